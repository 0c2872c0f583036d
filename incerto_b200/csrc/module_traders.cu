// Traders module: host side of kernels_traders.cu (examples/traders.rs).
#include <algorithm>
#include <cmath>

#include "kernels_traders.hpp"
#include "modules.hpp"

namespace incerto
{

static bool valid_comp(const Sim& s, int c) { return c >= 0 && c < static_cast<int>(s.comps.size()); }

// Trader.portfolio is engine-managed: 4 bits per stock, word-major.  Its width depends on the number of
// stocks, known once the spawners have been counted: called by run_spawners between counting and allocation.
int32_t traders_layout(Sim& s, const std::vector<uint64_t>& rows)
{
    TradersModule& tm = s.mods->traders;
    int portfolio = -1, stock_price = -1;
    uint32_t num_stocks = 0;
    for (auto& sys : s.systems)
        if (sys.kind >= INCERTO_SYS_TRADERS_BUY && sys.kind <= INCERTO_SYS_STOCKS_PRICE_CHANGE &&
            sys.params.size() == sizeof(incerto_traders_params))
        {
            incerto_traders_params p;
            std::memcpy(&p, sys.params.data(), sizeof(p));
            portfolio = p.portfolio;
            stock_price = p.stock_price;
        }
    for (auto& sp : s.spawns)
        if (sp.device_kind && sp.kind == INCERTO_SPAWN_TRADERS && sp.params.size() == sizeof(incerto_spawn_traders_params))
        {
            incerto_spawn_traders_params p;
            std::memcpy(&p, sp.params.data(), sizeof(p));
            portfolio = p.portfolio;
            num_stocks = std::max(num_stocks, p.num_stocks);
        }
    if (!valid_comp(s, portfolio)) return INCERTO_OK;
    if (valid_comp(s, stock_price))
        for (size_t ti = 0; ti < s.tables.size() && ti < rows.size(); ++ti)
            if (s.tables[ti].has(stock_price)) num_stocks = std::max<uint32_t>(num_stocks, static_cast<uint32_t>(std::min<uint64_t>(rows[ti], 1u << 20)));
    if (num_stocks == 0) num_stocks = 1;
    if (num_stocks > kTradersMaxStocks)
    {
        set_error("traders: at most %u stocks per replica (portfolio words are register-resident)", kTradersMaxStocks);
        return INCERTO_ERR_UNSUPPORTED;
    }
    tm.num_stocks_cap = num_stocks;
    s.comps[portfolio].managed_bytes = 4 * ((num_stocks + 7) / 8);
    for (auto& t : s.tables)
        if (Column* c = t.col(portfolio)) c->row_bytes = s.comps[portfolio].managed_bytes;
    return INCERTO_OK;
}

int32_t traders_bind(Sim& s)
{
    TradersModule& tm = s.mods->traders;
    bool first = true;
    for (auto& sys : s.systems)
    {
        if (sys.kind < INCERTO_SYS_TRADERS_BUY || sys.kind > INCERTO_SYS_STOCKS_PRICE_CHANGE) continue;
        if (sys.params.size() != sizeof(incerto_traders_params)) return INCERTO_ERR_INVALID_ARGUMENT;
        incerto_traders_params p;
        std::memcpy(&p, sys.params.data(), sizeof(p));
        for (int c : {p.trader_id, p.cash, p.portfolio, p.net_worth, p.stock_id, p.stock_price, p.stock_delisted})
            if (!valid_comp(s, c)) return INCERTO_ERR_INVALID_ARGUMENT;
        if (s.comps[p.cash].dtype != INCERTO_F64 || s.comps[p.net_worth].dtype != INCERTO_F64 ||
            s.comps[p.stock_price].dtype != INCERTO_F64 || s.comps[p.stock_delisted].dtype != INCERTO_MARKER ||
            s.comps[p.cash].lanes != 1 || s.comps[p.net_worth].lanes != 1 || s.comps[p.stock_price].lanes != 1)
        {
            set_error("traders: cash / net worth / stock price must be F64 columns and StockDelisted a marker");
            return INCERTO_ERR_INVALID_ARGUMENT;
        }
        if (!first && std::memcmp(&tm.p.trader_id, &p.trader_id, 7 * sizeof(incerto_component)) != 0)
        {
            set_error("traders systems must agree on their components");
            return INCERTO_ERR_INVALID_ARGUMENT;
        }
        const incerto_traders_params keep = tm.p;
        tm.p = p;
        if (!first)
        {
            // every system carries only the constants it uses; keep the ones seen earlier
            if (sys.kind != INCERTO_SYS_TRADERS_BUY)
            {
                tm.p.chance_buy = keep.chance_buy;
                tm.p.buy_shares_min = keep.buy_shares_min;
                tm.p.buy_shares_max = keep.buy_shares_max;
            }
            if (sys.kind != INCERTO_SYS_TRADERS_SELL) tm.p.chance_sell = keep.chance_sell;
            if (sys.kind != INCERTO_SYS_STOCKS_PRICE_CHANGE)
            {
                tm.p.price_change_mean = keep.price_change_mean;
                tm.p.price_change_std_dev = keep.price_change_std_dev;
            }
        }
        first = false;
        switch (sys.kind)
        {
        case INCERTO_SYS_TRADERS_BUY: tm.do_buy = true; break;
        case INCERTO_SYS_TRADERS_SELL: tm.do_sell = true; break;
        case INCERTO_SYS_TRADERS_NET_WORTH: tm.do_net_worth = true; break;
        default: tm.do_price = true; break;
        }
        for (int c : {p.trader_id, p.cash, p.portfolio, p.net_worth, p.stock_id, p.stock_price, p.stock_delisted})
            s.comps[c].exists = true; // the systems' queries register their components
    }
    if (tm.do_buy && (tm.p.buy_shares_min < 1 || tm.p.buy_shares_max > 15 || tm.p.buy_shares_min > tm.p.buy_shares_max))
    {
        set_error("traders: TRADER_BUY_NUM_SHARES must lie within 1..=15 (4-bit share counts; 0 means not owned)");
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    tm.active = true;
    tm.table_traders = table_for_components(s, {tm.p.cash, tm.p.portfolio, tm.p.net_worth});
    tm.table_stocks = table_for_components(s, {tm.p.stock_price});
    if (tm.table_stocks >= 0 && s.tables[tm.table_stocks].n > kTradersMaxStocks)
    {
        set_error("traders: at most %u stocks per replica", kTradersMaxStocks);
        return INCERTO_ERR_UNSUPPORTED;
    }
    return INCERTO_OK;
}

// Level 1 of a buy / sell decision: for p <= 1/32 the block lottery with the largest k in 5..8 such
// that 2^k p <= 1 (`kexp`, one Philox call per trader and step), level 2 then draws against 2^k p; above 1/32 one byte
// per stock below T1 = min(256, ceil(256 p)) (`kexp` = 0), level 2 against 256 p / T1.
static void two_level(double p, uint32_t& t1, uint64_t& thr2, uint32_t& kexp)
{
    kexp = 0;
    if (!(p > 0.0))
    {
        t1 = 0;
        thr2 = 0;
        return;
    }
    for (uint32_t k = 8; k >= 5; --k)
        if (p * static_cast<double>(1u << k) <= 1.0)
        {
            kexp = k;
            t1 = 0;
            thr2 = bernoulli_threshold(p * static_cast<double>(1u << k));
            return;
        }
    const double c = std::ceil(256.0 * p);
    t1 = c >= 256.0 ? 256u : static_cast<uint32_t>(c);
    thr2 = bernoulli_threshold(256.0 * p / static_cast<double>(t1));
}

static bool traders_fill(Sim& s, TradersModule& tm, TradersArgs& a, uint64_t step)
{
    std::memset(&a, 0, sizeof(a));
    if (tm.table_stocks < 0) return false;
    Table& st = s.tables[tm.table_stocks];
    a.price = static_cast<double*>(st.col(tm.p.stock_price)->d);
    a.stock_flags = st.d_flags;
    a.num_stocks = static_cast<uint32_t>(st.n);
    a.bit_delisted = 1u << s.comps[tm.p.stock_delisted].marker_bit;
    a.step = static_cast<uint32_t>(step);
    if (tm.table_traders >= 0)
    {
        Table& t = s.tables[tm.table_traders];
        a.cash = static_cast<double*>(t.col(tm.p.cash)->d);
        a.net_worth = static_cast<double*>(t.col(tm.p.net_worth)->d);
        a.portfolio = static_cast<uint32_t*>(t.col(tm.p.portfolio)->d);
        a.trader_flags = t.all_alive ? nullptr : t.d_flags;
        a.trader_uid = t.d_uid;
        a.n = t.n;
        a.n_cap = t.n_spawned;
    }
    a.do_buy = tm.do_buy;
    a.do_sell = tm.do_sell;
    a.do_net_worth = tm.do_net_worth;
    two_level(tm.p.chance_buy, a.buy_t1, a.buy_thr2, a.buy_kexp);
    two_level(tm.p.chance_sell, a.sell_t1, a.sell_thr2, a.sell_kexp);
    a.shares_min = tm.p.buy_shares_min;
    a.shares_range = tm.p.buy_shares_max - tm.p.buy_shares_min + 1;
    a.change_mean = tm.p.price_change_mean;
    a.change_std_dev = tm.p.price_change_std_dev;
    a.rk_buy = make_round_keys(make_key(s.seed, s.replica, TAG_TRADERS_BUY));
    a.rk_sell = make_round_keys(make_key(s.seed, s.replica, TAG_TRADERS_SELL));
    a.rk_price = make_round_keys(make_key(s.seed, s.replica, TAG_STOCKS_PRICE));
    return true;
}

int32_t traders_step(Sim& s, uint64_t first_step, uint32_t nsteps)
{
    TradersModule& tm = s.mods->traders;
    for (uint32_t i = 0; i < nsteps; ++i)
    {
        TradersArgs a;
        if (!traders_fill(s, tm, a, first_step + i)) return INCERTO_OK; // no stocks: every query is empty
        if (a.n && (tm.do_buy || tm.do_sell || tm.do_net_worth))
        {
            const double words = (a.num_stocks + 7) / 8;
            KTimer kt(s, "traders_step", (8.0 * words + 24.0) * a.n);
            launch_traders_step(a, s.stream);
        }
        if (tm.do_price && a.num_stocks)
        {
            KTimer kt(s, "stocks_price_change", 17.0 * a.num_stocks);
            launch_stocks_price_change(a, s.stream);
        }
        INC_CUDA(cudaGetLastError());
    }
    return INCERTO_OK;
}

// INCERTO_SPAWN_TRADERS (examples/traders.rs:116-127) / INCERTO_SPAWN_STOCKS (:130-137)
int32_t traders_spawn_device(Sim& s, const HostSpawn& sp, Table& t, uint64_t row_offset)
{
    if (sp.kind == INCERTO_SPAWN_TRADERS)
    {
        incerto_spawn_traders_params p;
        std::memcpy(&p, sp.params.data(), sizeof(p));
        if (row_offset != 0)
        {
            set_error("the traders device spawner must be the only spawner of its archetype");
            return INCERTO_ERR_INVALID_ARGUMENT;
        }
        if (s.comps[p.trader_id].dtype != INCERTO_U64 || s.comps[p.cash].dtype != INCERTO_F64 ||
            s.comps[p.net_worth].dtype != INCERTO_F64)
        {
            set_error("traders spawner: TraderId must be U64, cash and net worth F64");
            return INCERTO_ERR_INVALID_ARGUMENT;
        }
        const uint64_t n = p.num_traders;
        KTimer kt(s, "traders_spawn", (24.0 + t.col(p.portfolio)->row_bytes) * n);
        launch_iota_u64(static_cast<uint64_t*>(t.col(p.trader_id)->d), n, s.stream);
        launch_fill_f64(static_cast<double*>(t.col(p.cash)->d), p.initial_cash, n, s.stream);
        launch_fill_f64(static_cast<double*>(t.col(p.net_worth)->d), p.initial_cash, n, s.stream);
        INC_CUDA(cudaMemsetAsync(t.col(p.portfolio)->d, 0, t.col(p.portfolio)->row_bytes * t.n_spawned, s.stream));
        INC_CUDA(cudaMemsetAsync(t.d_flags, kAliveBit, n, s.stream));
        INC_CUDA(cudaGetLastError());
        return INCERTO_OK;
    }
    incerto_spawn_stocks_params p;
    std::memcpy(&p, sp.params.data(), sizeof(p));
    if (s.comps[p.stock_id].dtype != INCERTO_U64 || s.comps[p.stock_price].dtype != INCERTO_F64)
    {
        set_error("stocks spawner: StockId must be U64 and StockPrice F64");
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    if (row_offset != 0)
    {
        set_error("the stocks device spawner must be the only spawner of its archetype");
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    launch_iota_u64(static_cast<uint64_t*>(t.col(p.stock_id)->d), p.num_stocks, s.stream);
    launch_fill_f64(static_cast<double*>(t.col(p.stock_price)->d), p.initial_price, p.num_stocks, s.stream);
    INC_CUDA(cudaMemsetAsync(t.d_flags, kAliveBit, p.num_stocks, s.stream));
    INC_CUDA(cudaGetLastError());
    return INCERTO_OK;
}

// read_component of the managed portfolio: one byte (#shares) per stock, stocks_cap bytes per live trader
int32_t traders_read_portfolio(Sim& s, int table, uint8_t* out, size_t out_size, size_t* written)
{
    TradersModule& tm = s.mods->traders;
    Table& t = s.tables[table];
    Column* c = t.col(tm.p.portfolio);
    const uint32_t ns = tm.num_stocks_cap, words = (ns + 7) / 8;
    std::vector<uint32_t> h(static_cast<size_t>(words) * t.n_spawned);
    std::vector<uint32_t> rows, uids;
    INC_TRY(live_rows_by_uid(s, t, rows, uids));
    INC_CUDA(cudaMemcpyAsync(h.data(), c->d, h.size() * 4, cudaMemcpyDeviceToHost, s.stream));
    INC_CUDA(cudaStreamSynchronize(s.stream));
    size_t w = 0;
    for (uint32_t r : rows)
    {
        if (w + ns > out_size)
        {
            set_error("read_component: output buffer too small");
            return INCERTO_ERR_INVALID_ARGUMENT;
        }
        for (uint32_t k = 0; k < ns; ++k) out[w + k] = (h[static_cast<size_t>(k >> 3) * t.n_spawned + r] >> (4 * (k & 7))) & 0xfu;
        w += ns;
    }
    *written = w;
    return INCERTO_OK;
}

} // namespace incerto
