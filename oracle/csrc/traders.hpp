// ORACLE — test infrastructure, not product code (see philox.hpp).
// examples/traders.rs restated system by system; rng calls replaced by the keyed draws of DESIGN.md §2.
//   Trader :62-67  spawn_traders :116-127  spawn_stocks :130-137  stocks_price_change :140-163
//   traders_may_buy_shares :165-197  traders_may_sell_shares :199-236  traders_calculate_net_worth :239-263
// Pinned order: buy -> sell -> net worth -> price change, commands (insert StockDelisted)
// applied at the end of Update.  Stocks are visited in StockId order (the reference iterates a table query
// for buying and a HashMap for selling / net worth; any order is a legal run, the f64 results differ in
// the last bits and - through the affordability test - in rare decisions).
// Structure-faithful: portfolios are hash maps, price tables are rebuilt per system as in :206-210, :245-249.
#pragma once
#include <algorithm>
#include <limits>
#include <unordered_map>
#include <vector>

#include "detmath.hpp"
#include "lottery.hpp"
#include "philox.hpp"

namespace oracle
{

struct TradersParams
{
    uint64_t num_traders = 10;     // :36
    uint32_t num_stocks = 100;     // :37
    double initial_cash = 100.0;   // :39
    double chance_buy = 0.01;      // :40
    double chance_sell = 0.005;    // :41
    uint32_t shares_min = 1, shares_max = 5; // :42
    double initial_price = 5.0;    // :44
    double change_mean = 0.001;    // :45
    double change_std_dev = 0.2;   // :46
    bool buy = true, sell = true, net_worth = true, price_change = true;
};

// two-level Bernoulli of the contract: level 1 = one byte < T1 with T1 = min(256, ceil(256 p)),
// level 2 = a 32-bit draw below floor(256 p / T1 * 2^32)
// (for p <= 1/32 level 1 is the block lottery with P = 2^-k, the largest k in 5..8 such that 2^k p <= 1 - one call
// per trader and step, lottery.hpp - and level 2 draws against 2^k p)
struct TwoLevel
{
    uint32_t t1;
    uint64_t thr2;
    int kexp = 0;
    // level 1 for stock s of trader t
    bool level1(const Stream& rng, uint32_t t, uint32_t s, uint32_t step, LotteryCache& cache) const
    {
        if (kexp) return cache.wins_at(rng, t, s, 0, step, kexp); // block t, position s
        const U4 b = rng.block(t, s >> 4, step, 0);
        return ((b.w[(s & 15) >> 2] >> (8 * (s & 3))) & 0xffu) < t1;
    }
    explicit TwoLevel(double p)
    {
        if (!(p > 0.0))
        {
            t1 = 0;
            thr2 = 0;
            return;
        }
        for (int k = 8; k >= 5; --k)
            if (p * static_cast<double>(1u << k) <= 1.0)
            {
                kexp = k;
                t1 = 0;
                thr2 = threshold(p * static_cast<double>(1u << k));
                return;
            }
        const double c = std::ceil(256.0 * p);
        t1 = c >= 256.0 ? 256u : static_cast<uint32_t>(c);
        thr2 = threshold(256.0 * p / static_cast<double>(t1));
    }
};

struct TradersSim
{
    TradersParams P;
    uint64_t seed;
    uint32_t replica;
    uint64_t step = 1;
    struct Trader
    {
        double cash;
        std::unordered_map<uint32_t, uint32_t> portfolio; // StockId -> #shares
    };
    std::vector<Trader> traders;
    std::vector<double> net_worth;
    std::vector<double> price;
    std::vector<uint8_t> delisted;
    std::vector<std::vector<double>> series; // per step: net worth of every trader (record_time_series, :97)

    TradersSim(const TradersParams& p, uint64_t s, uint32_t r) : P(p), seed(s), replica(r) {}

    void spawn_traders() // :116-127
    {
        for (uint64_t i = 0; i < P.num_traders; ++i)
        {
            traders.push_back(Trader{P.initial_cash, {}});
            net_worth.push_back(P.initial_cash);
        }
    }
    void spawn_stocks() // :130-137
    {
        for (uint32_t i = 0; i < P.num_stocks; ++i)
        {
            price.push_back(P.initial_price);
            delisted.push_back(0);
        }
    }

    static uint32_t byte_of(const U4& b, uint32_t k) { return (b.w[k >> 2] >> (8 * (k & 3))) & 0xffu; }

    void traders_may_buy_shares() // :165-197
    {
        const Stream rng(seed, replica, T_TBUY);
        const TwoLevel tl(P.chance_buy);
        LotteryCache lottery;
        for (uint32_t t = 0; t < traders.size(); ++t)
        {
            Trader& tr = traders[t];
            for (uint32_t s = 0; s < price.size(); ++s)
            {
                if (delisted[s]) continue; // Without<StockDelisted>
                const double p = price[s];
                const bool already_owned = tr.portfolio.count(s) != 0;
                bool decides_to_buy = tl.level1(rng, t, s, static_cast<uint32_t>(step), lottery);
                U4 l2{};
                if (decides_to_buy)
                {
                    l2 = rng.block(t, s, static_cast<uint32_t>(step), 1);
                    decides_to_buy = bern(l2.w[0], tl.thr2);
                }
                if (!already_owned && decides_to_buy)
                {
                    const uint32_t num_shares = P.shares_min + below(l2.w[1], P.shares_max - P.shares_min + 1);
                    const double amount = static_cast<double>(num_shares) * p;
                    if (amount > tr.cash) continue;
                    tr.portfolio[s] = num_shares;
                    tr.cash -= amount;
                }
            }
        }
    }

    void traders_may_sell_shares() // :199-236
    {
        const Stream rng(seed, replica, T_TSEL);
        const TwoLevel tl(P.chance_sell);
        LotteryCache lottery;
        std::unordered_map<uint32_t, double> stock_prices;
        for (uint32_t s = 0; s < price.size(); ++s) stock_prices[s] = price[s];
        for (uint32_t t = 0; t < traders.size(); ++t)
        {
            Trader& tr = traders[t];
            std::vector<uint32_t> owned;
            for (auto& kv : tr.portfolio) owned.push_back(kv.first);
            std::sort(owned.begin(), owned.end()); // pinned iteration order
            std::vector<std::pair<uint32_t, uint32_t>> to_sell;
            for (uint32_t s : owned)
            {
                bool decides_to_sell = tl.level1(rng, t, s, static_cast<uint32_t>(step), lottery);
                if (decides_to_sell) decides_to_sell = bern(rng.block(t, s, static_cast<uint32_t>(step), 1).w[0], tl.thr2);
                if (decides_to_sell) to_sell.push_back({s, tr.portfolio[s]});
            }
            for (auto& [s, n] : to_sell)
            {
                tr.portfolio.erase(s);
                tr.cash += static_cast<double>(n) * stock_prices[s];
            }
        }
    }

    void traders_calculate_net_worth() // :239-263
    {
        std::unordered_map<uint32_t, double> stock_prices;
        for (uint32_t s = 0; s < price.size(); ++s) stock_prices[s] = price[s];
        for (uint32_t t = 0; t < traders.size(); ++t)
        {
            std::vector<uint32_t> owned;
            for (auto& kv : traders[t].portfolio) owned.push_back(kv.first);
            std::sort(owned.begin(), owned.end());
            double portfolio_value = 0.0;
            for (uint32_t s : owned) portfolio_value += static_cast<double>(traders[t].portfolio[s]) * stock_prices[s];
            net_worth[t] = traders[t].cash + portfolio_value;
        }
    }

    void stocks_price_change(std::vector<uint32_t>& commands) // :140-163
    {
        const Stream rng(seed, replica, T_TPRC);
        for (uint32_t s = 0; s < price.size(); ++s)
        {
            if (delisted[s]) continue;
            const U4 b = rng.block(s, 0, static_cast<uint32_t>(step), 0);
            const double price_change = P.change_mean + P.change_std_dev * det_normal(b.w[0], b.w[1], b.w[2], b.w[3]);
            price[s] = std::max(price[s] + price_change, 0.0);
            if (price[s] < std::numeric_limits<double>::epsilon()) commands.push_back(s);
        }
    }

    void run(uint64_t n)
    {
        for (uint64_t k = 0; k < n; ++k)
        {
            std::vector<uint32_t> commands;
            if (P.buy) traders_may_buy_shares();
            if (P.sell) traders_may_sell_shares();
            if (P.net_worth) traders_calculate_net_worth();
            if (P.price_change) stocks_price_change(commands);
            for (uint32_t s : commands) delisted[s] = 1;
            series.push_back(net_worth); // PostUpdate sample, interval 1
            step += 1;
        }
    }
};

} // namespace oracle
