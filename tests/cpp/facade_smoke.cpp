// The reference's tests/test_counter.rs:154-193 and examples/coin_toss.rs through the C++ façade.
// Exit code 0 = all assertions hold (needs a CUDA device); 77 = no device (build check only).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "incerto_b200.hpp"

using namespace incerto;

#define REQUIRE(x)                                                  \
    do                                                              \
    {                                                               \
        if (!(x))                                                   \
        {                                                           \
            std::fprintf(stderr, "%s:%d: %s failed\n", __FILE__, __LINE__, #x); \
            return 1;                                               \
        }                                                           \
    } while (0)

int main()
{
    try
    {
        // test_counter_time_series (tests/test_counter.rs:154-193)
        SimulationBuilder b;
        Component counter = b.component("MyCounter", INCERTO_U64);
        const uint64_t zero = 0;
        b.add_entity_spawner([&](Spawner& s) { s.spawn_batch(1, {{counter, &zero}}); });
        incerto_add_const_params p{};
        p.target = counter.handle;
        p.amount = 1;
        b.add_systems(INCERTO_SYS_ADD_CONST, p);
        b.record_aggregate_time_series(Sum(counter), 10);
        bool conflict = false;
        try
        {
            b.record_aggregate_time_series(Sum(counter), 12); // tests/test_builder.rs:27-42
        }
        catch (const SimulationError& e)
        {
            conflict = e.is(BuilderError::TimeSeriesRecordingConflict);
        }
        REQUIRE(conflict);
        Simulation sim = b.build();
        sim.run(100);
        REQUIRE(sim.sample_aggregate<uint64_t>(Sum(counter)) == 100);
        auto ts = sim.get_aggregate_time_series<uint64_t>(Sum(counter));
        REQUIRE(ts.len() == 10 && ts.values()[0] == 10 && ts.values()[9] == 100 && ts.duration() == 100);
        sim.run(10);
        ts = sim.get_aggregate_time_series<uint64_t>(Sum(counter));
        REQUIRE(ts.len() == 11 && ts.values()[10] == 110);
        REQUIRE(sim.sample_single<uint64_t>(counter) == 110);

        // examples/coin_toss.rs as-is
        SimulationBuilder c;
        c.set_seed(0x1CE27002025ull);
        Component heads = c.component("num_heads", INCERTO_U32), tosses = c.component("num_tosses", INCERTO_U32);
        std::vector<uint32_t> z(100, 0);
        c.add_entity_spawner([&](Spawner& s) { s.spawn_batch(100, {{heads, z.data()}, {tosses, z.data()}}); });
        incerto_coin_toss_params cp{heads.handle, tosses.handle, 0.5};
        c.add_systems(INCERTO_SYS_COIN_TOSS, cp);
        Simulation cs = c.build();
        cs.run(10000);
        const double odds = cs.sample_aggregate<double>(CoinOdds(heads, tosses));
        REQUIRE(odds > 0.49 && odds < 0.51);
        REQUIRE(cs.sample_aggregate<uint32_t>(Minimum(tosses)) == 10000);
        // examples/air_traffic_3d.rs check_conflicts, RNG-free: face neighbours conflict, same-cell and diagonal pairs do not
        SimulationBuilder ab;
        Component apos = ab.component("GridPosition3D", INCERTO_I16, 3), atgt = ab.component("Aircraft.target_altitude", INCERTO_I16);
        const int32_t bounds[6] = {0, 0, 0, 20, 20, 10};
        ab.add_spatial_grid(3, apos, atgt, bounds);
        const int16_t xyz[7 * 3] = {5, 5, 5, 5, 5, 6, 1, 1, 1, 1, 1, 1, 9, 9, 2, 10, 10, 2, 3, 7, 0};
        const int16_t tg[7] = {5, 6, 1, 1, 2, 2, 0};
        ab.add_entity_spawner([&](Spawner& s) { s.spawn_batch(7, {{apos, xyz}, {atgt, tg}}); });
        incerto_air_params ap{apos.handle, atgt.handle, 20, 10, 0.7};
        ab.add_systems(INCERTO_SYS_AIR_CHECK_CONFLICTS, ap);
        Simulation as = ab.build();
        as.run(1);
        REQUIRE(as.sample_aggregate<uint64_t>(AirConflicts(atgt)) == 1);
        REQUIRE(as.sample_aggregate<uint64_t>(Count(atgt)) == 7);

        // examples/traders.rs, RNG-free: p_buy = 1 with two shares at 5.0 exhausts the cash after ten purchases
        SimulationBuilder tb;
        incerto_traders_params tp{};
        Component tid = tb.component("TraderId", INCERTO_U64), cash = tb.component("Trader.cash", INCERTO_F64),
                  pf = tb.component("Trader.portfolio", INCERTO_U8), nw = tb.component("TraderNetWorth", INCERTO_F64),
                  sid = tb.component("StockId", INCERTO_U64), price = tb.component("StockPrice", INCERTO_F64),
                  delisted = tb.component("StockDelisted");
        tp.trader_id = tid.handle, tp.cash = cash.handle, tp.portfolio = pf.handle, tp.net_worth = nw.handle;
        tp.stock_id = sid.handle, tp.stock_price = price.handle, tp.stock_delisted = delisted.handle;
        tp.chance_buy = 1.0, tp.buy_shares_min = 2, tp.buy_shares_max = 2;
        incerto_spawn_traders_params st{tid.handle, cash.handle, pf.handle, nw.handle, 3, 100.0, 25};
        incerto_spawn_stocks_params ss{sid.handle, price.handle, delisted.handle, 25, 5.0};
        tb.add_device_spawner(INCERTO_SPAWN_TRADERS, &st, sizeof(st));
        tb.add_device_spawner(INCERTO_SPAWN_STOCKS, &ss, sizeof(ss));
        tb.add_systems(INCERTO_SYS_TRADERS_BUY, tp);
        tb.add_systems(INCERTO_SYS_TRADERS_NET_WORTH, tp);
        Simulation tsim = tb.build();
        tsim.run(1);
        REQUIRE(tsim.sample_aggregate<double>(Maximum(cash)) == 0.0);
        REQUIRE(tsim.sample_aggregate<double>(Minimum(nw)) == 100.0);
        const std::vector<uint8_t> shares = tsim.read_component<uint8_t>(pf, 25);
        REQUIRE(shares.size() == 75 && shares[0] == 2 && shares[9] == 2 && shares[10] == 0 && shares[74] == 0);
        std::printf("facade smoke ok: odds %.4f\n", odds);
        return 0;
    }
    catch (const SimulationError& e)
    {
        if (e.code == INCERTO_ERR_NO_DEVICE) return 77;
        std::fprintf(stderr, "SimulationError: %s\n", e.what());
        return 2;
    }
}
