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
