// ORACLE — test infrastructure, not product code (see philox.hpp).
// examples/coin_toss.rs restated: component :24-28, spawner :58-63, system :67-81,
// aggregate :33-49.
#pragma once
#include <vector>

#include "philox.hpp"

namespace oracle
{

struct CoinTosser
{
    uint64_t num_heads = 0;
    uint64_t num_tosses = 0;
};

struct CoinTossSim
{
    uint64_t seed;
    uint32_t replica;
    double odds_heads;
    uint64_t step = 1; // StepNumber seen by the first user step (simulation_builder.rs:50)
    std::vector<CoinTosser> tossers;

    CoinTossSim(uint64_t seed_, uint32_t replica_, uint64_t n, double odds) : seed(seed_), replica(replica_), odds_heads(odds), tossers(n) {}

    // the closure system of coin_toss.rs:67-81 with the rng call replaced by the keyed draw
    void system_toss()
    {
        const Stream rng(seed, replica, T_COIN);
        const uint64_t thr = threshold(odds_heads);
        uint32_t uid = 0;
        for (CoinTosser& t : tossers)
        {
            const U4 b = rng.block(uid >> 2, 0, static_cast<uint32_t>(step), 0);
            const bool coin_toss_heads = bern(b.w[uid & 3], thr);
            if (coin_toss_heads) t.num_heads += 1;
            t.num_tosses += 1;
            ++uid;
        }
    }
    void run(uint64_t n)
    {
        for (uint64_t i = 0; i < n; ++i)
        {
            system_toss();
            step += 1; // Last: StepNumber += 1 (step_number.rs:20-23)
        }
    }
    // coin_toss.rs:33-49
    double sample_aggregate() const
    {
        double odds_sum = 0.0;
        for (const CoinTosser& t : tossers) odds_sum += static_cast<double>(t.num_heads) / static_cast<double>(t.num_tosses);
        return odds_sum / static_cast<double>(tossers.size());
    }
};

} // namespace oracle
