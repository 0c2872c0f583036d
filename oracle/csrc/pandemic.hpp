// ORACLE — test infrastructure, not product code (see philox.hpp).
// examples/pandemic.rs restated system by system; rng calls replaced by the keyed draws of DESIGN.md §2.5.
//   Position :39-44   spawn_people :84-107   people_move :115-166   people_infect_others :169-192
//   people_infected_may_die :195-206   people_infected_may_recover :209-223   counts :74-78
// Pinned order: move -> infect -> die -> recover; Commands (insert Infected / despawn / remove Infected)
// are deferred to the end of Update and applied in issue order (bevy semantics).
#pragma once
#include <unordered_map>
#include <vector>

#include "philox.hpp"

namespace oracle
{

struct PandemicParams
{
    uint64_t initial_population = 1000; // :19
    uint32_t grid_size = 50;            // :20
    double chance_start_infected = 0.05; // :22
    double chance_infect_other = 0.02;  // :26
    double chance_recover = 0.001;      // :29
    double chance_die = 0.0005;         // :32
    bool move = true, infect = true, die = true, recover = true;
    bool nested_loop = true; // the reference's O(I x H) double loop; false = per-cell buckets (same pairs)
};

struct PandemicSim
{
    PandemicParams P;
    uint64_t seed;
    uint32_t replica;
    uint64_t step = 1;
    // entities in spawn order (uid = index); despawned entities keep their slot with alive = false
    std::vector<uint32_t> x, y;
    std::vector<uint8_t> alive, infected;

    PandemicSim(const PandemicParams& p, uint64_t s, uint32_t r) : P(p), seed(s), replica(r) {}

    void spawn_people() // :84-107
    {
        const Stream rng(seed, replica, T_PSPW);
        const uint64_t thr = threshold(P.chance_start_infected);
        for (uint64_t i = 0; i < P.initial_population; ++i)
        {
            const U4 b = rng.block(static_cast<uint32_t>(i), 0, 0, 0);
            x.push_back(below(b.w[0], P.grid_size)); // rng.random_range(0..GRID_SIZE)
            y.push_back(below(b.w[1], P.grid_size));
            infected.push_back(bern(b.w[2], thr) ? 1 : 0);
            alive.push_back(1);
        }
    }
    void spawn_explicit(uint64_t n, const uint16_t* xy, const uint8_t* inf)
    {
        for (uint64_t i = 0; i < n; ++i)
        {
            x.push_back(xy[2 * i]);
            y.push_back(xy[2 * i + 1]);
            infected.push_back(inf[i]);
            alive.push_back(1);
        }
    }

    void people_move() // :115-166
    {
        const Stream rng(seed, replica, T_PMOV);
        const uint32_t G = P.grid_size;
        for (uint32_t e = 0; e < x.size(); ++e)
        {
            if (!alive[e]) continue;
            // the three p = 1/2 draws of a person share one nibble (DESIGN.md §2 "nibbles"): nibble e & 7 of word
            // (e >> 3) & 3 of the call (e >> 5); draw k is true iff bit 3-k of the nibble is clear
            const uint32_t w = rng.block(e >> 5, 0, static_cast<uint32_t>(step), 0).w[(e >> 3) & 3];
            const uint32_t nib = (w >> (4 * (e & 7))) & 0xfu;
            const bool stay = !(nib & 8u), x_axis = !(nib & 4u), negative = !(nib & 2u);
            if (stay)
            {
                // do nothing
            }
            else if (x_axis)
            {
                if (negative)
                {
                    if (x[e] > 0) x[e] -= 1;
                }
                else if (x[e] < G - 1)
                    x[e] += 1;
            }
            else
            {
                if (negative)
                {
                    if (y[e] > 0) y[e] -= 1;
                }
                else if (y[e] < G - 1)
                    y[e] += 1;
            }
        }
    }

    struct Cmd
    {
        int kind; // 0 insert Infected, 1 despawn, 2 remove Infected
        uint32_t e;
    };

    void people_infect_others(std::vector<Cmd>& commands) // :169-192
    {
        const Stream rng(seed, replica, T_PINF);
        const uint64_t thr = threshold(P.chance_infect_other);
        if (P.nested_loop)
        {
            for (uint32_t i = 0; i < x.size(); ++i)
            {
                if (!alive[i] || !infected[i]) continue;
                for (uint32_t h = 0; h < x.size(); ++h)
                {
                    if (!alive[h] || infected[h]) continue;
                    if (x[i] == x[h] && y[i] == y[h])
                        if (bern(rng.block(h, i, static_cast<uint32_t>(step), 0).w[0], thr)) commands.push_back({0, h});
                }
            }
            return;
        }
        std::unordered_map<uint64_t, std::vector<uint32_t>> cells;
        for (uint32_t i = 0; i < x.size(); ++i)
            if (alive[i] && infected[i]) cells[(static_cast<uint64_t>(x[i]) << 32) | y[i]].push_back(i);
        for (uint32_t h = 0; h < x.size(); ++h)
        {
            if (!alive[h] || infected[h]) continue;
            auto it = cells.find((static_cast<uint64_t>(x[h]) << 32) | y[h]);
            if (it == cells.end()) continue;
            for (uint32_t i : it->second)
                if (bern(rng.block(h, i, static_cast<uint32_t>(step), 0).w[0], thr)) commands.push_back({0, h});
        }
    }

    bool fate_two_level() const
    {
        return (!P.die || P.chance_die * 256.0 <= 1.0) && (!P.recover || P.chance_recover * 256.0 <= 1.0);
    }
    // which: 0 die, 1 recover
    bool fate_draw(const Stream& rng, uint32_t e, int which) const
    {
        const bool two = fate_two_level();
        const double p = which == 0 ? P.chance_die : P.chance_recover;
        if (two)
        {
            const U4 l1 = rng.block(e >> 3, 0, static_cast<uint32_t>(step), 0);
            const uint32_t b = 2 * (e & 7) + which;
            if ((l1.w[b >> 2] >> (8 * (b & 3))) & 0xffu) return false;
        }
        const U4 l2 = rng.block(e, 0, static_cast<uint32_t>(step), 1);
        return bern(l2.w[which], threshold(two ? p * 256.0 : p));
    }
    void people_infected_may_die(std::vector<Cmd>& commands) // :195-206
    {
        const Stream rng(seed, replica, T_PFAT);
        for (uint32_t e = 0; e < x.size(); ++e)
            if (alive[e] && infected[e] && fate_draw(rng, e, 0)) commands.push_back({1, e});
    }
    void people_infected_may_recover(std::vector<Cmd>& commands) // :209-223
    {
        const Stream rng(seed, replica, T_PFAT);
        for (uint32_t e = 0; e < x.size(); ++e)
            if (alive[e] && infected[e] && fate_draw(rng, e, 1)) commands.push_back({2, e});
    }

    void run(uint64_t n)
    {
        for (uint64_t s = 0; s < n; ++s)
        {
            std::vector<Cmd> commands;
            if (P.move) people_move();
            if (P.infect) people_infect_others(commands);
            if (P.die) people_infected_may_die(commands);
            if (P.recover) people_infected_may_recover(commands);
            // end of Update: deferred commands, in issue order; remove/insert on a despawned entity is a no-op
            for (const Cmd& c : commands)
            {
                if (!alive[c.e]) continue;
                if (c.kind == 0)
                    infected[c.e] = 1;
                else if (c.kind == 1)
                    alive[c.e] = 0;
                else
                    infected[c.e] = 0;
            }
            step += 1;
        }
    }
    // simulation.count::<With<Person>>(), ::<With<Infected>>(), ::<(With<Person>, Without<Infected>)>() (:74-78)
    void counts(uint64_t out[3]) const
    {
        out[0] = out[1] = out[2] = 0;
        for (size_t e = 0; e < x.size(); ++e)
            if (alive[e])
            {
                out[0] += 1;
                out[1] += infected[e] ? 1 : 0;
                out[2] += infected[e] ? 0 : 1;
            }
    }
};

} // namespace oracle
