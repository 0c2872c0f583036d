// ORACLE — test infrastructure, not product code (see philox.hpp).
// examples/pandemic_spatial.rs restated system by system on the structure-faithful SpatialGrid of library.hpp;
// rng calls replaced by the keyed draws of DESIGN.md §2.
//   Person :74-79  Quarantined :82-86  ContactHistory :89-93  PandemicStats :96-138  spawn_population :279-313
//   people_move_with_social_distancing :316-411   disease_incubation_progression :414-433
//   spatial_disease_transmission :436-518         disease_recovery_and_death :521-541
//   update_contact_history :544-562               process_contact_tracing :565-608
//   update_quarantine_status :611-625
// Pinned order: incubation -> transmission -> recovery/death -> contact history -> contact
// tracing -> quarantine status -> move, Commands (despawn, try_insert Quarantined, remove Quarantined) applied
// in issue order at the end of Update.  With the move last, the PreUpdate grid snapshot and the positions
// agree for every system that pairs them.
#pragma once
#include <set>
#include <vector>

#include "library.hpp"
#include "philox.hpp"

namespace oracle
{

struct SpatialParams
{
    uint64_t initial_population = 2000; // :29
    int32_t grid_size = 40;             // :30
    double chance_start_infected = 0.02; // :33
    int32_t infection_radius = 2;       // :34
    double chance_infect_1 = 0.15;      // :35
    double chance_infect_2 = 0.05;      // :36
    double chance_recover = 0.03;       // :37
    double chance_die = 0.001;          // :38
    uint32_t incubation_period = 5;     // :39
    uint32_t crowding_threshold = 8;    // :43
    double social_compliance = 0.7;     // :44
    uint32_t quarantine_duration = 14;  // :48
    int32_t zone_x = 20, zone_y = 20, zone_radius = 8; // :52-55
    uint32_t history_clear_above = 20;  // :555
    double chance_attempt_move = 0.5;   // :332
    uint32_t systems = 0x7f; // bit k: system kind 40 + k enabled
};

enum : uint8_t
{
    D_HEALTHY = 0,
    D_INFECTIOUS = 0x80,
    D_RECOVERED = 0x81 // Exposed{n} = n (1..=127)
};

struct SpatialSim
{
    SpatialParams P;
    uint64_t seed;
    uint32_t replica;
    uint64_t step = 1;
    // entities in spawn order; despawned entities keep their slot with alive = false
    std::vector<Pos> position;
    std::vector<uint8_t> alive, disease, social;
    std::vector<uint32_t> quarantined; // remaining_duration, 0 = component absent
    std::vector<std::set<std::pair<int32_t, int32_t>>> history;
    std::vector<uint8_t> changed, removed;
    SpatialGrid grid;
    std::vector<std::vector<uint64_t>> series; // PandemicStats per step

    SpatialSim(const SpatialParams& p, uint64_t s, uint32_t r) : P(p), seed(s), replica(r)
    {
        grid.dim = 2;
        grid.bounds.some = true;
        grid.bounds.min = Pos{0, 0, 0};
        grid.bounds.max = Pos{p.grid_size - 1, p.grid_size - 1, 0}; // :166-169
    }

    void push_person(int32_t x, int32_t y, uint8_t d, uint8_t sd)
    {
        position.push_back(Pos{x, y, 0});
        alive.push_back(1);
        disease.push_back(d);
        social.push_back(sd);
        quarantined.push_back(0);
        history.emplace_back();
        changed.push_back(1);
        removed.push_back(0);
    }
    void spawn_population() // :279-313
    {
        const Stream rng(seed, replica, T_SSPW);
        for (uint64_t i = 0; i < P.initial_population; ++i)
        {
            const U4 b = rng.block(static_cast<uint32_t>(i), 0, 0, 0);
            const bool sd = bern(b.w[2], threshold(P.social_compliance));
            const bool exposed = bern(b.w[3], threshold(P.chance_start_infected));
            push_person(static_cast<int32_t>(below(b.w[0], P.grid_size)), static_cast<int32_t>(below(b.w[1], P.grid_size)),
                        exposed ? static_cast<uint8_t>(P.incubation_period) : static_cast<uint8_t>(D_HEALTHY), sd);
        }
    }

    // PreUpdate: spatial_grid_update_system then spatial_grid_cleanup_system (spatial_grid.rs:434-455)
    void grid_maintenance()
    {
        for (uint32_t e = 0; e < position.size(); ++e)
        {
            if (alive[e] && changed[e])
            {
                grid.insert_or_update(e, position[e]);
                changed[e] = 0;
            }
            if (removed[e])
            {
                grid.remove(e);
                removed[e] = 0;
            }
        }
    }

    struct Cmd
    {
        int kind; // 0 despawn, 1 try_insert Quarantined, 2 remove Quarantined
        uint32_t e;
    };

    void disease_incubation_progression() // :414-433
    {
        for (uint32_t e = 0; e < position.size(); ++e)
        {
            if (!alive[e]) continue;
            uint8_t& d = disease[e];
            if (d >= 1 && d < 0x80) d = d > 1 ? d - 1 : D_INFECTIOUS;
        }
    }

    void spatial_disease_transmission() // :436-518
    {
        const Stream rng(seed, replica, T_STRN);
        const int32_t R = P.infection_radius;
        std::vector<uint32_t> new_exposures;
        for (uint32_t s = 0; s < position.size(); ++s)
        {
            if (!alive[s] || quarantined[s] || disease[s] != D_INFECTIOUS) continue;
            std::vector<uint32_t> nearby;
            for (int32_t dx = -R; dx <= R; ++dx)
                for (int32_t dy = -R; dy <= R; ++dy)
                    if (std::abs(dx) + std::abs(dy) <= R)
                        grid.entities_at(Pos{position[s].x + dx, position[s].y + dy, 0}, [&](Entity e) { nearby.push_back(e); });
            for (uint32_t t : nearby)
            {
                if (t == s) continue;
                if (!alive[t] || quarantined[t]) continue; // query.get fails: Without<Quarantined>
                if (disease[t] != D_HEALTHY) continue;
                const int32_t dist = std::abs(position[s].x - position[t].x) + std::abs(position[s].y - position[t].y);
                const double chance = dist <= 1 ? P.chance_infect_1 : (dist == 2 ? P.chance_infect_2 : 0.0);
                if (bern(rng.block(t, s, static_cast<uint32_t>(step), 0).w[0], threshold(chance))) new_exposures.push_back(t);
            }
        }
        for (uint32_t t : new_exposures) disease[t] = static_cast<uint8_t>(P.incubation_period);
    }

    void disease_recovery_and_death(std::vector<Cmd>& commands) // :521-541
    {
        const Stream rng(seed, replica, T_SFAT);
        for (uint32_t e = 0; e < position.size(); ++e)
        {
            if (!alive[e] || disease[e] != D_INFECTIOUS) continue;
            const U4 b = rng.block(e, 0, static_cast<uint32_t>(step), 0);
            if (bern(b.w[0], threshold(P.chance_die)))
                commands.push_back({0, e});
            else if (bern(b.w[1], threshold(P.chance_recover)))
                disease[e] = D_RECOVERED;
        }
    }

    void update_contact_history() // :544-562
    {
        for (uint32_t e = 0; e < position.size(); ++e)
        {
            if (!alive[e]) continue;
            history[e].insert({position[e].x, position[e].y});
            if (history[e].size() > 14)
                if (history[e].size() > P.history_clear_above) history[e].clear();
        }
    }

    void process_contact_tracing(std::vector<Cmd>& commands) // :565-608
    {
        for (uint32_t a = 0; a < position.size(); ++a)
        {
            if (!alive[a] || quarantined[a]) continue;
            for (auto& loc : history[a])
                grid.entities_at(Pos{loc.first, loc.second, 0}, [&](Entity b) {
                    if (b == a) return;
                    if (alive[b] && !quarantined[b]) commands.push_back({1, b});
                });
        }
    }

    void update_quarantine_status(std::vector<Cmd>& commands) // :611-625
    {
        for (uint32_t e = 0; e < position.size(); ++e)
        {
            if (!alive[e] || !quarantined[e]) continue;
            if (quarantined[e] > 1)
                quarantined[e] -= 1;
            else
                commands.push_back({2, e});
        }
    }

    void people_move_with_social_distancing() // :316-411
    {
        const Stream rng(seed, replica, T_SMOV);
        const int32_t G = P.grid_size;
        for (uint32_t e = 0; e < position.size(); ++e)
        {
            if (!alive[e] || quarantined[e]) continue;
            const U4 r = rng.block(e, 0, static_cast<uint32_t>(step), 0);
            if (!bern(r.w[0], threshold(P.chance_attempt_move))) continue;
            const Pos p = position[e];
            const Pos directions[4] = {Pos{p.x, p.y - 1, 0}, Pos{p.x - 1, p.y, 0}, Pos{p.x + 1, p.y, 0}, Pos{p.x, p.y + 1, 0}};
            std::vector<Pos> best_moves;
            size_t min_crowding = ~static_cast<size_t>(0);
            for (const Pos& q : directions)
            {
                if (q.x < 0 || q.x >= G || q.y < 0 || q.y >= G) continue;
                if (std::abs(q.x - P.zone_x) + std::abs(q.y - P.zone_y) <= P.zone_radius)
                    if (social[e]) continue;
                const size_t people = grid.count_at(q);
                if (social[e])
                {
                    if (people < min_crowding)
                    {
                        min_crowding = people;
                        best_moves.clear();
                        best_moves.push_back(q);
                    }
                    else if (people == min_crowding)
                        best_moves.push_back(q);
                }
                else
                    best_moves.push_back(q);
            }
            if (!best_moves.empty())
            {
                const Pos chosen = best_moves[below(r.w[1], static_cast<uint32_t>(best_moves.size()))];
                if (grid.count_at(chosen) < P.crowding_threshold)
                {
                    position[e] = chosen;
                    changed[e] = 1;
                }
            }
        }
    }

    void stats(uint64_t out[6]) const // :107-138
    {
        for (int i = 0; i < 6; ++i) out[i] = 0;
        for (uint32_t e = 0; e < position.size(); ++e)
        {
            if (!alive[e]) continue;
            const uint8_t d = disease[e];
            out[d == D_HEALTHY ? 0 : (d == D_INFECTIOUS ? 2 : (d == D_RECOVERED ? 3 : 1))] += 1;
            out[5] += 1;
        }
        out[4] = P.initial_population - out[5];
    }

    void run(uint64_t n)
    {
        for (uint64_t k = 0; k < n; ++k)
        {
            grid_maintenance();
            std::vector<Cmd> commands;
            if (P.systems & 1) disease_incubation_progression();
            if (P.systems & 2) spatial_disease_transmission();
            if (P.systems & 4) disease_recovery_and_death(commands);
            if (P.systems & 8) update_contact_history();
            if (P.systems & 16) process_contact_tracing(commands);
            if (P.systems & 32) update_quarantine_status(commands);
            if (P.systems & 64) people_move_with_social_distancing();
            for (const Cmd& c : commands)
            {
                if (!alive[c.e]) continue; // try_insert / remove on a despawned entity: no-op
                if (c.kind == 0)
                {
                    alive[c.e] = 0;
                    removed[c.e] = 1;
                }
                else if (c.kind == 1)
                    quarantined[c.e] = P.quarantine_duration;
                else
                    quarantined[c.e] = 0;
            }
            std::vector<uint64_t> st(6);
            stats(st.data());
            if (st[5]) series.push_back(st); // skipped when no Person is left (time_series.rs:79)
            step += 1;
        }
    }
};

} // namespace oracle
