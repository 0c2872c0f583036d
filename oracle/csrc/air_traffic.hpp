// ORACLE — test infrastructure, not product code (see philox.hpp).
// examples/air_traffic_3d.rs restated system by system on the structure-faithful SpatialGrid of
// library.hpp; rng calls replaced by the keyed draws of DESIGN.md §2.
//   Aircraft :21-28   spawn_aircraft :61-90   move_aircraft :93-121   check_conflicts :124-162
// Pinned order: check_conflicts -> move_aircraft, so that the snapshot grid and the
// positions agree inside a step.  The grid is maintained in PreUpdate (spatial_grid.rs:418-455): during
// step k it holds the positions as of the end of step k-1.
#pragma once
#include <vector>

#include "library.hpp"
#include "philox.hpp"

namespace oracle
{

struct AirParams
{
    uint64_t num_aircraft = 15;     // :18
    int32_t airspace_size = 20;     // :16
    int32_t airspace_height = 10;   // :17
    double chance_horizontal_move = 0.7; // :108
    bool check = true, move = true;
    bool declaration_order = false; // move first (the order of the tuple at :217): stale grid, fresh positions
};

struct AirSim
{
    AirParams P;
    uint64_t seed;
    uint32_t replica;
    uint64_t step = 1;
    std::vector<Pos> position;          // GridPosition3D
    std::vector<int32_t> target_altitude; // Aircraft.target_altitude
    std::vector<uint8_t> aircraft_type; // Aircraft.aircraft_type (display only)
    std::vector<uint8_t> changed;       // bevy Changed<GridPosition3D>
    SpatialGrid grid;
    uint64_t conflicts = 0;             // `conflicts` of the last check_conflicts run (each pair counted twice)
    std::vector<uint64_t> series;       // conflicts / 2 per step

    AirSim(const AirParams& p, uint64_t s, uint32_t r) : P(p), seed(s), replica(r)
    {
        grid.dim = 3;
        grid.bounds.some = true;
        grid.bounds.min = Pos{0, 0, 0};
        grid.bounds.max = Pos{p.airspace_size, p.airspace_size, p.airspace_height}; // :209-212 (inclusive max)
    }

    void spawn_aircraft() // :61-90
    {
        const Stream rng(seed, replica, T_ASPW);
        for (uint64_t i = 0; i < P.num_aircraft; ++i)
        {
            const U4 a = rng.block(static_cast<uint32_t>(i), 0, 0, 0);
            const U4 b = rng.block(static_cast<uint32_t>(i), 0, 0, 1);
            Pos p;
            p.x = static_cast<int32_t>(below(a.w[0], P.airspace_size));
            p.y = static_cast<int32_t>(below(a.w[1], P.airspace_size));
            p.z = static_cast<int32_t>(below(a.w[2], P.airspace_height));
            aircraft_type.push_back(static_cast<uint8_t>(below(a.w[3], 3)));
            target_altitude.push_back(static_cast<int32_t>(below(b.w[0], P.airspace_height)));
            position.push_back(p);
            changed.push_back(1);
        }
    }
    void spawn_explicit(uint64_t n, const int32_t* xyz, const int32_t* target)
    {
        for (uint64_t i = 0; i < n; ++i)
        {
            position.push_back(Pos{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]});
            target_altitude.push_back(target[i]);
            aircraft_type.push_back(0);
            changed.push_back(1);
        }
    }

    void spatial_grid_update_system() // spatial_grid.rs:434-443
    {
        for (uint32_t e = 0; e < position.size(); ++e)
            if (changed[e])
            {
                grid.insert_or_update(e, position[e]);
                changed[e] = 0;
            }
    }

    void move_aircraft() // :93-121
    {
        const Stream rng(seed, replica, T_AMOV);
        for (uint32_t e = 0; e < position.size(); ++e)
        {
            Pos& p = position[e];
            const Pos before = p;
            if (p.z < target_altitude[e])
                p.z += 1;
            else if (p.z > target_altitude[e])
                p.z -= 1;
            const U4 r = rng.block(e, 0, static_cast<uint32_t>(step), 0);
            if (bern(r.w[0], threshold(P.chance_horizontal_move)))
            {
                const int32_t dx = static_cast<int32_t>(below(r.w[1], 3)) - 1; // rng.random_range(-1..=1)
                const int32_t dy = static_cast<int32_t>(below(r.w[2], 3)) - 1;
                p.x = std::min(std::max(p.x + dx, 0), P.airspace_size - 1);
                p.y = std::min(std::max(p.y + dy, 0), P.airspace_size - 1);
                changed[e] = 1; // `*position = ...` marks the component changed even when the value is equal
            }
            if (!(p == before)) changed[e] = 1;
            // (DerefMut on the altitude branch also marks it changed; the grid update is idempotent either way)
        }
    }

    void check_conflicts() // :124-162
    {
        uint64_t c = 0;
        for (uint32_t e = 0; e < position.size(); ++e)
        {
            const Pos& p = position[e];
            grid.neighbors_of(p, [&](Entity other) {
                if (other == e) return;
                const Pos& q = position[other]; // query.get(nearby_entity): the *current* position
                const int64_t dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
                if (dx * dx + dy * dy + dz * dz <= 1) c += 1;
            });
        }
        conflicts = c;
    }

    void run(uint64_t n)
    {
        for (uint64_t s = 0; s < n; ++s)
        {
            spatial_grid_update_system(); // PreUpdate
            if (P.declaration_order)
            {
                if (P.move) move_aircraft();
                if (P.check) check_conflicts();
            }
            else
            {
                if (P.check) check_conflicts();
                if (P.move) move_aircraft();
            }
            series.push_back(conflicts / 2); // :160 "Divide by 2 since each conflict is counted twice"
            step += 1;
        }
    }
};

} // namespace oracle
