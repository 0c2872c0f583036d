// ORACLE — test infrastructure, not product code (see philox.hpp).
// examples/forest_fire.rs restated system by system, on the hash-map SpatialGrid of
// library.hpp, with every rng call replaced by the keyed draw of DESIGN.md §2.4.
//   CellState :48-62   spawn_forest_grid :234-279   fire_spread_system :282-329
//   burn_progression_system :332-351   regrowth_system :354-365   FireStats :104-134
// Pinned system order (the reference leaves it open, simulation_builder.rs:62-64):
//   spread -> burn progression -> regrowth.
#pragma once
#include <unordered_set>
#include <vector>

#include "library.hpp"
#include "lottery.hpp"
#include "philox.hpp"

namespace oracle
{

enum class CellKind : uint8_t
{
    Healthy,
    Burning,
    Burned,
    Empty
};

struct ForestCell
{
    CellKind kind = CellKind::Empty;
    uint32_t remaining_burn_time = 0;
};

struct FireStats
{
    uint64_t healthy_count = 0, burning_count = 0, burned_count = 0, empty_count = 0, total_cells = 0;
};

struct ForestParams
{
    int32_t width = 50, height = 50;          // :33-34
    double initial_forest_density = 0.7;      // :38
    double fire_spread_probability = 0.6;     // :39
    uint32_t burn_duration = 3;               // :40
    double regrowth_probability = 0.001;      // :41
    uint32_t initial_fire_count = 3;          // :42
    bool spread = true, progression = true, regrowth = true; // which systems were added
};

// device encoding used at the C ABI (include/incerto_b200.h)
inline uint8_t encode_cell(const ForestCell& c)
{
    switch (c.kind)
    {
    case CellKind::Healthy: return 0x40;
    case CellKind::Empty: return 0x80;
    case CellKind::Burned: return 0;
    default: return static_cast<uint8_t>(c.remaining_burn_time);
    }
}
inline ForestCell decode_cell(uint8_t v)
{
    ForestCell c;
    if (v == 0x40)
        c.kind = CellKind::Healthy;
    else if (v == 0x80)
        c.kind = CellKind::Empty;
    else if (v == 0)
        c.kind = CellKind::Burned;
    else
    {
        c.kind = CellKind::Burning;
        c.remaining_burn_time = v;
    }
    return c;
}

struct ForestSim
{
    ForestParams P;
    uint64_t seed;
    uint32_t replica;
    uint64_t step = 1;
    // entities in spawn order: uid == index
    std::vector<Pos> position;
    std::vector<ForestCell> cell;
    SpatialGrid grid;
    std::vector<FireStats> series; // record_aggregate_time_series::<ForestCell, FireStats>(1), :168

    ForestSim(const ForestParams& p, uint64_t seed_, uint32_t replica_) : P(p), seed(seed_), replica(replica_)
    {
        grid.dim = 2;
        grid.bounds.some = true; // :151-154
        grid.bounds.min = Pos{0, 0, 0};
        grid.bounds.max = Pos{P.width - 1, P.height - 1, 0};
    }

    // spawn_forest_grid, :234-279
    void spawn_forest_grid()
    {
        const Stream rng_spawn(seed, replica, T_FSPW);
        const Stream rng_ignite(seed, replica, T_FIGN);
        const uint64_t thr = threshold(P.initial_forest_density);
        for (int32_t x = 0; x < P.width; ++x)
            for (int32_t y = 0; y < P.height; ++y)
            {
                const uint32_t uid = static_cast<uint32_t>(position.size());
                const U4 b = rng_spawn.block(uid >> 2, 0, 0, 0);
                ForestCell c;
                c.kind = bern(b.w[uid & 3], thr) ? CellKind::Healthy : CellKind::Empty;
                position.push_back(Pos{x, y, 0});
                cell.push_back(c);
            }
        // `healthy_positions` holds every grid position in x-major order (:260-263)
        const uint32_t n_positions = static_cast<uint32_t>(P.width) * static_cast<uint32_t>(P.height);
        for (uint32_t k = 0; k < P.initial_fire_count; ++k)
        {
            const U4 b = rng_ignite.block(k, 0, 0, 0);
            const uint32_t idx = below(b.w[0], n_positions); // choose(&mut rng)
            ForestCell burning;
            burning.kind = CellKind::Burning;
            burning.remaining_burn_time = P.burn_duration;
            position.push_back(Pos{static_cast<int32_t>(idx / P.height), static_cast<int32_t>(idx % P.height), 0});
            cell.push_back(burning);
        }
    }
    void spawn_explicit(const std::vector<Pos>& pos, const std::vector<uint8_t>& states)
    {
        for (size_t i = 0; i < pos.size(); ++i)
        {
            position.push_back(pos[i]);
            cell.push_back(decode_cell(states[i]));
        }
    }

    // PreUpdate maintenance (spatial_grid.rs:434-443): positions never change in this example,
    // so every entity is inserted once, during the first step.
    void spatial_grid_update_system()
    {
        if (grid.num_entities() == position.size()) return;
        for (Entity e = 0; e < position.size(); ++e) grid.insert_or_update(e, position[e]);
    }

    uint32_t base_cells() const { return static_cast<uint32_t>(P.width) * static_cast<uint32_t>(P.height); }

    // one spread draw for (healthy target T, burning source S)
    bool spread_draw(const Stream& rng, uint64_t thr, Entity target, Entity source) const
    {
        if (source < base_cells())
        {
            // direction of S as seen from T: 0 N(0,-1), 1 W(-1,0), 2 E(+1,0), 3 S(0,+1)
            const int dx = position[source].x - position[target].x, dy = position[source].y - position[target].y;
            const int d = (dx == 0 && dy == -1) ? 0 : (dx == -1 && dy == 0) ? 1 : (dx == 1 && dy == 0) ? 2 : 3;
            const U4 b = rng.block(target, 0, static_cast<uint32_t>(step), 0);
            return bern(b.w[d], thr);
        }
        const U4 b = rng.block(target, source, static_cast<uint32_t>(step), 1);
        return bern(b.w[0], thr);
    }

    // fire_spread_system, :282-329
    void fire_spread_system()
    {
        const Stream rng(seed, replica, T_FSPR);
        const uint64_t thr = threshold(P.fire_spread_probability);
        std::unordered_set<Pos, PosHash> spread_positions;
        for (Entity burning_entity = 0; burning_entity < position.size(); ++burning_entity)
        {
            if (cell[burning_entity].kind != CellKind::Burning) continue;
            grid.orthogonal_neighbors_of(position[burning_entity], [&](Entity neighbor_entity) {
                if (cell[neighbor_entity].kind == CellKind::Healthy)
                    if (spread_draw(rng, thr, neighbor_entity, burning_entity)) spread_positions.insert(position[neighbor_entity]);
            });
        }
        for (Entity e = 0; e < position.size(); ++e)
            if (spread_positions.count(position[e]))
            {
                cell[e].kind = CellKind::Burning;
                cell[e].remaining_burn_time = P.burn_duration;
            }
    }

    // burn_progression_system, :332-351
    void burn_progression_system()
    {
        for (ForestCell& c : cell)
            if (c.kind == CellKind::Burning)
            {
                if (c.remaining_burn_time > 1)
                    c.remaining_burn_time -= 1;
                else
                    c.kind = CellKind::Burned;
            }
    }

    // regrowth draw of entity uid: two-level when p*256 <= 1
    mutable LotteryCache regrow_lottery;
    bool regrow_draw(const Stream& rng, Entity uid) const
    {
        const double p = P.regrowth_probability;
        if (p * 256.0 <= 1.0)
        {
            // level 1: uid is a winner of its block's lottery (256 independent Bernoulli(2^-8) draws, one call)
            if (!regrow_lottery.wins(rng, static_cast<uint32_t>(uid), 0, static_cast<uint32_t>(step))) return false;
            const U4 l2 = rng.block(uid >> 2, 0, static_cast<uint32_t>(step), 1);
            return bern(l2.w[uid & 3], threshold(p * 256.0));
        }
        const U4 l2 = rng.block(uid >> 2, 0, static_cast<uint32_t>(step), 1);
        return bern(l2.w[uid & 3], threshold(p));
    }

    // regrowth_system, :354-365 (`&&` short-circuits: only Empty cells draw)
    void regrowth_system()
    {
        const Stream rng(seed, replica, T_FREG);
        for (Entity e = 0; e < cell.size(); ++e)
            if (cell[e].kind == CellKind::Empty && regrow_draw(rng, e)) cell[e].kind = CellKind::Healthy;
    }

    FireStats sample_aggregate() const // :104-134
    {
        FireStats s;
        for (const ForestCell& c : cell)
        {
            switch (c.kind)
            {
            case CellKind::Healthy: s.healthy_count += 1; break;
            case CellKind::Burning: s.burning_count += 1; break;
            case CellKind::Burned: s.burned_count += 1; break;
            case CellKind::Empty: s.empty_count += 1; break;
            }
        }
        s.total_cells = cell.size();
        return s;
    }

    void run(uint64_t n)
    {
        for (uint64_t i = 0; i < n; ++i)
        {
            spatial_grid_update_system();               // PreUpdate
            if (P.spread) fire_spread_system();         // Update, pinned order
            if (P.progression) burn_progression_system();
            if (P.regrowth) regrowth_system();
            if (!cell.empty()) series.push_back(sample_aggregate()); // PostUpdate, interval 1 (time_series.rs:68-87)
            step += 1;                                  // Last
        }
    }
};

} // namespace oracle
