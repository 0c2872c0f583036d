// ORACLE — test infrastructure, not product code (see philox.hpp).
//
// CPU restatement of the incerto library pieces the example systems rely on:
//   SpatialGrid<T, C>      src/plugins/spatial_grid.rs:194-389 (HashMap<pos, HashSet<Entity>> + reverse map)
//   GridCoordinates        src/plugins/spatial_grid.rs:55-109  (neighbour enumeration orders)
//   GridBounds::contains   src/plugins/spatial_grid.rs:208-242 (inclusive on both ends)
//   aggregators            src/util/aggregate.rs:73-192
// Structure-faithful on purpose (hash maps, one serial pass per system): it doubles as the
// CPU baseline that stands in for the Bevy run (no Rust toolchain in this environment).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace oracle
{

struct Pos
{
    int32_t x = 0, y = 0, z = 0;
    bool operator==(const Pos& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct PosHash
{
    size_t operator()(const Pos& p) const
    {
        uint64_t h = static_cast<uint32_t>(p.x) * 0x9E3779B97F4A7C15ull;
        h ^= (static_cast<uint64_t>(static_cast<uint32_t>(p.y)) + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
        h ^= (static_cast<uint64_t>(static_cast<uint32_t>(p.z)) + 0x165667B1ull) * 0x9E3779B185EBCA87ull;
        return static_cast<size_t>(h ^ (h >> 29));
    }
};

struct Bounds
{
    bool some = false;
    Pos min, max;
    // spatial_grid.rs:208-242
    bool contains(const Pos& p, int dim) const
    {
        if (min.x > max.x || min.y > max.y || (dim == 3 && min.z > max.z)) throw std::runtime_error("bounds: min > max");
        return p.x >= min.x && p.x <= max.x && p.y >= min.y && p.y <= max.y && (dim == 2 || (p.z >= min.z && p.z <= max.z));
    }
};

// spatial_grid.rs:55-109
inline std::vector<Pos> neighbors(const Pos& p, int dim)
{
    std::vector<Pos> r;
    if (dim == 2)
    {
        // NORTH_WEST, NORTH, NORTH_EAST, WEST, EAST, SOUTH_WEST, SOUTH, SOUTH_EAST with NORTH = (0,-1)
        const int d[8][2] = {{-1, -1}, {0, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {0, 1}, {1, 1}};
        for (auto& o : d) r.push_back(Pos{p.x + o[0], p.y + o[1], 0});
    }
    else
    {
        for (int dx = -1; dx <= 1; ++dx)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dz = -1; dz <= 1; ++dz)
                    if (dx || dy || dz) r.push_back(Pos{p.x + dx, p.y + dy, p.z + dz});
    }
    return r;
}
inline std::vector<Pos> neighbors_orthogonal(const Pos& p, int dim)
{
    std::vector<Pos> r;
    if (dim == 2)
    {
        const int d[4][2] = {{0, -1}, {-1, 0}, {1, 0}, {0, 1}}; // NORTH, WEST, EAST, SOUTH
        for (auto& o : d) r.push_back(Pos{p.x + o[0], p.y + o[1], 0});
    }
    else
    {
        const int d[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}}; // W E N S DOWN UP
        for (auto& o : d) r.push_back(Pos{p.x + o[0], p.y + o[1], p.z + o[2]});
    }
    return r;
}

using Entity = uint32_t;

class SpatialGrid
{
  public:
    int dim = 2;
    Bounds bounds;
    std::unordered_map<Pos, std::unordered_set<Entity>, PosHash> position_to_entities;
    std::unordered_map<Entity, Pos> entity_to_position;

    bool in_bounds(const Pos& p) const { return !bounds.some || bounds.contains(p, dim); } // :265-271
    void insert_or_update(Entity e, const Pos& p)                                         // :274-290
    {
        if (!in_bounds(p)) throw std::out_of_range("entity outside spatial grid bounds");
        remove(e);
        position_to_entities[p].insert(e);
        entity_to_position[e] = p;
    }
    bool remove(Entity e) // :295-311
    {
        auto it = entity_to_position.find(e);
        if (it == entity_to_position.end()) return false;
        auto pit = position_to_entities.find(it->second);
        if (pit == position_to_entities.end()) throw std::logic_error("entity found in one hashmap but not the other?");
        pit->second.erase(e);
        if (pit->second.empty()) position_to_entities.erase(pit);
        entity_to_position.erase(it);
        return true;
    }
    template <typename F>
    void entities_at(const Pos& p, F&& f) const // :314-320
    {
        auto it = position_to_entities.find(p);
        if (it == position_to_entities.end()) return;
        for (Entity e : it->second) f(e);
    }
    size_t count_at(const Pos& p) const
    {
        auto it = position_to_entities.find(p);
        return it == position_to_entities.end() ? 0 : it->second.size();
    }
    template <typename F>
    void neighbors_of(const Pos& p, F&& f) const // :331-350
    {
        for (const Pos& q : neighbors(p, dim))
            if (in_bounds(q)) entities_at(q, f);
    }
    template <typename F>
    void orthogonal_neighbors_of(const Pos& p, F&& f) const // :355-372
    {
        for (const Pos& q : neighbors_orthogonal(p, dim))
            if (in_bounds(q)) entities_at(q, f);
    }
    bool is_empty(const Pos& p) const { return count_at(p) == 0; }      // :376-381
    size_t num_entities() const { return entity_to_position.size(); }   // :385-388
};

// ---- aggregators, src/util/aggregate.rs
template <typename T>
T agg_minimum(const std::vector<T>& v) // :73-88
{
    if (v.empty()) throw std::runtime_error("sample_aggregate called with empty slice");
    T m = v[0];
    for (const T& x : v)
    {
        if (x != x) throw std::runtime_error("cannot compare values");
        if (x < m) m = x;
    }
    return m;
}
template <typename T>
T agg_maximum(const std::vector<T>& v) // :90-105
{
    if (v.empty()) throw std::runtime_error("sample_aggregate called with empty slice");
    T m = v[0];
    for (const T& x : v)
    {
        if (x != x) throw std::runtime_error("cannot compare values");
        if (x > m) m = x;
    }
    return m;
}
template <typename T>
T agg_mean(const std::vector<T>& v) // :157-192: sum and division in T
{
    if (v.empty()) throw std::runtime_error("empty");
    T sum = T(0);
    for (const T& x : v) sum = static_cast<T>(sum + x);
    const T cnt = static_cast<T>(v.size());
    return static_cast<T>(sum / cnt);
}
template <typename T>
T agg_percentile(std::vector<T> v, unsigned P) // :107-135
{
    if (v.empty()) throw std::runtime_error("empty");
    for (const T& x : v)
        if (x != x) throw std::runtime_error("cannot compare values");
    std::sort(v.begin(), v.end());
    const double percentage = static_cast<double>(P) / 100.0;
    const double idx = std::floor(percentage * static_cast<double>(v.size()));
    const size_t i = static_cast<size_t>(idx);
    if (i >= v.size()) throw std::out_of_range("percentile index out of range");
    return v[i];
}
template <typename T>
T agg_median(std::vector<T> v) // :137-155
{
    if (v.empty()) throw std::runtime_error("empty");
    for (const T& x : v)
        if (x != x) throw std::runtime_error("cannot compare values");
    std::sort(v.begin(), v.end());
    return v[v.size() / 2];
}

} // namespace oracle
