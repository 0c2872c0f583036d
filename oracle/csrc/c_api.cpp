// ORACLE — test infrastructure, not product code (see philox.hpp).
// extern "C" surface for the ctypes-driven tests and bench.py's CPU baseline.
#include <cstring>
#include <thread>

#include "air_traffic.hpp"
#include "coin_toss.hpp"
#include "forest_fire.hpp"
#include "library.hpp"
#include "pandemic.hpp"
#include "pandemic_spatial.hpp"
#include "lottery.hpp"
#include "philox.hpp"
#include "traders.hpp"

using namespace oracle;

extern "C"
{

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    const U4 r = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    std::memcpy(out, r.w, 16);
}
uint64_t oracle_threshold(double p) { return threshold(p); }
// winners of one block lottery (lottery.hpp) under key (k0, k1); returns their number, positions into out[<= 48]
uint32_t oracle_lottery256(uint32_t k0, uint32_t k1, uint32_t block, uint32_t sub, uint32_t step, int kexp, uint8_t* out)
{
    Stream s(0, 0, 0);
    s.k0 = k0;
    s.k1 = k1;
    const std::vector<uint8_t> w = lottery256(s, block, sub, step, kexp);
    for (size_t i = 0; i < w.size(); ++i) out[i] = w[i];
    return static_cast<uint32_t>(w.size());
}
uint32_t oracle_below(uint32_t u, uint32_t n) { return below(u, n); }

// ---------------------------------------------------------------- coin toss
// heads/tosses: n u64 each, updated in place; returns the aggregate (coin_toss.rs:33-49)
double oracle_coin_toss_run(uint64_t seed, uint32_t replica, uint64_t n, double odds, uint64_t first_step,
                            uint64_t steps, uint64_t* heads, uint64_t* tosses)
{
    CoinTossSim sim(seed, replica, n, odds);
    sim.step = first_step;
    for (uint64_t i = 0; i < n; ++i)
    {
        sim.tossers[i].num_heads = heads[i];
        sim.tossers[i].num_tosses = tosses[i];
    }
    sim.run(steps);
    for (uint64_t i = 0; i < n; ++i)
    {
        heads[i] = sim.tossers[i].num_heads;
        tosses[i] = sim.tossers[i].num_tosses;
    }
    return sim.sample_aggregate();
}

// ------------------------------------------------------------- forest fire
struct ForestHandle
{
    ForestSim sim;
    ForestHandle(const ForestParams& p, uint64_t seed, uint32_t replica) : sim(p, seed, replica) {}
};

void* oracle_forest_new(uint64_t seed, uint32_t replica, int32_t width, int32_t height, double density,
                        double p_spread, uint32_t burn_duration, double p_regrow, uint32_t fire_count,
                        uint32_t systems_mask /* 1 spread, 2 progression, 4 regrowth */)
{
    ForestParams p;
    p.width = width;
    p.height = height;
    p.initial_forest_density = density;
    p.fire_spread_probability = p_spread;
    p.burn_duration = burn_duration;
    p.regrowth_probability = p_regrow;
    p.initial_fire_count = fire_count;
    p.spread = systems_mask & 1;
    p.progression = systems_mask & 2;
    p.regrowth = systems_mask & 4;
    return new ForestHandle(p, seed, replica);
}
void oracle_forest_free(void* h) { delete static_cast<ForestHandle*>(h); }
void oracle_forest_spawn(void* h) { static_cast<ForestHandle*>(h)->sim.spawn_forest_grid(); }
// explicit entities: n positions (x,y int32 pairs) and n state codes (device encoding)
void oracle_forest_spawn_explicit(void* h, uint64_t n, const int32_t* xy, const uint8_t* states)
{
    std::vector<Pos> pos(n);
    std::vector<uint8_t> st(states, states + n);
    for (uint64_t i = 0; i < n; ++i) pos[i] = Pos{xy[2 * i], xy[2 * i + 1], 0};
    static_cast<ForestHandle*>(h)->sim.spawn_explicit(pos, st);
}
void oracle_forest_run(void* h, uint64_t steps) { static_cast<ForestHandle*>(h)->sim.run(steps); }
uint64_t oracle_forest_num_entities(void* h) { return static_cast<ForestHandle*>(h)->sim.cell.size(); }
void oracle_forest_cells(void* h, uint8_t* states_out, int32_t* xy_out)
{
    ForestSim& s = static_cast<ForestHandle*>(h)->sim;
    for (size_t i = 0; i < s.cell.size(); ++i)
    {
        if (states_out) states_out[i] = encode_cell(s.cell[i]);
        if (xy_out)
        {
            xy_out[2 * i] = s.position[i].x;
            xy_out[2 * i + 1] = s.position[i].y;
        }
    }
}
void oracle_forest_stats(void* h, uint64_t out5[5])
{
    const FireStats f = static_cast<ForestHandle*>(h)->sim.sample_aggregate();
    out5[0] = f.healthy_count;
    out5[1] = f.burning_count;
    out5[2] = f.burned_count;
    out5[3] = f.empty_count;
    out5[4] = f.total_cells;
}
uint64_t oracle_forest_series_len(void* h) { return static_cast<ForestHandle*>(h)->sim.series.size(); }
void oracle_forest_series(void* h, uint64_t* out /* len * 5 */)
{
    ForestSim& s = static_cast<ForestHandle*>(h)->sim;
    for (size_t i = 0; i < s.series.size(); ++i)
    {
        out[5 * i + 0] = s.series[i].healthy_count;
        out[5 * i + 1] = s.series[i].burning_count;
        out[5 * i + 2] = s.series[i].burned_count;
        out[5 * i + 3] = s.series[i].empty_count;
        out[5 * i + 4] = s.series[i].total_cells;
    }
}

// CPU baseline helper: `threads` independent replicas of the same forest run side by side
// (Bevy itself runs these three conflicting systems serially on one core; independent
// replicas are the only multi-core axis).  Returns cell-updates performed.
uint64_t oracle_forest_bench(uint64_t seed, int32_t width, int32_t height, uint64_t steps, uint32_t threads)
{
    std::vector<std::thread> pool;
    std::vector<uint64_t> done(threads, 0);
    for (uint32_t t = 0; t < threads; ++t)
        pool.emplace_back([&, t]() {
            ForestParams p;
            p.width = width;
            p.height = height;
            ForestSim sim(p, seed, t);
            sim.spawn_forest_grid();
            sim.run(steps);
            done[t] = static_cast<uint64_t>(sim.cell.size()) * steps;
        });
    uint64_t total = 0;
    for (uint32_t t = 0; t < threads; ++t)
    {
        pool[t].join();
        total += done[t];
    }
    return total;
}

// ---------------------------------------------------------------- pandemic
void* oracle_pandemic_new(uint64_t seed, uint32_t replica, uint64_t population, uint32_t grid_size, double p_start,
                          double p_infect, double p_recover, double p_die, uint32_t systems_mask /* 1 move 2 infect 4 die 8 recover */,
                          uint32_t nested_loop)
{
    PandemicParams p;
    p.initial_population = population;
    p.grid_size = grid_size;
    p.chance_start_infected = p_start;
    p.chance_infect_other = p_infect;
    p.chance_recover = p_recover;
    p.chance_die = p_die;
    p.move = systems_mask & 1;
    p.infect = systems_mask & 2;
    p.die = systems_mask & 4;
    p.recover = systems_mask & 8;
    p.nested_loop = nested_loop != 0;
    return new PandemicSim(p, seed, replica);
}
void oracle_pandemic_free(void* h) { delete static_cast<PandemicSim*>(h); }
void oracle_pandemic_spawn(void* h) { static_cast<PandemicSim*>(h)->spawn_people(); }
void oracle_pandemic_spawn_explicit(void* h, uint64_t n, const uint16_t* xy, const uint8_t* infected)
{
    static_cast<PandemicSim*>(h)->spawn_explicit(n, xy, infected);
}
void oracle_pandemic_run(void* h, uint64_t steps) { static_cast<PandemicSim*>(h)->run(steps); }
uint64_t oracle_pandemic_num_slots(void* h) { return static_cast<PandemicSim*>(h)->x.size(); }
// per spawn slot: position, alive, infected
void oracle_pandemic_state(void* h, uint16_t* xy, uint8_t* alive, uint8_t* infected)
{
    PandemicSim& s = *static_cast<PandemicSim*>(h);
    for (size_t i = 0; i < s.x.size(); ++i)
    {
        xy[2 * i] = static_cast<uint16_t>(s.x[i]);
        xy[2 * i + 1] = static_cast<uint16_t>(s.y[i]);
        alive[i] = s.alive[i];
        infected[i] = s.infected[i];
    }
}
void oracle_pandemic_counts(void* h, uint64_t out3[3]) { static_cast<PandemicSim*>(h)->counts(out3); }


// ------------------------------------------------------------ air_traffic_3d
void* oracle_air_new(uint64_t seed, uint32_t replica, uint64_t n, int32_t size, int32_t height, double p_move,
                     uint32_t systems_mask /* 1 check_conflicts, 2 move */)
{
    AirParams p;
    p.num_aircraft = n;
    p.airspace_size = size;
    p.airspace_height = height;
    p.chance_horizontal_move = p_move;
    p.check = systems_mask & 1;
    p.move = systems_mask & 2;
    return new AirSim(p, seed, replica);
}
void oracle_air_free(void* h) { delete static_cast<AirSim*>(h); }
void oracle_air_spawn(void* h) { static_cast<AirSim*>(h)->spawn_aircraft(); }
void oracle_air_spawn_explicit(void* h, uint64_t n, const int32_t* xyz, const int32_t* target)
{
    static_cast<AirSim*>(h)->spawn_explicit(n, xyz, target);
}
int oracle_air_run(void* h, uint64_t steps)
{
    try
    {
        static_cast<AirSim*>(h)->run(steps);
    }
    catch (const std::out_of_range&)
    {
        return 34;
    }
    return 0;
}
uint64_t oracle_air_num(void* h) { return static_cast<AirSim*>(h)->position.size(); }
void oracle_air_state(void* h, int32_t* xyz, int32_t* target)
{
    AirSim& s = *static_cast<AirSim*>(h);
    for (size_t i = 0; i < s.position.size(); ++i)
    {
        xyz[3 * i] = s.position[i].x;
        xyz[3 * i + 1] = s.position[i].y;
        xyz[3 * i + 2] = s.position[i].z;
        target[i] = s.target_altitude[i];
    }
}
uint64_t oracle_air_series(void* h, uint64_t* out, uint64_t cap)
{
    AirSim& s = *static_cast<AirSim*>(h);
    for (size_t i = 0; i < s.series.size() && i < cap; ++i) out[i] = s.series[i];
    return s.series.size();
}

// -------------------------------------------------------- pandemic_spatial
// params: 12 doubles / ints flattened, see oracle/spatial.py
void* oracle_spatial_new(uint64_t seed, uint32_t replica, uint64_t population, int32_t grid_size, const double* prob /*6*/,
                         const int32_t* ints /*9*/, uint32_t systems)
{
    SpatialParams p;
    p.initial_population = population;
    p.grid_size = grid_size;
    p.chance_start_infected = prob[0];
    p.chance_infect_1 = prob[1];
    p.chance_infect_2 = prob[2];
    p.chance_recover = prob[3];
    p.chance_die = prob[4];
    p.social_compliance = prob[5];
    p.chance_attempt_move = prob[6];
    p.infection_radius = ints[0];
    p.incubation_period = static_cast<uint32_t>(ints[1]);
    p.crowding_threshold = static_cast<uint32_t>(ints[2]);
    p.quarantine_duration = static_cast<uint32_t>(ints[3]);
    p.zone_x = ints[4];
    p.zone_y = ints[5];
    p.zone_radius = ints[6];
    p.history_clear_above = static_cast<uint32_t>(ints[7]);
    p.systems = systems;
    return new SpatialSim(p, seed, replica);
}
void oracle_spatial_free(void* h) { delete static_cast<SpatialSim*>(h); }
void oracle_spatial_spawn(void* h) { static_cast<SpatialSim*>(h)->spawn_population(); }
void oracle_spatial_spawn_explicit(void* h, uint64_t n, const int32_t* xy, const uint8_t* disease, const uint8_t* social)
{
    SpatialSim& s = *static_cast<SpatialSim*>(h);
    for (uint64_t i = 0; i < n; ++i) s.push_person(xy[2 * i], xy[2 * i + 1], disease[i], social[i]);
}
int oracle_spatial_run(void* h, uint64_t steps)
{
    try
    {
        static_cast<SpatialSim*>(h)->run(steps);
    }
    catch (const std::out_of_range&)
    {
        return 34;
    }
    return 0;
}
uint64_t oracle_spatial_num_slots(void* h) { return static_cast<SpatialSim*>(h)->position.size(); }
// per spawn slot; history: 21 u32 per slot = len, then the packed (x | y << 16) keys in ascending order
void oracle_spatial_state(void* h, int32_t* xy, uint8_t* alive, uint8_t* disease, uint8_t* social, uint32_t* quarantined,
                          uint32_t* history)
{
    SpatialSim& s = *static_cast<SpatialSim*>(h);
    for (size_t i = 0; i < s.position.size(); ++i)
    {
        xy[2 * i] = s.position[i].x;
        xy[2 * i + 1] = s.position[i].y;
        alive[i] = s.alive[i];
        disease[i] = s.disease[i];
        social[i] = s.social[i];
        quarantined[i] = s.quarantined[i];
        if (history)
        {
            uint32_t* hrow = history + 21 * i;
            std::vector<uint32_t> keys;
            for (auto& loc : s.history[i]) keys.push_back(static_cast<uint32_t>(loc.first) | (static_cast<uint32_t>(loc.second) << 16));
            std::sort(keys.begin(), keys.end());
            hrow[0] = static_cast<uint32_t>(keys.size());
            for (size_t k = 0; k < 20; ++k) hrow[1 + k] = k < keys.size() ? keys[k] : 0;
        }
    }
}
void oracle_spatial_stats(void* h, uint64_t out6[6]) { static_cast<SpatialSim*>(h)->stats(out6); }
uint64_t oracle_spatial_series(void* h, uint64_t* out /* len * 6 */, uint64_t cap)
{
    SpatialSim& s = *static_cast<SpatialSim*>(h);
    for (size_t i = 0; i < s.series.size() && i < cap; ++i)
        for (int k = 0; k < 6; ++k) out[6 * i + k] = s.series[i][k];
    return s.series.size();
}

// ------------------------------------------------------------------ traders
void* oracle_traders_new(uint64_t seed, uint32_t replica, uint64_t num_traders, uint32_t num_stocks, double cash,
                         double p_buy, double p_sell, uint32_t shares_min, uint32_t shares_max, double price,
                         double mean, double std_dev, uint32_t systems_mask /* 1 buy 2 sell 4 net worth 8 price */)
{
    TradersParams p;
    p.num_traders = num_traders;
    p.num_stocks = num_stocks;
    p.initial_cash = cash;
    p.chance_buy = p_buy;
    p.chance_sell = p_sell;
    p.shares_min = shares_min;
    p.shares_max = shares_max;
    p.initial_price = price;
    p.change_mean = mean;
    p.change_std_dev = std_dev;
    p.buy = systems_mask & 1;
    p.sell = systems_mask & 2;
    p.net_worth = systems_mask & 4;
    p.price_change = systems_mask & 8;
    TradersSim* s = new TradersSim(p, seed, replica);
    s->spawn_traders();
    s->spawn_stocks();
    return s;
}
void oracle_traders_free(void* h) { delete static_cast<TradersSim*>(h); }
void oracle_traders_set_prices(void* h, const double* prices)
{
    TradersSim& s = *static_cast<TradersSim*>(h);
    for (size_t i = 0; i < s.price.size(); ++i) s.price[i] = prices[i];
}
void oracle_traders_run(void* h, uint64_t steps) { static_cast<TradersSim*>(h)->run(steps); }
// cash, net worth: num_traders each; shares: num_traders x num_stocks; price, delisted: num_stocks
void oracle_traders_state(void* h, double* cash, double* net_worth, uint8_t* shares, double* price, uint8_t* delisted)
{
    TradersSim& s = *static_cast<TradersSim*>(h);
    const size_t ns = s.price.size();
    for (size_t t = 0; t < s.traders.size(); ++t)
    {
        cash[t] = s.traders[t].cash;
        net_worth[t] = s.net_worth[t];
        for (size_t k = 0; k < ns; ++k) shares[t * ns + k] = 0;
        for (auto& kv : s.traders[t].portfolio) shares[t * ns + kv.first] = static_cast<uint8_t>(kv.second);
    }
    for (size_t k = 0; k < ns; ++k)
    {
        price[k] = s.price[k];
        delisted[k] = s.delisted[k];
    }
}
// net worth of one trader at every recorded step
uint64_t oracle_traders_series(void* h, uint64_t trader, double* out, uint64_t cap)
{
    TradersSim& s = *static_cast<TradersSim*>(h);
    for (size_t i = 0; i < s.series.size() && i < cap; ++i) out[i] = s.series[i][trader];
    return s.series.size();
}
double oracle_det_log(double x) { return det_log(x); }
double oracle_det_cos2pi(double u) { return det_cos2pi(u); }
double oracle_det_normal(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) { return det_normal(w0, w1, w2, w3); }

// ------------------------------------------------------- library: aggregators
#define ORACLE_AGG(T, NAME)                                                                          \
    int oracle_aggregate_##NAME(const T* v, uint64_t n, uint32_t kind, uint32_t P, T* out)           \
    {                                                                                                \
        try                                                                                          \
        {                                                                                            \
            std::vector<T> x(v, v + n);                                                              \
            switch (kind)                                                                            \
            {                                                                                        \
            case 3: *out = agg_minimum<T>(x); break;                                                 \
            case 4: *out = agg_maximum<T>(x); break;                                                 \
            case 5: *out = agg_mean<T>(x); break;                                                    \
            case 6: *out = agg_median<T>(x); break;                                                  \
            case 7: *out = agg_percentile<T>(x, P); break;                                           \
            default: return -1;                                                                      \
            }                                                                                        \
            return 0;                                                                                \
        }                                                                                            \
        catch (const std::out_of_range&)                                                             \
        {                                                                                            \
            return 36;                                                                               \
        }                                                                                            \
        catch (const std::exception&)                                                                \
        {                                                                                            \
            return 35;                                                                               \
        }                                                                                            \
    }
ORACLE_AGG(uint8_t, u8)
ORACLE_AGG(uint16_t, u16)
ORACLE_AGG(uint32_t, u32)
ORACLE_AGG(uint64_t, u64)
ORACLE_AGG(int8_t, i8)
ORACLE_AGG(int16_t, i16)
ORACLE_AGG(int32_t, i32)
ORACLE_AGG(int64_t, i64)
ORACLE_AGG(float, f32)
ORACLE_AGG(double, f64)

// ------------------------------------------------------- library: spatial grid
void* oracle_grid_new(int dim, int has_bounds, const int32_t* bounds)
{
    SpatialGrid* g = new SpatialGrid;
    g->dim = dim;
    g->bounds.some = has_bounds != 0;
    if (has_bounds)
    {
        g->bounds.min = Pos{bounds[0], bounds[1], dim == 3 ? bounds[2] : 0};
        g->bounds.max = Pos{bounds[dim], bounds[dim + 1], dim == 3 ? bounds[dim + 2] : 0};
    }
    return g;
}
void oracle_grid_free(void* g) { delete static_cast<SpatialGrid*>(g); }
int oracle_grid_insert(void* g, uint32_t e, const int32_t* p)
{
    SpatialGrid* G = static_cast<SpatialGrid*>(g);
    try
    {
        G->insert_or_update(e, Pos{p[0], p[1], G->dim == 3 ? p[2] : 0});
    }
    catch (const std::out_of_range&)
    {
        return 34;
    }
    return 0;
}
void oracle_grid_remove(void* g, uint32_t e) { static_cast<SpatialGrid*>(g)->remove(e); }
uint64_t oracle_grid_query(void* g, int mode /*0 at, 1 neighbors, 2 orthogonal*/, const int32_t* p, uint32_t* out,
                           uint64_t cap)
{
    SpatialGrid* G = static_cast<SpatialGrid*>(g);
    const Pos q{p[0], p[1], G->dim == 3 ? p[2] : 0};
    uint64_t n = 0;
    auto f = [&](Entity e) {
        if (n < cap) out[n] = e;
        ++n;
    };
    if (mode == 0)
        G->entities_at(q, f);
    else if (mode == 1)
        G->neighbors_of(q, f);
    else
        G->orthogonal_neighbors_of(q, f);
    return n;
}
uint64_t oracle_grid_num_entities(void* g) { return static_cast<SpatialGrid*>(g)->num_entities(); }
int oracle_grid_in_bounds(void* g, const int32_t* p)
{
    SpatialGrid* G = static_cast<SpatialGrid*>(g);
    return G->in_bounds(Pos{p[0], p[1], G->dim == 3 ? p[2] : 0}) ? 1 : 0;
}
uint64_t oracle_neighbor_offsets(int dim, int orthogonal, int32_t* out /* up to 26*3 */)
{
    const Pos o{0, 0, 0};
    const std::vector<Pos> v = orthogonal ? neighbors_orthogonal(o, dim) : neighbors(o, dim);
    for (size_t i = 0; i < v.size(); ++i)
    {
        out[3 * i] = v[i].x;
        out[3 * i + 1] = v[i].y;
        out[3 * i + 2] = v[i].z;
    }
    return v.size();
}

} // extern "C"
