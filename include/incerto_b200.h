/*
 * incerto_b200.h — C ABI of the B200-native execution engine for incerto's
 * per-step simulation hot path.
 *
 * The reference (haath/incerto v0.5.0) has no FFI of its own: its boundary is
 * the Rust method surface of `SimulationBuilder` (src/simulation_builder.rs:39-306)
 * and `Simulation` (src/simulation.rs:22-259).  Every entry point below names
 * the reference method it replaces.  A Rust `incerto-cuda-sys` crate (see
 * INTEGRATION.md) binds exactly these symbols; in this repository the same
 * symbols are driven by the C++ facade `include/incerto_b200.hpp` and by the
 * ctypes mirror `incerto_b200/simulation.py`.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types cross the boundary;
 *  - every call returns an `int32_t` status (`incerto_status`); nothing unwinds;
 *  - input buffers are borrowed for the duration of the call and copied;
 *  - pointers handed out by `incerto_sim_get_*time_series` are borrowed from the
 *    handle and stay valid until the next `run`/`reset`/`destroy` on it
 *    (mirrors `TimeSeries<'a>` borrowing `&self`, src/simulation.rs:186-216);
 *  - a handle is not thread-safe: one handle <-> one host thread <-> one device
 *    (`Simulation::run` takes `&mut self`, src/simulation.rs:25);
 *  - there is no CPU fallback: without a CUDA device `build` fails with
 *    INCERTO_ERR_NO_DEVICE.
 */
#ifndef INCERTO_B200_H
#define INCERTO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define INCERTO_B200_ABI_VERSION 1

typedef struct incerto_builder incerto_builder; /* SimulationBuilder, src/simulation_builder.rs:25-29 */
typedef struct incerto_sim incerto_sim;         /* Simulation,        src/simulation.rs:17-20 */

/* ------------------------------------------------------------------ status */
typedef enum incerto_status
{
    INCERTO_OK = 0,
    /* SamplingError, src/error.rs:12-43 (same order) */
    INCERTO_ERR_COMPONENT_DOES_NOT_EXIST = 1,
    INCERTO_ERR_SINGLE_NO_ENTITIES = 2,
    INCERTO_ERR_SINGLE_MULTIPLE_ENTITIES = 3,
    INCERTO_ERR_AGGREGATE_NO_ENTITIES = 4,
    INCERTO_ERR_TIME_SERIES_NOT_RECORDED = 5,
    INCERTO_ERR_ENTITY_IDENTIFIER_NOT_FOUND = 6,
    INCERTO_ERR_ENTITY_IDENTIFIER_NOT_UNIQUE = 7,
    /* BuilderError, src/error.rs:47-52 */
    INCERTO_ERR_TIME_SERIES_RECORDING_CONFLICT = 16,
    /* engine codes; the reference panics or cannot express these */
    INCERTO_ERR_INVALID_ARGUMENT = 32, /* e.g. sample_interval == 0 (reference asserts, simulation_builder.rs:194,271) */
    INCERTO_ERR_CUDA = 33,
    INCERTO_ERR_OUT_OF_BOUNDS = 34,    /* entity outside spatial grid bounds (reference panics, spatial_grid.rs:276-279) */
    INCERTO_ERR_NOT_COMPARABLE = 35,   /* NaN in Minimum/Maximum/Median/Percentile (reference panics, aggregate.rs:221-232) */
    INCERTO_ERR_PERCENTILE_RANGE = 36, /* Percentile index == len, i.e. P == 100 (reference indexes out of range, aggregate.rs:131-132) */
    INCERTO_ERR_UNSUPPORTED = 37,
    INCERTO_ERR_NO_DEVICE = 38,
    INCERTO_ERR_COMM = 39
} incerto_status;

/* Human-readable detail of the last non-OK status on this thread. */
const char* incerto_last_error_message(void);
uint32_t incerto_abi_version(void);

/* --------------------------------------------------------------- components */
/* A reference `Component` (Rust struct) maps to one or more engine components:
 * one per numeric field ("column"), or a marker for field-less structs.
 * Storage dtype is the device layout. */
typedef enum incerto_dtype
{
    INCERTO_MARKER = 0, /* field-less component (Person, Infected, GroupA ...): one bit per row */
    INCERTO_U8 = 1,
    INCERTO_U16 = 2,
    INCERTO_U32 = 3,
    INCERTO_U64 = 4, /* Rust usize */
    INCERTO_I8 = 5,
    INCERTO_I16 = 6,
    INCERTO_I32 = 7,
    INCERTO_I64 = 8,
    INCERTO_F32 = 9,
    INCERTO_F64 = 10
} incerto_dtype;

typedef int32_t incerto_component; /* >= 0; negative values are never valid */

/* Register a component type.  `lanes` is 1 for scalars, 2/3 for IVec2/IVec3-like
 * components (GridPosition<T>, src/plugins/spatial_grid.rs:117-118); multi-lane
 * data crosses the ABI interleaved (x0,y0,x1,y1,...) exactly as a Rust
 * `Vec<IVec2>` lies in memory.  Returns the handle in *out. */
int32_t incerto_builder_register_component(incerto_builder* b, const char* name, incerto_dtype dtype,
                                           uint32_t lanes, incerto_component* out);

/* ------------------------------------------------------------------ builder */
/* SimulationBuilder::new, src/simulation_builder.rs:42-56.  The step counter
 * starts so that the first user step observes StepNumber == 1 (:50). */
int32_t incerto_builder_new(incerto_builder** out);
void incerto_builder_free(incerto_builder* b); /* drop without build */

/* Engine knobs with no reference counterpart (the reference seeds from the OS). */
int32_t incerto_builder_set_seed(incerto_builder* b, uint64_t seed, uint32_t replica);
int32_t incerto_builder_set_device(incerto_builder* b, int32_t cuda_device_ordinal);

/* Spatial decomposition of ONE world over `world` processes (one per GPU): rank r owns grid x in
 * [extent * r / world, extent * (r + 1) / world).
 *  - forest_fire: the dense grid is split by columns of x; each rank holds its slab + one halo column per side, the
 *    halo exchange is fused into the step kernel (peer memory) or runs over NCCL.
 *  - air_traffic_3d: the airspace is split by x.  Every rank spawns the whole population and keeps the full table;
 *    a row is alive on the rank that owns the aircraft's position.  After every step the ranks exchange one fixed-size
 *    mailbox per neighbour: the aircraft that crossed the cut (they come alive on the receiver) and the positions of the
 *    boundary-column aircraft (ghosts for check_conflicts).  Needs a communicator (incerto_sim_attach_comm) or the
 *    manual transport below with one step per run.  The local AIR_CONFLICTS aggregate of one rank is its number of
 *    conflict ENDS (a pair across a cut has one end on either side); incerto_sim_sample_aggregate_global and the
 *    aggregate time series (collective) report the world's conflicts. */
int32_t incerto_builder_set_row_partition(incerto_builder* b, uint32_t rank, uint32_t world);

/* One column of a bulk spawn. `data` holds n*lanes values of the component's
 * dtype; for markers it is NULL (every row carries the marker) or n bytes of
 * 0/1 (per-row presence). */
typedef struct incerto_column_data
{
    incerto_component component;
    const void* data;
} incerto_column_data;

/* SimulationBuilder::add_entity_spawner (:154-158) + Spawner::spawn
 * (src/spawner.rs:8-11), in bulk: spawns `n` entities carrying the listed
 * components.  Spawners run once, in insertion order, inside `build`
 * (:295-305); entity uid = spawn ordinal inside its archetype table. */
int32_t incerto_builder_add_entity_spawner(incerto_builder* b, uint64_t n, const incerto_column_data* columns,
                                           uint32_t n_columns);

/* The same without the copy: the column buffers must stay valid and unchanged until incerto_builder_build
 * returns (a spawner closure's staging buffers do).  A simulation built this way cannot be `reset`. */
int32_t incerto_builder_add_entity_spawner_borrowed(incerto_builder* b, uint64_t n, const incerto_column_data* columns,
                                                    uint32_t n_columns);

/* Built-in device-side spawners: the examples' spawn functions restated with
 * the Philox contract so that populations too large to upload are created in
 * HBM directly. */
typedef enum incerto_spawner_kind
{
    INCERTO_SPAWN_COIN_TOSSERS = 1,    /* examples/coin_toss.rs:58-63 */
    INCERTO_SPAWN_FOREST_GRID = 2,     /* examples/forest_fire.rs:234-279 */
    INCERTO_SPAWN_PANDEMIC_PEOPLE = 3, /* examples/pandemic.rs:84-107 */
    INCERTO_SPAWN_SPATIAL_POPULATION = 4, /* examples/pandemic_spatial.rs:279-313 */
    INCERTO_SPAWN_TRADERS = 5,         /* examples/traders.rs:116-127 */
    INCERTO_SPAWN_STOCKS = 6,          /* examples/traders.rs:130-137 */
    INCERTO_SPAWN_AIRCRAFT = 7         /* examples/air_traffic_3d.rs:61-90 */
} incerto_spawner_kind;

int32_t incerto_builder_add_device_spawner(incerto_builder* b, uint32_t kind, const void* params,
                                           uint32_t params_size);

/* ------------------------------------------------------------------ systems */
/* SimulationBuilder::add_systems (:66-70).  A reference system is arbitrary
 * Rust; here each system is a hand-written sm_100a kernel selected by kind and
 * bound to components through its params struct.  Within a step systems run
 * in ascending `kind` order — the pinned linearisation of DESIGN.md §4 (the
 * reference leaves the order unspecified, simulation_builder.rs:62-64). */
typedef enum incerto_system_kind
{
    /* generic, used by the ports of tests/test_counter.rs */
    INCERTO_SYS_ADD_CONST = 1,  /* counter.0 += c                tests/test_counter.rs:41-45,71-76,109-121 */
    INCERTO_SYS_ADD_COLUMN = 2, /* counter.0 += id.0             tests/test_counter.rs:202-207 */
    /* examples/coin_toss.rs:67-81 */
    INCERTO_SYS_COIN_TOSS = 10,
    /* examples/forest_fire.rs: spread :282-329, progression :332-351, regrowth :354-365 */
    INCERTO_SYS_FOREST_SPREAD = 20,
    INCERTO_SYS_FOREST_BURN_PROGRESSION = 21,
    INCERTO_SYS_FOREST_REGROWTH = 22,
    /* examples/pandemic.rs: move :115-166, infect :169-192, die :195-206, recover :209-223 */
    INCERTO_SYS_PANDEMIC_MOVE = 30,
    INCERTO_SYS_PANDEMIC_INFECT = 31,
    INCERTO_SYS_PANDEMIC_DIE = 32,
    INCERTO_SYS_PANDEMIC_RECOVER = 33,
    /* examples/pandemic_spatial.rs (canonical order) */
    INCERTO_SYS_SPATIAL_INCUBATION = 40,      /* :414-433 */
    INCERTO_SYS_SPATIAL_TRANSMISSION = 41,    /* :436-518 */
    INCERTO_SYS_SPATIAL_RECOVERY_DEATH = 42,  /* :521-541 */
    INCERTO_SYS_SPATIAL_CONTACT_HISTORY = 43, /* :544-562 */
    INCERTO_SYS_SPATIAL_CONTACT_TRACING = 44, /* :565-608 */
    INCERTO_SYS_SPATIAL_QUARANTINE = 45,      /* :611-625 */
    INCERTO_SYS_SPATIAL_MOVE = 46,            /* :316-411 */
    /* examples/traders.rs */
    INCERTO_SYS_TRADERS_BUY = 50,          /* :165-197 */
    INCERTO_SYS_TRADERS_SELL = 51,         /* :199-236 */
    INCERTO_SYS_TRADERS_NET_WORTH = 52,    /* :239-263 */
    INCERTO_SYS_STOCKS_PRICE_CHANGE = 53,  /* :140-163 */
    /* examples/air_traffic_3d.rs */
    INCERTO_SYS_AIR_CHECK_CONFLICTS = 60, /* :124-162 */
    INCERTO_SYS_AIR_MOVE = 61             /* :93-121 */
} incerto_system_kind;

typedef struct incerto_system_desc
{
    uint32_t kind;
    const void* params; /* the incerto_*_params struct of that kind */
    uint32_t params_size;
} incerto_system_desc;

int32_t incerto_builder_add_systems(incerto_builder* b, const incerto_system_desc* systems, uint32_t n);

/* Query filter: With<..>/Without<..> lists (bevy QueryFilter as used by the
 * reference: simulation.rs:136,163; tests/test_counter.rs:115,134). */
typedef struct incerto_filter
{
    const incerto_component* with;
    uint32_t n_with;
    const incerto_component* without;
    uint32_t n_without;
} incerto_filter;

/* --- params: generic */
typedef struct incerto_add_const_params
{
    incerto_component target; /* integer column */
    int64_t amount;
    incerto_component with[4]; /* With<..> filter, n_with entries used */
    uint32_t n_with;
} incerto_add_const_params;

typedef struct incerto_add_column_params
{
    incerto_component target;
    incerto_component source;
} incerto_add_column_params;

/* --- params: coin_toss (examples/coin_toss.rs:18-28,67-81) */
typedef struct incerto_coin_toss_params
{
    incerto_component num_heads;  /* U32 or U64 column */
    incerto_component num_tosses; /* same dtype */
    double odds_heads;            /* ODDS_HEADS :20 */
} incerto_coin_toss_params;

typedef struct incerto_spawn_coin_tossers_params
{
    incerto_component num_heads, num_tosses;
    uint64_t num_entities; /* NUM_ENTITIES :19 */
} incerto_spawn_coin_tossers_params;

/* --- params: forest_fire (examples/forest_fire.rs:32-44) */
/* ForestCell.state device encoding (u8): Burned = 0, Burning{t} = t (1..=255),
 * Healthy = 0x40 ... see DESIGN.md §3.2; the facade translates CellState. */
#define INCERTO_CELL_BURNED 0u
#define INCERTO_CELL_BURNING(t) ((uint8_t)(t)) /* 1 <= t <= 63 */
#define INCERTO_CELL_HEALTHY 0x40u
#define INCERTO_CELL_EMPTY 0x80u

typedef struct incerto_forest_params
{
    incerto_component position; /* GridPosition2D: 2 lanes of I16 (or U16 / I32 for a host-spawned grid) */
    incerto_component cell;     /* ForestCell.state, U8 */
    double probability;         /* FIRE_SPREAD_PROBABILITY :39 (spread) / REGROWTH_PROBABILITY :41 (regrowth); unused for progression */
    uint32_t burn_duration;     /* BURN_DURATION :40 (spread) */
} incerto_forest_params;

typedef struct incerto_spawn_forest_params
{
    incerto_component position, cell;
    int32_t width, height;          /* GRID_WIDTH/GRID_HEIGHT :33-34 */
    double initial_forest_density;  /* :38 */
    uint32_t burn_duration;         /* :40 */
    uint32_t initial_fire_count;    /* :42 */
} incerto_spawn_forest_params;

/* --- params: pandemic (examples/pandemic.rs:18-33) */
typedef struct incerto_pandemic_params
{
    incerto_component person;   /* marker */
    incerto_component position; /* Position{x,y}: U16 x 2 lanes */
    incerto_component infected; /* marker */
    uint32_t grid_size;         /* GRID_SIZE :20 */
    double probability;         /* CHANCE_INFECT_OTHER / CHANCE_DIE / CHANCE_RECOVER; unused for move */
} incerto_pandemic_params;

typedef struct incerto_spawn_pandemic_params
{
    incerto_component person, position, infected;
    uint64_t initial_population; /* :19 */
    uint32_t grid_size;          /* :20 */
    double chance_start_infected; /* :22 */
} incerto_spawn_pandemic_params;

/* --- params: air_traffic_3d (examples/air_traffic_3d.rs:14-19) */
typedef struct incerto_air_params
{
    incerto_component position;        /* GridPosition3D, I32 x 3 lanes at the ABI */
    incerto_component target_altitude; /* Aircraft.target_altitude, an I16 column (device layout; the façades narrow the reference's usize) */
    int32_t airspace_size;             /* AIRSPACE_SIZE :16 */
    int32_t airspace_height;           /* AIRSPACE_HEIGHT :17 */
    double chance_horizontal_move;     /* 0.7 :108 */
} incerto_air_params;

typedef struct incerto_spawn_aircraft_params
{
    incerto_component position, target_altitude;
    uint64_t num_aircraft; /* :18 */
    int32_t airspace_size, airspace_height;
} incerto_spawn_aircraft_params;

/* --- params: pandemic_spatial (examples/pandemic_spatial.rs:28-58) */
/* Person.disease_state device encoding (u8): Healthy = 0, Exposed{n} = n (1..=127),
 * Infectious = 0x80, Recovered = 0x81. */
#define INCERTO_DISEASE_HEALTHY 0u
#define INCERTO_DISEASE_EXPOSED(n) ((uint8_t)(n))
#define INCERTO_DISEASE_INFECTIOUS 0x80u
#define INCERTO_DISEASE_RECOVERED 0x81u

typedef struct incerto_spatial_params
{
    incerto_component position;          /* GridPosition2D */
    incerto_component disease_state;     /* Person.disease_state, U8 */
    incerto_component social_distancing; /* Person.social_distancing, marker */
    incerto_component quarantined;       /* Quarantined.remaining_duration, U8; 0 = component absent */
    incerto_component contact_history;   /* ContactHistory, engine-managed fixed-capacity set */
    int32_t grid_size;                   /* :30 */
    int32_t infection_radius;            /* :34 (2) */
    double chance_infect_distance_1;     /* :35 */
    double chance_infect_distance_2;     /* :36 */
    double chance_recover;               /* :37 */
    double chance_die;                   /* :38 */
    uint32_t incubation_period;          /* :39 */
    uint32_t crowding_threshold;         /* :43 */
    uint32_t contact_quarantine_duration; /* :48 */
    int32_t quarantine_center_x, quarantine_center_y, quarantine_radius; /* :52-55 */
    uint32_t contact_history_clear_above; /* 20, :555 */
    double chance_attempt_move;          /* 0.5 :332 */
} incerto_spatial_params;

typedef struct incerto_spawn_spatial_params
{
    incerto_component position, disease_state, social_distancing, quarantined, contact_history;
    uint64_t initial_population;       /* :29 */
    int32_t grid_size;                 /* :30 */
    double social_distance_compliance; /* :44 */
    double chance_start_infected;      /* :33 */
    uint32_t incubation_period;        /* :39 */
} incerto_spawn_spatial_params;

/* --- params: traders (examples/traders.rs:35-47) */
typedef struct incerto_traders_params
{
    incerto_component trader_id;    /* TraderId, U64 */
    incerto_component cash;         /* Trader.cash, F64 */
    incerto_component portfolio;    /* Trader.portfolio: engine-managed, num_stocks x 4-bit share counts */
    incerto_component net_worth;    /* TraderNetWorth, F64 */
    incerto_component stock_id;     /* StockId, U64 */
    incerto_component stock_price;  /* StockPrice, F64 */
    incerto_component stock_delisted; /* marker */
    double chance_buy;              /* :40 */
    double chance_sell;             /* :41 */
    uint32_t buy_shares_min, buy_shares_max; /* :42 */
    double price_change_mean, price_change_std_dev; /* :45-46 */
} incerto_traders_params;

typedef struct incerto_spawn_traders_params
{
    incerto_component trader_id, cash, portfolio, net_worth;
    uint64_t num_traders; /* :36 */
    double initial_cash;  /* :39 */
    uint32_t num_stocks;  /* portfolio capacity, :37 */
} incerto_spawn_traders_params;

typedef struct incerto_spawn_stocks_params
{
    incerto_component stock_id, stock_price, stock_delisted;
    uint32_t num_stocks;   /* :37 */
    double initial_price;  /* :44 */
} incerto_spawn_stocks_params;

/* ----------------------------------------------------------- spatial grid */
/* SimulationBuilder::add_spatial_grid{,_2d,_3d} (:114-139): index of entities
 * having `position` (a 2- or 3-lane component) and `component`.  `bounds` is
 * {min x,y[,z], max x,y[,z]} inclusive (GridBounds, spatial_grid.rs:45-53) or
 * NULL.  On the device the HashMap of the reference becomes a counting sort by
 * cell key rebuilt at the start of every step (the PreUpdate maintenance of
 * spatial_grid.rs:410-455 observed at step granularity).  Unbounded grids
 * (bounds == NULL, simulation_builder.rs:114-121) accept every position; their index is built on demand, when the
 * host queries it, over the bounding box of the entities the last PreUpdate saw (at most 2^28 cells).  The example
 * systems that iterate a grid on the device (forest_fire, air_traffic_3d) need bounds. */
typedef int32_t incerto_grid;
int32_t incerto_builder_add_spatial_grid(incerto_builder* b, uint32_t dim, const int32_t* bounds,
                                         incerto_component position, incerto_component component,
                                         incerto_grid* out);

/* -------------------------------------------------------------- aggregates */
/* `Out` of SampleAggregate<Out> (src/traits.rs:11-20). */
typedef enum incerto_aggregate_kind
{
    INCERTO_AGG_SUM = 1,        /* tests' `components.iter().map(|c| c.0).sum()`; out: component dtype widened to 64 bit */
    INCERTO_AGG_COUNT = 2,      /* `components.len()` (air_traffic_3d.rs:53-59); out u64 */
    INCERTO_AGG_MINIMUM = 3,    /* Minimum<T>  src/util/aggregate.rs:73-88; out T */
    INCERTO_AGG_MAXIMUM = 4,    /* Maximum<T>  :90-105 */
    INCERTO_AGG_MEAN = 5,       /* Mean<T>     :157-192 (T arithmetic: integer division for ints) */
    INCERTO_AGG_MEDIAN = 6,     /* Median<T>   :137-155 = sorted[len/2] */
    INCERTO_AGG_PERCENTILE = 7, /* Percentile<T,P> :107-135 = sorted[floor(P/100*len)] */
    INCERTO_AGG_COIN_ODDS = 16,      /* examples/coin_toss.rs:33-49; out f64 */
    INCERTO_AGG_FIRE_STATS = 17,     /* examples/forest_fire.rs:104-134; out incerto_fire_stats */
    INCERTO_AGG_PANDEMIC_STATS = 18, /* examples/pandemic_spatial.rs:107-138; out incerto_pandemic_stats */
    INCERTO_AGG_AIR_CONFLICTS = 19   /* conflicts/2 of the last step (air_traffic_3d.rs:130-161); out u64 */
} incerto_aggregate_kind;

typedef struct incerto_fire_stats
{
    uint64_t healthy_count, burning_count, burned_count, empty_count, total_cells;
} incerto_fire_stats;

typedef struct incerto_pandemic_stats
{
    uint64_t healthy_count, exposed_count, infectious_count, recovered_count, dead_count, total_population;
} incerto_pandemic_stats;

typedef struct incerto_aggregate_desc
{
    incerto_component component;  /* C */
    incerto_component component2; /* second column where the aggregate needs one (COIN_ODDS: num_tosses); else -1 */
    uint32_t kind;                /* incerto_aggregate_kind */
    uint32_t percentile;          /* P for PERCENTILE (0..=100) */
    uint64_t param;               /* PANDEMIC_STATS: INITIAL_POPULATION (pandemic_spatial.rs:134) */
    incerto_filter filter;        /* F; empty = () */
} incerto_aggregate_desc;

/* SimulationBuilder::record_aggregate_time_series{,_filtered} (:230-283).
 * Conflict (same component, filter, out kind) -> TIME_SERIES_RECORDING_CONFLICT;
 * interval == 0 -> INVALID_ARGUMENT. */
int32_t incerto_builder_record_aggregate_time_series(incerto_builder* b, const incerto_aggregate_desc* agg,
                                                     uint64_t sample_interval);

/* SimulationBuilder::record_time_series::<C, I, O> (:185-206): per-entity series
 * of `component`, entities identified by `identifier`. */
int32_t incerto_builder_record_time_series(incerto_builder* b, incerto_component component,
                                           incerto_component identifier, uint64_t sample_interval);

/* SimulationBuilder::build (:295-305). Consumes the builder (also on failure). */
int32_t incerto_builder_build(incerto_builder* b, incerto_sim** out);

/* --------------------------------------------------------------- simulation */
/* Simulation::run (src/simulation.rs:25-31). Returns after the device finished. */
int32_t incerto_sim_run(incerto_sim* s, uint64_t num_steps);
/* Same, but only enqueues; pair with incerto_sim_synchronize. */
int32_t incerto_sim_run_async(incerto_sim* s, uint64_t num_steps);
int32_t incerto_sim_synchronize(incerto_sim* s);
void incerto_sim_destroy(incerto_sim* s);

/* Re-run the spawners under another replica key, reusing all device memory
 * ("restart the simulation from the beginning", simulation_builder.rs:22-24). */
int32_t incerto_sim_reset(incerto_sim* s, uint32_t replica);

/* Simulation::count::<F> (:163-173). */
int32_t incerto_sim_count(incerto_sim* s, const incerto_filter* filter, uint64_t* out);
/* Simulation::sample_aggregate / sample_aggregate_filtered (:107-152). `out`
 * receives the kind's record (see incerto_aggregate_kind). */
int32_t incerto_sim_sample_aggregate(incerto_sim* s, const incerto_aggregate_desc* agg, void* out, size_t out_size);
/* The same aggregate over the entities of ALL ranks of the attached communicator (cross-replica statistics: every rank
 * holds one replica, or one shard).  Collective: every rank calls it with the same aggregate.  Count / Sum / Minimum /
 * Maximum / Mean combine the per-rank records over NCCL; Median / Percentile run as a distributed radix select (global
 * count, the 256-bin histogram of every pass summed over the ranks), so the answer is the one a single process
 * holding all the entities would give (src/util/aggregate.rs:107-155).  Also COIN_ODDS and FIRE_STATS. */
int32_t incerto_sim_sample_aggregate_global(incerto_sim* s, const incerto_aggregate_desc* agg, void* out, size_t out_size);
/* Simulation::sample::<C, Id, Out> (:43-65). `id_value`: one value of the
 * identifier's dtype; `out`: `lanes` values of the component's dtype. */
int32_t incerto_sim_sample(incerto_sim* s, incerto_component component, incerto_component identifier,
                           const void* id_value, void* out, size_t out_size);
/* Simulation::sample_single::<C, Out> (:79-93). */
int32_t incerto_sim_sample_single(incerto_sim* s, incerto_component component, void* out, size_t out_size);

/* Simulation::get_aggregate_time_series{,_filtered} (:226-258).  *values is an
 * array of *len records of *record_size bytes, *time the step stamps.
 * COLLECTIVE for the FireStats series of a row-partitioned forest with a communicator attached: the per-rank partial
 * records are all-reduced when the series is first read after a run (then cached until the next run), so every rank
 * of the communicator must read it - the same number of times between runs - or none may. */
int32_t incerto_sim_get_aggregate_time_series(incerto_sim* s, const incerto_aggregate_desc* agg,
                                              const uint64_t** time, const void** values, uint64_t* len,
                                              uint64_t* record_size, uint64_t* sample_interval);
/* Simulation::get_time_series::<C, Id, Out> (:186-216). */
int32_t incerto_sim_get_time_series(incerto_sim* s, incerto_component component, incerto_component identifier,
                                    const void* id_value, const uint64_t** time, const void** values,
                                    uint64_t* len, uint64_t* record_size, uint64_t* sample_interval);

/* ------------------------------------------------- SpatialGrid<T, C> queries */
/* Host-side views of the device index, mirroring src/plugins/spatial_grid.rs:
 * entities_at :314-320, neighbors_of :331-350, orthogonal_neighbors_of :355-372,
 * is_empty :376-381, num_entities :385-388, position_of :324-327, in_bounds :265-271.
 * Entities are reported by uid. `pos` has `dim` lanes. */
int32_t incerto_sim_grid_entities_at(incerto_sim* s, incerto_grid g, const int32_t* pos, uint64_t* out_uids,
                                     uint64_t capacity, uint64_t* n_out);
int32_t incerto_sim_grid_neighbors_of(incerto_sim* s, incerto_grid g, const int32_t* pos, uint32_t orthogonal,
                                      uint64_t* out_uids, uint64_t capacity, uint64_t* n_out);
int32_t incerto_sim_grid_is_empty(incerto_sim* s, incerto_grid g, const int32_t* pos, uint32_t* out);
int32_t incerto_sim_grid_num_entities(incerto_sim* s, incerto_grid g, uint64_t* out);
int32_t incerto_sim_grid_position_of(incerto_sim* s, incerto_grid g, uint64_t uid, int32_t* out_pos);
int32_t incerto_sim_grid_in_bounds(incerto_sim* s, incerto_grid g, const int32_t* pos, uint32_t* out);

/* ------------------------------------------------------ engine introspection */
int32_t incerto_sim_step_number(incerto_sim* s, uint64_t* out); /* StepNumber, src/plugins/step_number.rs:6 */
/* Number of live rows carrying `component` and their column contents in row
 * order (dead rows skipped), with their uids.  Used by the parity tests. */
int32_t incerto_sim_num_entities(incerto_sim* s, incerto_component component, uint64_t* out);
int32_t incerto_sim_read_component(incerto_sim* s, incerto_component component, void* out, size_t out_size,
                                   uint64_t* uids_out /* may be NULL */);
/* Presence (0/1) of `marker` on every live row of the tables holding the data component `rows_of`,
 * in the row order of incerto_sim_read_component(rows_of). */
int32_t incerto_sim_read_marker(incerto_sim* s, incerto_component marker, incerto_component rows_of, uint8_t* out,
                                size_t out_size);
/* Device time (ms, CUDA events on the engine's stream) and kernel launches of
 * the last incerto_sim_run / run_async+synchronize. */
int32_t incerto_sim_last_run_stats(incerto_sim* s, double* device_ms, uint64_t* kernel_launches);
/* Independent replicas can run concurrently on one GPU: one handle (one stream) each, incerto_sim_run_async on all of
 * them, then incerto_sim_synchronize.  This returns the device time from the start of `first`'s last run to the end of
 * `last`'s last run (same device, both synchronised). */
int32_t incerto_sim_span_ms(incerto_sim* first, incerto_sim* last, double* ms_out);
/* Per-kernel accumulated device time since build: writes up to `capacity`
 * entries; names are static strings. */
typedef struct incerto_kernel_stat
{
    const char* name;
    uint64_t launches;
    double device_ms; /* only accumulated while profiling is enabled */
    double algorithmic_bytes;
} incerto_kernel_stat;
int32_t incerto_sim_set_profiling(incerto_sim* s, uint32_t enabled);
int32_t incerto_sim_kernel_stats(incerto_sim* s, incerto_kernel_stat* out, uint32_t capacity, uint32_t* n_out);
/* Write >= 2x L2 worth of HBM so the next kernel starts cold (bench hygiene). */
int32_t incerto_sim_flush_l2(incerto_sim* s);
/* Bench hygiene in one call: `num_steps` steps of Simulation::run (src/simulation.rs:25-31), each started from a
 * flushed L2 and all enqueued back to back (no host round trip between the steps, so the ranks of a row-partitioned
 * grid stay in lock step on the device).  incerto_sim_last_run_stats then reports the SUM over the steps of the
 * CUDA-event time around each step alone (the flushes are not timed; with a communicator attached an untimed
 * all-reduce lines the ranks up again after each flush). */
int32_t incerto_sim_run_cold(incerto_sim* s, uint64_t num_steps);

/* ----------------------------------------------------------------- multi-GPU */
/* Attach an NCCL communicator (one process per GPU).  `unique_id` is the
 * 128-byte ncclUniqueId created on rank 0 (incerto_comm_unique_id) and
 * distributed by the host (torch.distributed in this repo).  Used for the
 * forest-fire halo exchange and for cross-replica aggregate reductions.
 * Communicators are cached per (unique id, device) for the life of the process: handles that attach with the same id
 * share one ncclComm (created by the first attach, ~0.5 s), so building simulation after simulation does not pay for
 * communicator setup again.  The handles sharing it are driven from one thread, one collective at a time. */
int32_t incerto_comm_unique_id(uint8_t out[128]);
int32_t incerto_sim_attach_comm(incerto_sim* s, const uint8_t unique_id[128], uint32_t rank, uint32_t world);
/* Manual halo transport for the row-partitioned forest (any host transport instead of NCCL; also how
 * the partition is tested on a single GPU).  When no communicator is attached, `run` leaves the halo
 * columns alone and the host moves them between steps: export the boundary of side 0 (lowest owned
 * column) / 1 (highest) — H cell bytes followed by the overlay-entity states — and import it into the
 * neighbour's halo of the opposite side.
 * For an x-partitioned air_traffic_3d world the same two calls move the per-step mailboxes: export = this rank's outbox
 * for the neighbour on `side` after a step (call with a small `capacity` to learn the size through *n_out), import =
 * the neighbour's outbox into this rank's inbox of that side; the import revives the aircraft that migrated in.  One
 * step per `run` call, then both directions of every cut, before the next step. */
int32_t incerto_sim_halo_export(incerto_sim* s, uint32_t side, void* out, size_t capacity, size_t* n_out);
int32_t incerto_sim_halo_import(incerto_sim* s, uint32_t side, const void* in, size_t n);
/* Columns [begin, end) of a W-column grid owned by `rank` of `world` (pure host arithmetic). */
int32_t incerto_partition_range(int32_t width, uint32_t rank, uint32_t world, int32_t* begin, int32_t* end);
/* All-reduce `n` u64 (sum) / f64 (sum, min, max) values in place across the
 * attached communicator: cross-replica aggregate statistics. op: 0 sum, 1 min, 2 max. */
int32_t incerto_sim_allreduce_u64(incerto_sim* s, uint64_t* values, uint32_t n, uint32_t op);
int32_t incerto_sim_allreduce_f64(incerto_sim* s, double* values, uint32_t n, uint32_t op);

#ifdef __cplusplus
}
#endif
#endif /* INCERTO_B200_H */
