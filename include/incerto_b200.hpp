// incerto_b200.hpp — header-only C++ host façade over the C ABI (include/incerto_b200.h).
//
// Mirrors the reference's SimulationBuilder / Simulation method names, argument meaning and error
// behaviour (src/simulation_builder.rs:39-306, src/simulation.rs:22-259, src/error.rs).  The reference
// is compiled Rust whose toolchain is absent here, so this is the compiled-language host side.
#pragma once
#include <cstdint>
#include <functional>
#include <initializer_list>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "incerto_b200.h"

namespace incerto
{

// src/error.rs:12-43
enum class SamplingError
{
    ComponentDoesNotExist = 1,
    SingleNoEntities,
    SingleMultipleEntities,
    AggregateNoEntities,
    TimeSeriesNotRecorded,
    EntityIdentifierNotFound,
    EntityIdentifierNotUnique
};
// src/error.rs:47-52
enum class BuilderError
{
    TimeSeriesRecordingConflict = 16
};

struct SimulationError : std::runtime_error
{
    int32_t code;
    explicit SimulationError(int32_t c) : std::runtime_error(describe(c)), code(c) {}
    bool is(SamplingError e) const { return code == static_cast<int32_t>(e); }
    bool is(BuilderError e) const { return code == static_cast<int32_t>(e); }
    static std::string describe(int32_t c)
    {
        static const char* names[] = {"Ok", "ComponentDoesNotExist", "SingleNoEntities", "SingleMultipleEntities",
                                      "AggregateNoEntities", "TimeSeriesNotRecorded", "EntityIdentifierNotFound",
                                      "EntityIdentifierNotUnique"};
        if (c >= 1 && c <= 7) return names[c];
        if (c == 16) return "TimeSeriesRecordingConflict";
        return "incerto_b200 status " + std::to_string(c) + ": " + incerto_last_error_message();
    }
};
inline void check(int32_t status)
{
    if (status != INCERTO_OK) throw SimulationError(status);
}

struct Component
{
    incerto_component handle = -1;
    incerto_dtype dtype = INCERTO_MARKER;
    uint32_t lanes = 1;
};
struct With
{
    Component c;
};
struct Without
{
    Component c;
};

class Filter
{
  public:
    Filter() = default;
    Filter(std::initializer_list<With> w, std::initializer_list<Without> wo = {})
    {
        for (auto& x : w) with_.push_back(x.c.handle);
        for (auto& x : wo) without_.push_back(x.c.handle);
    }
    incerto_filter raw() const
    {
        return incerto_filter{with_.data(), static_cast<uint32_t>(with_.size()), without_.data(),
                              static_cast<uint32_t>(without_.size())};
    }

  private:
    std::vector<incerto_component> with_, without_;
};

// Out of SampleAggregate<Out> (src/traits.rs:11-20)
struct Aggregate
{
    Component component, component2;
    uint32_t kind = 0, percentile = 0;
    uint64_t param = 0;
    Filter filter;
    incerto_aggregate_desc raw() const
    {
        return incerto_aggregate_desc{component.handle, component2.handle, kind, percentile, param, filter.raw()};
    }
    Aggregate filtered(Filter f) const
    {
        Aggregate a = *this;
        a.filter = std::move(f);
        return a;
    }
};
inline Aggregate Sum(Component c) { return Aggregate{c, {}, INCERTO_AGG_SUM}; }
inline Aggregate Count(Component c) { return Aggregate{c, {}, INCERTO_AGG_COUNT}; }
inline Aggregate Minimum(Component c) { return Aggregate{c, {}, INCERTO_AGG_MINIMUM}; } // util/aggregate.rs:73-88
inline Aggregate Maximum(Component c) { return Aggregate{c, {}, INCERTO_AGG_MAXIMUM}; } // :90-105
inline Aggregate Mean(Component c) { return Aggregate{c, {}, INCERTO_AGG_MEAN}; }       // :157-192
inline Aggregate Median(Component c) { return Aggregate{c, {}, INCERTO_AGG_MEDIAN}; }   // :137-155
inline Aggregate Percentile(Component c, uint32_t p) { return Aggregate{c, {}, INCERTO_AGG_PERCENTILE, p}; } // :107-135
inline Aggregate CoinOdds(Component heads, Component tosses) { return Aggregate{heads, tosses, INCERTO_AGG_COIN_ODDS}; }
inline Aggregate FireStats(Component cell) { return Aggregate{cell, {}, INCERTO_AGG_FIRE_STATS}; }
// examples/pandemic_spatial.rs:107-138 (param = INITIAL_POPULATION) and the conflicts / 2 of air_traffic_3d.rs:130-161
inline Aggregate PandemicStats(Component disease_state, uint64_t initial_population)
{
    return Aggregate{disease_state, {}, INCERTO_AGG_PANDEMIC_STATS, 0, initial_population};
}
inline Aggregate AirConflicts(Component c) { return Aggregate{c, {}, INCERTO_AGG_AIR_CONFLICTS}; }

// src/spawner.rs:3-12 — bulk form: n entities, one column per component
class Spawner
{
  public:
    struct Batch
    {
        uint64_t n;
        std::vector<incerto_column_data> cols;
    };
    void spawn_batch(uint64_t n, std::initializer_list<std::pair<Component, const void*>> columns)
    {
        Batch b{n, {}};
        for (auto& c : columns) b.cols.push_back(incerto_column_data{c.first.handle, c.second});
        batches.push_back(std::move(b));
    }
    std::vector<Batch> batches;
};

template <typename T>
struct TimeSeries // src/types/times_series.rs:1-93; borrows from the Simulation until its next run/reset
{
    const uint64_t* time_ = nullptr;
    const T* values_ = nullptr;
    uint64_t len_ = 0, interval_ = 0;
    uint64_t len() const { return len_; }
    bool is_empty() const { return len_ == 0; }
    uint64_t duration() const { return len_ ? time_[len_ - 1] : 0; }
    uint64_t sample_interval() const { return interval_; }
    const uint64_t* time() const { return time_; }
    const T* values() const { return values_; }
};

class Simulation // src/simulation.rs:17-259
{
  public:
    explicit Simulation(incerto_sim* h) : h_(h) {}
    Simulation(Simulation&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    Simulation(const Simulation&) = delete;
    ~Simulation()
    {
        if (h_) incerto_sim_destroy(h_);
    }
    void run(uint64_t num_steps) { check(incerto_sim_run(h_, num_steps)); } // :25-31
    // engine additions with no reference counterpart: enqueue only / wait; several handles (one stream each) may be
    // in flight on one GPU, span_ms gives the device time from the start of this handle's run to the end of `last`'s
    void run_async(uint64_t num_steps) { check(incerto_sim_run_async(h_, num_steps)); }
    void synchronize() { check(incerto_sim_synchronize(h_)); }
    double span_ms(Simulation& last)
    {
        double ms = 0.0;
        check(incerto_sim_span_ms(h_, last.h_, &ms));
        return ms;
    }
    void reset(uint32_t replica) { check(incerto_sim_reset(h_, replica)); }
    // one process per GPU: attach the NCCL communicator (id from incerto_comm_unique_id on rank 0, broadcast by the host)
    void attach_comm(const uint8_t unique_id[128], uint32_t rank, uint32_t world)
    {
        check(incerto_sim_attach_comm(h_, unique_id, rank, world));
    }
    uint64_t count(const Filter& f)                                          // :163-173
    {
        uint64_t out = 0;
        const incerto_filter raw = f.raw();
        check(incerto_sim_count(h_, &raw, &out));
        return out;
    }
    template <typename Out>
    Out sample_aggregate(const Aggregate& a) // :107-152
    {
        Out out{};
        const incerto_aggregate_desc d = a.raw();
        check(incerto_sim_sample_aggregate(h_, &d, &out, sizeof(out)));
        return out;
    }
    // the same aggregate over the entities of all ranks of the attached communicator (collective)
    template <typename Out>
    Out sample_aggregate_global(const Aggregate& a)
    {
        Out out{};
        const incerto_aggregate_desc d = a.raw();
        check(incerto_sim_sample_aggregate_global(h_, &d, &out, sizeof(out)));
        return out;
    }
    template <typename Out, typename Id>
    Out sample(Component c, Component identifier, const Id& id) // :43-65
    {
        Out out{};
        check(incerto_sim_sample(h_, c.handle, identifier.handle, &id, &out, sizeof(out)));
        return out;
    }
    template <typename Out>
    Out sample_single(Component c) // :79-93
    {
        Out out{};
        check(incerto_sim_sample_single(h_, c.handle, &out, sizeof(out)));
        return out;
    }
    template <typename Out>
    TimeSeries<Out> get_aggregate_time_series(const Aggregate& a) // :226-258
    {
        TimeSeries<Out> ts;
        const void* v = nullptr;
        uint64_t rs = 0;
        const incerto_aggregate_desc d = a.raw();
        check(incerto_sim_get_aggregate_time_series(h_, &d, &ts.time_, &v, &ts.len_, &rs, &ts.interval_));
        if (rs != sizeof(Out) && ts.len_) throw std::logic_error("record size mismatch");
        ts.values_ = static_cast<const Out*>(v);
        return ts;
    }
    // column contents of the live entities in uid (= spawn) order; `out` holds n * lanes values
    template <typename T>
    std::vector<T> read_component(Component c, uint32_t values_per_entity = 0)
    {
        uint64_t n = 0;
        check(incerto_sim_num_entities(h_, c.handle, &n));
        std::vector<T> out(n * (values_per_entity ? values_per_entity : c.lanes));
        if (n) check(incerto_sim_read_component(h_, c.handle, out.data(), out.size() * sizeof(T), nullptr));
        return out;
    }
    uint64_t step_number()
    {
        uint64_t s = 0;
        check(incerto_sim_step_number(h_, &s));
        return s;
    }
    incerto_sim* raw() { return h_; }

  private:
    incerto_sim* h_;
};

class SimulationBuilder // src/simulation_builder.rs:25-306
{
  public:
    SimulationBuilder() { check(incerto_builder_new(&h_)); } // :42-56
    ~SimulationBuilder()
    {
        if (h_) incerto_builder_free(h_);
    }
    SimulationBuilder(const SimulationBuilder&) = delete;
    SimulationBuilder& set_seed(uint64_t seed, uint32_t replica = 0)
    {
        check(incerto_builder_set_seed(h_, seed, replica));
        return *this;
    }
    Component component(const char* name, incerto_dtype dtype = INCERTO_MARKER, uint32_t lanes = 1)
    {
        Component c{-1, dtype, lanes};
        check(incerto_builder_register_component(h_, name, dtype, lanes, &c.handle));
        return c;
    }
    SimulationBuilder& add_entity_spawner(std::function<void(Spawner&)> fn) // :154-158
    {
        spawners_.push_back(std::move(fn));
        return *this;
    }
    SimulationBuilder& add_device_spawner(uint32_t kind, const void* params, uint32_t size)
    {
        device_spawners_.push_back({spawners_.size(), kind, std::vector<uint8_t>(static_cast<const uint8_t*>(params), static_cast<const uint8_t*>(params) + size)});
        spawners_.push_back(nullptr);
        return *this;
    }
    template <typename Params>
    SimulationBuilder& add_systems(uint32_t kind, const Params& p) // :66-70
    {
        const incerto_system_desc d{kind, &p, static_cast<uint32_t>(sizeof(Params))};
        check(incerto_builder_add_systems(h_, &d, 1));
        return *this;
    }
    SimulationBuilder& add_spatial_grid(uint32_t dim, Component position, Component c, const int32_t* bounds, incerto_grid* out = nullptr) // :114-139
    {
        check(incerto_builder_add_spatial_grid(h_, dim, bounds, position.handle, c.handle, out));
        return *this;
    }
    SimulationBuilder& record_aggregate_time_series(const Aggregate& a, uint64_t sample_interval) // :230-283
    {
        const incerto_aggregate_desc d = a.raw();
        check(incerto_builder_record_aggregate_time_series(h_, &d, sample_interval));
        return *this;
    }
    SimulationBuilder& record_time_series(Component c, Component identifier, uint64_t sample_interval) // :185-206
    {
        check(incerto_builder_record_time_series(h_, c.handle, identifier.handle, sample_interval));
        return *this;
    }
    Simulation build() // :295-305 — spawners run in insertion order
    {
        for (size_t i = 0; i < spawners_.size(); ++i)
        {
            if (!spawners_[i])
            {
                for (auto& d : device_spawners_)
                    if (d.index == i) check(incerto_builder_add_device_spawner(h_, d.kind, d.params.data(), static_cast<uint32_t>(d.params.size())));
                continue;
            }
            Spawner s;
            spawners_[i](s);
            for (auto& b : s.batches) check(incerto_builder_add_entity_spawner(h_, b.n, b.cols.data(), static_cast<uint32_t>(b.cols.size())));
        }
        incerto_sim* sim = nullptr;
        incerto_builder* h = h_;
        h_ = nullptr; // consumed, also on failure
        check(incerto_builder_build(h, &sim));
        return Simulation(sim);
    }

  private:
    struct DeviceSpawn
    {
        size_t index;
        uint32_t kind;
        std::vector<uint8_t> params;
    };
    incerto_builder* h_ = nullptr;
    std::vector<std::function<void(Spawner&)>> spawners_;
    std::vector<DeviceSpawn> device_spawners_;
};

} // namespace incerto
