// Sampling API, aggregate time series, spatial-grid views.
//
// Reference surface restated here:
//   Simulation::sample / sample_single / sample_aggregate{,_filtered} / count   src/simulation.rs:43-173
//   Simulation::get_time_series / get_aggregate_time_series{,_filtered}           src/simulation.rs:186-258
//   SimulationBuilder::record_time_series / record_aggregate_time_series*         src/simulation_builder.rs:185-283
//   AggregateTimeSeriesPlugin::time_series_sample                                 src/plugins/time_series.rs:68-87
//   TimeSeriesPlugin::{time_series_init,time_series_sample}                       src/plugins/time_series.rs:133-162
//   SpatialGrid queries                                                           src/plugins/spatial_grid.rs:244-389
#include <algorithm>
#include <limits>
#include <array>
#include <cmath>
#include <cstdlib>

#include "kernels_grid.hpp"
#include "modules.hpp"

namespace incerto
{

// raw record written by the device for the generic aggregate kinds
struct GenericRaw
{
    ReduceRecord rec;
    uint64_t sel_value;
    uint64_t sel_error;
};

static bool is_generic_kind(uint32_t k) { return k >= INCERTO_AGG_SUM && k <= INCERTO_AGG_PERCENTILE; }

static size_t raw_size(uint32_t kind)
{
    if (is_generic_kind(kind)) return sizeof(GenericRaw);
    switch (kind)
    {
    case INCERTO_AGG_COIN_ODDS: return 16;
    case INCERTO_AGG_FIRE_STATS: return sizeof(incerto_fire_stats);
    case INCERTO_AGG_PANDEMIC_STATS: return sizeof(incerto_pandemic_stats);
    case INCERTO_AGG_AIR_CONFLICTS: return 16;
    default: return 0;
    }
}

static size_t final_size(const Sim& s, const AggSpec& a)
{
    switch (a.kind)
    {
    case INCERTO_AGG_SUM:
    case INCERTO_AGG_COUNT:
    case INCERTO_AGG_AIR_CONFLICTS: return 8;
    case INCERTO_AGG_MINIMUM:
    case INCERTO_AGG_MAXIMUM:
    case INCERTO_AGG_MEAN:
    case INCERTO_AGG_MEDIAN:
    case INCERTO_AGG_PERCENTILE: return dtype_size(s.comps[a.component].dtype);
    case INCERTO_AGG_COIN_ODDS: return 8;
    case INCERTO_AGG_FIRE_STATS: return sizeof(incerto_fire_stats);
    case INCERTO_AGG_PANDEMIC_STATS: return sizeof(incerto_pandemic_stats);
    default: return 0;
    }
}

static int32_t parse_filter(const incerto_filter* f, FilterSpec& out)
{
    out.with.clear();
    out.without.clear();
    if (!f) return INCERTO_OK;
    if ((f->n_with && !f->with) || (f->n_without && !f->without)) return INCERTO_ERR_INVALID_ARGUMENT;
    for (uint32_t i = 0; i < f->n_with; ++i) out.with.push_back(f->with[i]);
    for (uint32_t i = 0; i < f->n_without; ++i) out.without.push_back(f->without[i]);
    std::sort(out.with.begin(), out.with.end());
    std::sort(out.without.begin(), out.without.end());
    return INCERTO_OK;
}

static int32_t parse_agg(const Sim& s, const incerto_aggregate_desc* d, AggSpec& a)
{
    if (!d) return INCERTO_ERR_INVALID_ARGUMENT;
    a.component = d->component;
    a.component2 = d->component2;
    a.kind = d->kind;
    a.percentile = d->percentile;
    a.param = d->param;
    INC_TRY(parse_filter(&d->filter, a.filter));
    if (a.component < 0 || a.component >= static_cast<int>(s.comps.size()))
    {
        set_error("aggregate names an unregistered component handle %d", a.component);
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    if (raw_size(a.kind) == 0)
    {
        set_error("unknown aggregate kind %u", a.kind);
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    if (a.kind == INCERTO_AGG_PERCENTILE && a.percentile > 100)
    {
        set_error("percentile must be between 0 and 100"); // aggregate.rs:110
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    if (a.kind == INCERTO_AGG_COIN_ODDS && (a.component2 < 0 || a.component2 >= static_cast<int>(s.comps.size())))
        return INCERTO_ERR_INVALID_ARGUMENT;
    return INCERTO_OK;
}

// tables an aggregate over `a.component` ranges over
static std::vector<int> agg_tables(Sim& s, const AggSpec& a)
{
    std::vector<int> r;
    const ComponentInfo& ci = s.comps[a.component];
    for (size_t ti = 0; ti < s.tables.size(); ++ti)
    {
        Table& t = s.tables[ti];
        if (ci.dtype != INCERTO_MARKER && !t.has(a.component)) continue;
        RowFilter rf = row_filter(s, t, a.filter);
        if (rf.impossible) continue;
        r.push_back(static_cast<int>(ti));
    }
    return r;
}

// Launch the kernels of one aggregate over one table; the raw record lands at d_raw.
// ---- aggregates over the entities of ALL ranks of the attached communicator (incerto_sim_sample_aggregate_global):
// the per-rank raw record is combined over NCCL before it is finalised, and Median / Percentile run as a distributed
// radix select (global count, the 256-bin histogram of every pass summed over the ranks).
static int32_t global_combine_record(Sim& s, GenericRaw* d_raw, int dt)
{
    GenericRaw g;
    INC_CUDA(cudaMemcpyAsync(&g, d_raw, sizeof(g), cudaMemcpyDeviceToHost, s.stream));
    INC_CUDA(cudaStreamSynchronize(s.stream));
    const bool flt = dtype_is_float(dt), sgn = dtype_is_signed(dt);
    const uint64_t flip = (!flt && sgn) ? 0x8000000000000000ull : 0ull; // signed order -> unsigned order
    if (g.rec.count == 0)
    {
        // identities, whatever the reduction kernel left behind for an empty selection
        g.rec.sum_bits = 0;
        if (flt)
        {
            const double inf = std::numeric_limits<double>::infinity(), ninf = -inf;
            std::memcpy(&g.rec.min_bits, &inf, 8);
            std::memcpy(&g.rec.max_bits, &ninf, 8);
        }
        else
        {
            g.rec.min_bits = ~0ull ^ flip;
            g.rec.max_bits = 0ull ^ flip;
        }
    }
    uint64_t cnt[2] = {g.rec.count, g.rec.nan_seen};
    INC_TRY(comm_allreduce_host(s, &cnt[0], 1, 0, 0));
    INC_TRY(comm_allreduce_host(s, &cnt[1], 1, 0, 2));
    uint64_t mn = g.rec.min_bits ^ flip, mx = g.rec.max_bits ^ flip;
    INC_TRY(comm_allreduce_host(s, &g.rec.sum_bits, 1, flt ? 1 : 0, 0));
    INC_TRY(comm_allreduce_host(s, &mn, 1, flt ? 1 : 0, 1));
    INC_TRY(comm_allreduce_host(s, &mx, 1, flt ? 1 : 0, 2));
    g.rec.count = cnt[0];
    g.rec.nan_seen = static_cast<uint32_t>(cnt[1]);
    g.rec.min_bits = mn ^ flip;
    g.rec.max_bits = mx ^ flip;
    INC_CUDA(cudaMemcpyAsync(d_raw, &g, sizeof(g), cudaMemcpyHostToDevice, s.stream));
    INC_CUDA(cudaStreamSynchronize(s.stream));
    return INCERTO_OK;
}

struct SelectHook
{
    Sim* s;
    uint64_t* state;
    int32_t rc;
};
static int32_t select_allreduce_histogram(void* ctx)
{
    SelectHook* h = static_cast<SelectHook*>(ctx);
    h->rc = comm_allreduce(*h->s, h->state + 4, 256, 0, 0);
    return h->rc;
}

static int32_t aggregate_launch(Sim& s, const AggSpec& a, int ti, uint8_t* d_raw, bool global = false)
{
    if (global && !is_generic_kind(a.kind) && a.kind != INCERTO_AGG_COIN_ODDS && a.kind != INCERTO_AGG_FIRE_STATS &&
        a.kind != INCERTO_AGG_AIR_CONFLICTS)
    {
        set_error("sample_aggregate_global: aggregate kind %u is not available across ranks", a.kind);
        return INCERTO_ERR_UNSUPPORTED;
    }
    Table& t = s.tables[ti];
    const ComponentInfo& ci = s.comps[a.component];
    FilterSpec f = a.filter;
    if (ci.dtype == INCERTO_MARKER) f.with.push_back(a.component);
    RowFilter rf = row_filter(s, t, f);
    const uint8_t* flags = rf.trivial ? nullptr : t.d_flags;
    ForestModule& fm = s.mods->forest;
    const bool dense = fm.active && fm.table == ti;

    if (a.kind == INCERTO_AGG_FIRE_STATS)
    {
        if (!dense)
        {
            set_error("FIRE_STATS needs the dense forest table");
            return INCERTO_ERR_UNSUPPORTED;
        }
        uint64_t st[5];
        INC_TRY(forest_stats_now(s, st));
        if (global && s.part_world <= 1) INC_TRY(comm_allreduce_host(s, st, 5, 0, 0));
        INC_CUDA(cudaMemcpyAsync(d_raw, st, sizeof(st), cudaMemcpyHostToDevice, s.stream));
        INC_CUDA(cudaStreamSynchronize(s.stream));
        return INCERTO_OK;
    }
    if (dense)
    {
        if (a.kind != INCERTO_AGG_COUNT)
        {
            set_error("only COUNT and FIRE_STATS aggregates are available on the dense forest table");
            return INCERTO_ERR_UNSUPPORTED;
        }
        GenericRaw g;
        std::memset(&g, 0, sizeof(g));
        g.rec.count = t.n;
        INC_CUDA(cudaMemcpyAsync(d_raw, &g, sizeof(g), cudaMemcpyHostToDevice, s.stream));
        INC_CUDA(cudaStreamSynchronize(s.stream));
        if (global) INC_TRY(global_combine_record(s, reinterpret_cast<GenericRaw*>(d_raw), INCERTO_U64));
        return INCERTO_OK;
    }
    if (a.kind == INCERTO_AGG_PANDEMIC_STATS)
    {
        SpatialModule& sm = s.mods->spatial;
        if (!sm.active || sm.table != ti || a.component != sm.p.disease_state)
        {
            set_error("PANDEMIC_STATS is sampled on the disease_state column of the pandemic_spatial systems");
            return INCERTO_ERR_UNSUPPORTED;
        }
        return spatial_stats_now(s, a.param, d_raw);
    }
    if (a.kind == INCERTO_AGG_AIR_CONFLICTS)
    {
        AirModule& am = s.mods->air;
        if (!am.active || am.table != ti || !am.d_acc)
        {
            set_error("AIR_CONFLICTS needs the check_conflicts system on this component's table");
            return INCERTO_ERR_UNSUPPORTED;
        }
        INC_CUDA(cudaMemcpyAsync(d_raw, am.d_acc, 16, cudaMemcpyDeviceToDevice, s.stream));
        // one airspace split over the ranks: a pair across a cut is counted once on either side, so the halves only
        // add up globally (conflicts = sum over the ranks / 2)
        if (global) INC_TRY(comm_allreduce(s, d_raw, 2, 0, 0));
        return INCERTO_OK;
    }
    if (a.kind == INCERTO_AGG_COIN_ODDS)
    {
        Column* h = t.col(a.component);
        Column* c = t.col(a.component2);
        if (!h || !c || s.comps[a.component].dtype != s.comps[a.component2].dtype ||
            (ci.dtype != INCERTO_U32 && ci.dtype != INCERTO_U64))
        {
            set_error("COIN_ODDS needs num_heads/num_tosses columns of equal U32/U64 dtype in one table");
            return INCERTO_ERR_INVALID_ARGUMENT;
        }
        INC_TRY(ensure_scratch(s, 2 * sizeof(double) * kNumSMs * 8));
        KTimer kt(s, "coin_odds", 2.0 * t.n * dtype_size(ci.dtype));
        launch_coin_odds(h->d, c->d, ci.dtype, flags, rf.need_set, rf.need_clear, t.n,
                         reinterpret_cast<double*>(d_raw), s.d_scratch, s.d_ticket, s.stream);
        if (global)
        {
            // record = { f64 sum of heads/tosses, u64 count }
            uint64_t rec[2];
            INC_CUDA(cudaMemcpyAsync(rec, d_raw, 16, cudaMemcpyDeviceToHost, s.stream));
            INC_CUDA(cudaStreamSynchronize(s.stream));
            INC_TRY(comm_allreduce_host(s, &rec[0], 1, 1, 0));
            INC_TRY(comm_allreduce_host(s, &rec[1], 1, 0, 0));
            INC_CUDA(cudaMemcpyAsync(d_raw, rec, 16, cudaMemcpyHostToDevice, s.stream));
            INC_CUDA(cudaStreamSynchronize(s.stream));
        }
        return INCERTO_OK;
    }
    if (!is_generic_kind(a.kind))
    {
        set_error("aggregate kind %u is not implemented in this build", a.kind);
        return INCERTO_ERR_UNSUPPORTED;
    }
    GenericRaw* raw = reinterpret_cast<GenericRaw*>(d_raw);
    INC_TRY(ensure_scratch(s, 64 * kNumSMs * 8 + kSelectStateWords * 8 + 256));
    if (ci.dtype == INCERTO_MARKER || a.kind == INCERTO_AGG_COUNT)
    {
        INC_CUDA(cudaMemsetAsync(raw, 0, sizeof(GenericRaw), s.stream));
        if (rf.trivial)
        {
            const uint64_t n = t.n;
            INC_CUDA(cudaMemcpyAsync(&raw->rec.count, &n, 8, cudaMemcpyHostToDevice, s.stream));
            INC_CUDA(cudaStreamSynchronize(s.stream));
        }
        else
        {
            KTimer kt(s, "count_rows", 1.0 * t.n);
            launch_count_rows(t.d_flags, rf.need_set, rf.need_clear, t.n, &raw->rec.count, s.d_scratch, s.d_ticket,
                              s.stream);
        }
        if (global) INC_TRY(global_combine_record(s, raw, INCERTO_U64));
        return INCERTO_OK;
    }
    Column* col = t.col(a.component);
    {
        KTimer kt(s, "reduce_column", 1.0 * t.n * dtype_size(ci.dtype));
        launch_reduce_column(col->d, ci.dtype, ci.stride(), 0, flags, rf.need_set, rf.need_clear, t.n, &raw->rec,
                             s.d_scratch, s.d_ticket, s.stream);
    }
    if (global) INC_TRY(global_combine_record(s, raw, ci.dtype));
    if (a.kind == INCERTO_AGG_MEDIAN || a.kind == INCERTO_AGG_PERCENTILE)
    {
        uint64_t* state = reinterpret_cast<uint64_t*>(s.d_scratch + 64 * kNumSMs * 8);
        KTimer kt(s, "radix_select", 1.0 * t.n * dtype_size(ci.dtype) * dtype_size(ci.dtype));
        SelectHook hook{&s, state, INCERTO_OK};
        launch_radix_select(col->d, ci.dtype, ci.stride(), 0, flags, rf.need_set, rf.need_clear, t.n, &raw->rec,
                            a.kind == INCERTO_AGG_MEDIAN ? 1u : 2u, a.percentile, state, s.stream,
                            global ? select_allreduce_histogram : nullptr, &hook);
        INC_TRY(hook.rc);
        INC_CUDA(cudaMemcpyAsync(&raw->sel_value, &state[0], 8, cudaMemcpyDeviceToDevice, s.stream));
        INC_CUDA(cudaMemcpyAsync(&raw->sel_error, &state[3], 8, cudaMemcpyDeviceToDevice, s.stream));
    }
    else
    {
        INC_CUDA(cudaMemsetAsync(&raw->sel_value, 0, 16, s.stream));
    }
    INC_CUDA(cudaGetLastError());
    return INCERTO_OK;
}

template <typename T>
static void store_as(void* out, T v)
{
    std::memcpy(out, &v, sizeof(T));
}

// value widened to 64 bit (see ReduceRecord) -> component dtype
static void narrow_store(int dt, uint64_t bits, void* out)
{
    switch (dt)
    {
    case INCERTO_U8: store_as<uint8_t>(out, static_cast<uint8_t>(bits)); break;
    case INCERTO_U16: store_as<uint16_t>(out, static_cast<uint16_t>(bits)); break;
    case INCERTO_U32: store_as<uint32_t>(out, static_cast<uint32_t>(bits)); break;
    case INCERTO_U64: store_as<uint64_t>(out, bits); break;
    case INCERTO_I8: store_as<int8_t>(out, static_cast<int8_t>(static_cast<int64_t>(bits))); break;
    case INCERTO_I16: store_as<int16_t>(out, static_cast<int16_t>(static_cast<int64_t>(bits))); break;
    case INCERTO_I32: store_as<int32_t>(out, static_cast<int32_t>(static_cast<int64_t>(bits))); break;
    case INCERTO_I64: store_as<int64_t>(out, static_cast<int64_t>(bits)); break;
    case INCERTO_F32:
    {
        double d;
        std::memcpy(&d, &bits, 8);
        store_as<float>(out, static_cast<float>(d));
        break;
    }
    case INCERTO_F64: store_as<uint64_t>(out, bits); break;
    default: break;
    }
}

// Combine the per-table raw records and convert to the ABI record.
// Returns AGGREGATE_NO_ENTITIES when nothing matched.
// `reduced`: the raw records were summed over the ranks of the communicator (only matters for AIR_CONFLICTS of an
// x-partitioned airspace, whose per-rank records count conflict ENDS: a pair across a cut has one end on either side)
static int32_t aggregate_finalize(const Sim& s, const AggSpec& a, const uint8_t* raws, size_t n_raws, void* out,
                                  uint64_t* count_out, bool reduced = false)
{
    const int dt = s.comps[a.component].dtype;
    if (is_generic_kind(a.kind))
    {
        GenericRaw tot;
        std::memset(&tot, 0, sizeof(tot));
        bool first = true;
        for (size_t i = 0; i < n_raws; ++i)
        {
            GenericRaw g;
            std::memcpy(&g, raws + i * sizeof(GenericRaw), sizeof(g));
            if (!g.rec.count) continue;
            if (first)
            {
                tot = g;
                first = false;
                continue;
            }
            tot.rec.count += g.rec.count;
            tot.rec.nan_seen |= g.rec.nan_seen;
            if (dtype_is_float(dt))
            {
                double x, y;
                std::memcpy(&x, &tot.rec.sum_bits, 8);
                std::memcpy(&y, &g.rec.sum_bits, 8);
                x += y;
                std::memcpy(&tot.rec.sum_bits, &x, 8);
                double mn0, mn1, mx0, mx1;
                std::memcpy(&mn0, &tot.rec.min_bits, 8);
                std::memcpy(&mn1, &g.rec.min_bits, 8);
                std::memcpy(&mx0, &tot.rec.max_bits, 8);
                std::memcpy(&mx1, &g.rec.max_bits, 8);
                if (mn1 < mn0) tot.rec.min_bits = g.rec.min_bits;
                if (mx1 > mx0) tot.rec.max_bits = g.rec.max_bits;
            }
            else if (dtype_is_signed(dt))
            {
                tot.rec.sum_bits += g.rec.sum_bits;
                if (static_cast<int64_t>(g.rec.min_bits) < static_cast<int64_t>(tot.rec.min_bits)) tot.rec.min_bits = g.rec.min_bits;
                if (static_cast<int64_t>(g.rec.max_bits) > static_cast<int64_t>(tot.rec.max_bits)) tot.rec.max_bits = g.rec.max_bits;
            }
            else
            {
                tot.rec.sum_bits += g.rec.sum_bits;
                if (g.rec.min_bits < tot.rec.min_bits) tot.rec.min_bits = g.rec.min_bits;
                if (g.rec.max_bits > tot.rec.max_bits) tot.rec.max_bits = g.rec.max_bits;
            }
        }
        if (count_out) *count_out = tot.rec.count;
        if (tot.rec.count == 0) return INCERTO_ERR_AGGREGATE_NO_ENTITIES;
        switch (a.kind)
        {
        case INCERTO_AGG_COUNT: store_as<uint64_t>(out, tot.rec.count); return INCERTO_OK;
        case INCERTO_AGG_SUM: store_as<uint64_t>(out, tot.rec.sum_bits); return INCERTO_OK;
        case INCERTO_AGG_MINIMUM:
        case INCERTO_AGG_MAXIMUM:
            if (tot.rec.nan_seen)
            {
                set_error("error during aggregate sampling, cannot compare values (NaN)"); // aggregate.rs:227-231
                return INCERTO_ERR_NOT_COMPARABLE;
            }
            narrow_store(dt, a.kind == INCERTO_AGG_MINIMUM ? tot.rec.min_bits : tot.rec.max_bits, out);
            return INCERTO_OK;
        case INCERTO_AGG_MEAN:
        {
            // sum / cnt in T's own arithmetic (aggregate.rs:168-173)
            if (dt == INCERTO_F64)
            {
                double sum;
                std::memcpy(&sum, &tot.rec.sum_bits, 8);
                store_as<double>(out, sum / static_cast<double>(tot.rec.count));
            }
            else if (dt == INCERTO_F32)
            {
                double sum;
                std::memcpy(&sum, &tot.rec.sum_bits, 8);
                store_as<float>(out, static_cast<float>(sum) / static_cast<float>(tot.rec.count));
            }
            else if (dtype_is_signed(dt))
            {
                // `components.len() as T` then T division (truncating toward zero)
                int64_t cnt = static_cast<int64_t>(tot.rec.count);
                int64_t sum = static_cast<int64_t>(tot.rec.sum_bits);
                uint64_t sb = static_cast<uint64_t>(sum), cb = static_cast<uint64_t>(cnt);
                int64_t s_t, c_t;
                uint8_t tmp[8];
                narrow_store(dt, sb, tmp);
                s_t = dt == INCERTO_I8 ? *reinterpret_cast<int8_t*>(tmp) : dt == INCERTO_I16 ? *reinterpret_cast<int16_t*>(tmp) : dt == INCERTO_I32 ? *reinterpret_cast<int32_t*>(tmp) : *reinterpret_cast<int64_t*>(tmp);
                narrow_store(dt, cb, tmp);
                c_t = dt == INCERTO_I8 ? *reinterpret_cast<int8_t*>(tmp) : dt == INCERTO_I16 ? *reinterpret_cast<int16_t*>(tmp) : dt == INCERTO_I32 ? *reinterpret_cast<int32_t*>(tmp) : *reinterpret_cast<int64_t*>(tmp);
                if (c_t == 0)
                {
                    set_error("Mean: entity count truncates to zero in the sample type");
                    return INCERTO_ERR_INVALID_ARGUMENT;
                }
                narrow_store(dt, static_cast<uint64_t>(s_t / c_t), out);
            }
            else
            {
                const uint64_t mask = dtype_size(dt) == 8 ? ~0ull : ((1ull << (8 * dtype_size(dt))) - 1);
                const uint64_t s_t = tot.rec.sum_bits & mask, c_t = tot.rec.count & mask;
                if (c_t == 0)
                {
                    set_error("Mean: entity count truncates to zero in the sample type");
                    return INCERTO_ERR_INVALID_ARGUMENT;
                }
                narrow_store(dt, s_t / c_t, out);
            }
            return INCERTO_OK;
        }
        case INCERTO_AGG_MEDIAN:
        case INCERTO_AGG_PERCENTILE:
            if (n_raws > 1)
            {
                bool many = false;
                size_t nz = 0;
                for (size_t i = 0; i < n_raws; ++i)
                {
                    GenericRaw g;
                    std::memcpy(&g, raws + i * sizeof(GenericRaw), sizeof(g));
                    if (g.rec.count) ++nz;
                }
                many = nz > 1;
                if (many)
                {
                    set_error("Median/Percentile across several archetype tables is not supported");
                    return INCERTO_ERR_UNSUPPORTED;
                }
            }
            if (tot.rec.nan_seen)
            {
                set_error("error during aggregate sampling, cannot compare values (NaN)");
                return INCERTO_ERR_NOT_COMPARABLE;
            }
            if (tot.sel_error)
            {
                set_error("Percentile index equals the sample count (P = 100): the reference indexes out of range");
                return INCERTO_ERR_PERCENTILE_RANGE;
            }
            narrow_store(dt, tot.sel_value, out);
            return INCERTO_OK;
        default: return INCERTO_ERR_INVALID_ARGUMENT;
        }
    }
    if (a.kind == INCERTO_AGG_COIN_ODDS)
    {
        double sum = 0.0;
        uint64_t cnt = 0;
        for (size_t i = 0; i < n_raws; ++i)
        {
            double sm;
            uint64_t c;
            std::memcpy(&sm, raws + i * 16, 8);
            std::memcpy(&c, raws + i * 16 + 8, 8);
            if (c)
            {
                sum += sm;
                cnt += c;
            }
        }
        if (count_out) *count_out = cnt;
        if (!cnt) return INCERTO_ERR_AGGREGATE_NO_ENTITIES;
        store_as<double>(out, sum / static_cast<double>(cnt)); // coin_toss.rs:48
        return INCERTO_OK;
    }
    if (a.kind == INCERTO_AGG_AIR_CONFLICTS)
    {
        uint64_t twice = 0, indexed = 0;
        for (size_t i = 0; i < n_raws; ++i)
        {
            uint64_t v[2];
            std::memcpy(v, raws + i * 16, 16);
            twice += v[0];
            indexed += v[1];
        }
        if (count_out) *count_out = indexed;
        if (!indexed) return INCERTO_ERR_AGGREGATE_NO_ENTITIES;
        // "Divide by 2 since each conflict is counted twice", air_traffic_3d.rs:160.  One rank of a partitioned
        // airspace reports its conflict ends un-halved: only their sum over the ranks is even.
        store_as<uint64_t>(out, (s.mods->air.part && !reduced) ? twice : twice / 2);
        return INCERTO_OK;
    }
    if (a.kind == INCERTO_AGG_PANDEMIC_STATS)
    {
        incerto_pandemic_stats ps;
        std::memcpy(&ps, raws, sizeof(ps));
        if (count_out) *count_out = ps.total_population;
        if (!ps.total_population) return INCERTO_ERR_AGGREGATE_NO_ENTITIES;
        std::memcpy(out, &ps, sizeof(ps));
        return INCERTO_OK;
    }
    if (a.kind == INCERTO_AGG_FIRE_STATS)
    {
        incerto_fire_stats fs;
        std::memcpy(&fs, raws, sizeof(fs));
        if (count_out) *count_out = fs.total_cells;
        if (!fs.total_cells) return INCERTO_ERR_AGGREGATE_NO_ENTITIES;
        std::memcpy(out, &fs, sizeof(fs));
        return INCERTO_OK;
    }
    return INCERTO_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------- series
static int32_t series_reserve(Sim& s, SeriesInst& se, uint64_t slots)
{
    if (slots <= se.capacity) return INCERTO_OK;
    uint64_t cap = std::max<uint64_t>(slots, se.capacity * 2);
    cap = std::max<uint64_t>(cap, 64);
    const uint64_t rec = se.per_entity ? se.record_size * se.n_rows : raw_size(se.agg.kind);
    uint8_t* nv = nullptr;
    uint64_t* nc = nullptr;
    INC_CUDA(dev_alloc(&nv, cap * rec));
    INC_CUDA(dev_alloc(&nc, cap * sizeof(uint64_t)));
    INC_CUDA(cudaMemsetAsync(nc, 0, cap * sizeof(uint64_t), s.stream));
    if (se.d_values)
    {
        INC_CUDA(cudaMemcpyAsync(nv, se.d_values, se.capacity * rec, cudaMemcpyDeviceToDevice, s.stream));
        INC_CUDA(cudaMemcpyAsync(nc, se.d_counts, se.capacity * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s.stream));
        INC_CUDA(cudaStreamSynchronize(s.stream));
        INC_CUDA(dev_free(se.d_values));
        INC_CUDA(dev_free(se.d_counts));
    }
    se.d_values = nv;
    se.d_counts = nc;
    se.capacity = cap;
    return INCERTO_OK;
}

int32_t series_prepare(Sim& s, uint64_t first_step, uint64_t nsteps)
{
    const uint64_t last = first_step + nsteps - 1;
    for (size_t i = 0; i < s.series.size(); ++i)
    {
        SeriesInst& se = s.series[i];
        if (se.per_entity && se.n_rows == 0)
        {
            const int ti = table_for_components(s, {se.component, se.identifier});
            if (ti < 0) continue; // nothing to record
            se.n_rows = s.tables[ti].n;
            se.record_size = s.comps[se.component].row_bytes();
        }
        INC_TRY(series_reserve(s, se, last / se.interval + 1));
        se.host_valid = false;
    }
    return INCERTO_OK;
}

int32_t series_sample_step(Sim& s, uint64_t step)
{
    for (size_t i = 0; i < s.series.size(); ++i)
    {
        SeriesInst& se = s.series[i];
        if (step % se.interval != 0) continue; // step_number.is_multiple_of(sample_interval), time_series.rs:75
        const uint64_t slot = step / se.interval - 1;
        if (slot >= se.capacity) continue;
        if (static_cast<int>(i) == s.mods->forest.series)
        {
            se.n_slots = std::max(se.n_slots, slot + 1); // written by the fused forest kernel
            continue;
        }
        if (se.per_entity)
        {
            const int ti = table_for_components(s, {se.component, se.identifier});
            if (ti < 0) continue;
            Table& t = s.tables[ti];
            if (t.n != se.n_rows)
            {
                set_error("per-entity time series: table was compacted or grew while recording");
                return INCERTO_ERR_UNSUPPORTED;
            }
            Column* c = t.col(se.component);
            const size_t bytes = se.n_rows * se.record_size;
            KTimer kt(s, "series_snapshot", 2.0 * bytes);
            INC_CUDA(cudaMemcpyAsync(se.d_values + slot * bytes, c->d, bytes, cudaMemcpyDeviceToDevice, s.stream));
            se.n_slots = std::max(se.n_slots, slot + 1);
            continue;
        }
        std::vector<int> tabs = agg_tables(s, se.agg);
        if (tabs.empty())
        {
            se.n_slots = std::max(se.n_slots, slot + 1); // count stays 0: sample skipped
            continue;
        }
        if (tabs.size() > 1)
        {
            set_error("aggregate time series over a component living in several archetype tables is not supported");
            return INCERTO_ERR_UNSUPPORTED;
        }
        const size_t rs = raw_size(se.agg.kind);
        INC_TRY(aggregate_launch(s, se.agg, tabs[0], se.d_values + slot * rs));
        // count: first u64 of GenericRaw; second u64 of coin odds
        const size_t count_off = is_generic_kind(se.agg.kind) ? 0 : ((se.agg.kind == INCERTO_AGG_COIN_ODDS || se.agg.kind == INCERTO_AGG_AIR_CONFLICTS) ? 8 : (se.agg.kind == INCERTO_AGG_PANDEMIC_STATS ? 40 : 32));
        INC_CUDA(cudaMemcpyAsync(se.d_counts + slot, se.d_values + slot * rs + count_off, 8, cudaMemcpyDeviceToDevice,
                                 s.stream));
        se.n_slots = std::max(se.n_slots, slot + 1);
    }
    return INCERTO_OK;
}

void series_after_run(Sim& s)
{
    for (auto& se : s.series) se.host_valid = false;
}

void series_clear(Sim& s)
{
    for (auto& se : s.series)
    {
        se.n_slots = 0;
        se.host_valid = false;
        se.h_time.clear();
        se.h_values.clear();
        if (se.d_counts) cudaMemsetAsync(se.d_counts, 0, se.capacity * sizeof(uint64_t), s.stream);
    }
}

static int32_t series_materialize(Sim& s, SeriesInst& se)
{
    if (se.host_valid) return INCERTO_OK;
    se.h_time.clear();
    se.h_values.clear();
    const size_t rs = raw_size(se.agg.kind);
    const size_t fs = final_size(s, se.agg);
    se.record_size = fs;
    if (se.n_slots)
    {
        std::vector<uint8_t> raw(se.n_slots * rs);
        std::vector<uint64_t> counts(se.n_slots);
        INC_CUDA(cudaMemcpyAsync(raw.data(), se.d_values, raw.size(), cudaMemcpyDeviceToHost, s.stream));
        INC_CUDA(cudaMemcpyAsync(counts.data(), se.d_counts, counts.size() * 8, cudaMemcpyDeviceToHost, s.stream));
        INC_CUDA(cudaStreamSynchronize(s.stream));
        bool reduced = false;
        if (s.comm && s.part_world > 1 &&
            (se.agg.kind == INCERTO_AGG_FIRE_STATS || (se.agg.kind == INCERTO_AGG_AIR_CONFLICTS && s.mods->air.part)))
        {
            reduced = true;
            // row-partitioned forest / x-partitioned airspace: every rank holds partial counts.  COLLECTIVE: every rank of the communicator must
            // read the series the same number of times between runs (documented at incerto_sim_get_aggregate_time_series)
            uint64_t* d = nullptr;
            const size_t words = raw.size() / 8 + counts.size();
            INC_CUDA(dev_alloc(&d, words * 8));
            INC_CUDA(cudaMemcpyAsync(d, raw.data(), raw.size(), cudaMemcpyHostToDevice, s.stream));
            INC_CUDA(cudaMemcpyAsync(d + raw.size() / 8, counts.data(), counts.size() * 8, cudaMemcpyHostToDevice, s.stream));
            int32_t st = comm_allreduce(s, d, static_cast<uint32_t>(words), 0, 0);
            cudaError_t ce = cudaSuccess;
            if (st == INCERTO_OK)
            {
                ce = cudaMemcpyAsync(raw.data(), d, raw.size(), cudaMemcpyDeviceToHost, s.stream);
                if (ce == cudaSuccess)
                    ce = cudaMemcpyAsync(counts.data(), d + raw.size() / 8, counts.size() * 8, cudaMemcpyDeviceToHost, s.stream);
                if (ce == cudaSuccess) ce = cudaStreamSynchronize(s.stream);
            }
            dev_free(d);
            INC_TRY(st);
            INC_CUDA(ce);
        }
        std::vector<uint8_t> rec(fs ? fs : 8);
        for (uint64_t k = 0; k < se.n_slots; ++k)
        {
            if (counts[k] == 0) continue; // `if !component_values.is_empty()`, time_series.rs:79
            uint64_t cnt = 0;
            const int32_t st = aggregate_finalize(s, se.agg, raw.data() + k * rs, 1, rec.data(), &cnt, reduced);
            if (st == INCERTO_ERR_AGGREGATE_NO_ENTITIES) continue;
            INC_TRY(st);
            se.h_time.push_back((k + 1) * se.interval);
            se.h_values.insert(se.h_values.end(), rec.begin(), rec.begin() + fs);
        }
    }
    se.host_valid = true;
    return INCERTO_OK;
}

// -------------------------------------------------------------------- grids
static int32_t grid_alloc_dense(GridInst& g)
{
    if (g.d_cell_start) return INCERTO_OK;
    INC_CUDA(dev_alloc(&g.d_cell_start, (g.n_cells + 1) * 4));
    INC_CUDA(dev_alloc(&g.d_cursor, g.n_cells * 4));
    INC_CUDA(dev_alloc(&g.d_sorted_rows, std::max<uint64_t>(g.cap_rows, 1) * 4));
    return INCERTO_OK;
}

int32_t grids_allocate(Sim& s)
{
    for (auto& g : s.grids)
    {
        g.table = table_for_components(s, {g.position, g.component});
        if (g.table < 0) continue; // nothing spawned yet carries both
        if (s.mods->forest.active && s.mods->forest.table == g.table) continue; // analytic on the dense grid
        const ComponentInfo& pc = s.comps[g.position];
        if ((pc.dtype != INCERTO_I16 && pc.dtype != INCERTO_U16) || pc.lanes != g.dim)
        {
            set_error("spatial grid: position component must have %u lanes of I16/U16 (device layout)", g.dim);
            return INCERTO_ERR_INVALID_ARGUMENT;
        }
        if (!g.bounded)
        {
            // add_spatial_grid(None) (simulation_builder.rs:114-121; in_bounds is "always true if no bounds are set",
            // spatial_grid.rs:265-270): the index is built on demand over the bounding box of the positions the last
            // PreUpdate saw (grid_ensure_built); no position is ever refused
            g.lazy = true;
            g.n_cells = 1;
            g.cap_rows = s.tables[g.table].n;
            continue;
        }
        uint64_t cells = 1;
        for (uint32_t d = 0; d < g.dim; ++d) cells *= static_cast<uint64_t>(g.bmax[d] - g.bmin[d] + 1);
        if (cells >= (1ull << 32))
        {
            set_error("spatial grid has too many cells (%llu)", static_cast<unsigned long long>(cells));
            return INCERTO_ERR_UNSUPPORTED;
        }
        g.n_cells = cells;
        g.cap_rows = s.tables[g.table].n;
        if (const char* e = std::getenv("INCERTO_B200_GRID_LAZY_CELLS"))
            if (cells >= std::strtoull(e, nullptr, 10)) g.lazy = true;
        if (cells > (1ull << 27)) g.lazy = true; // >= 1.5 GB of cell tables: build on demand only
        if (!g.lazy) INC_TRY(grid_alloc_dense(g));
    }
    return INCERTO_OK;
}

static GridDesc grid_desc(const GridInst& g)
{
    GridDesc d;
    d.dim = g.dim;
    for (int i = 0; i < 3; ++i)
    {
        d.bmin[i] = i < static_cast<int>(g.dim) ? g.bmin[i] : 0;
        d.ext[i] = i < static_cast<int>(g.dim) ? g.bmax[i] - g.bmin[i] + 1 : 1;
    }
    return d;
}

static int32_t grid_rebuild(Sim& s, GridInst& g, uint64_t step, bool last_of_run)
{
    if (g.table < 0 || g.n_cells == 0) return INCERTO_OK;
    Table& t = s.tables[g.table];
    FilterSpec f;
    f.with.push_back(g.component);
    RowFilter rf = row_filter(s, t, f);
    Column* pc = t.col(g.position);
    if (g.lazy)
    {
        // only what a host query after this run can observe is kept: the positions at the last PreUpdate
        if (!last_of_run) return INCERTO_OK;
        if (!g.d_snap_pos)
        {
            INC_CUDA(dev_alloc(&g.d_snap_pos, std::max<uint64_t>(g.cap_rows, 1) * pc->row_bytes));
            INC_CUDA(dev_alloc(&g.d_snap_flags, std::max<uint64_t>(g.cap_rows, 1)));
        }
        KTimer kt(s, "grid_snapshot", 2.0 * t.n * (pc->row_bytes + 1));
        INC_CUDA(cudaMemcpyAsync(g.d_snap_pos, pc->d, t.n * pc->row_bytes, cudaMemcpyDeviceToDevice, s.stream));
        INC_CUDA(cudaMemcpyAsync(g.d_snap_flags, t.d_flags, t.n, cudaMemcpyDeviceToDevice, s.stream));
        g.snap_n = t.n;
        g.snap_valid = true;
        g.snap_built = false;
        g.built_step = step;
        return INCERTO_OK;
    }
    INC_TRY(ensure_scratch(s, scan_scratch_bytes(g.n_cells) + 256));
    KTimer kt(s, "grid_build", 16.0 * t.n + 12.0 * g.n_cells);
    launch_grid_build(grid_desc(g), static_cast<const int16_t*>(pc->d), s.comps[g.position].stride(),
                      rf.trivial ? nullptr : t.d_flags, rf.need_set, t.n, g.d_cell_start, g.d_cursor, g.d_sorted_rows,
                      reinterpret_cast<uint32_t*>(s.d_scratch), s.d_error, s.stream);
    INC_CUDA(cudaGetLastError());
    g.built_step = step;
    return INCERTO_OK;
}

// host query on a lazy grid: build the dense table from the snapshot of the last PreUpdate
static int32_t grid_ensure_built(Sim& s, GridInst& g)
{
    if (!g.lazy || !g.snap_valid || g.snap_built) return INCERTO_OK;
    FilterSpec f;
    f.with.push_back(g.component);
    RowFilter rf = row_filter(s, s.tables[g.table], f);
    if (!g.bounded)
    {
        // the table covers the bounding box of the snapshot (the reference's HashMap has no extent at all)
        const uint32_t stride = s.comps[g.position].stride();
        std::vector<int16_t> pos(g.snap_n * stride);
        std::vector<uint8_t> fl(g.snap_n);
        INC_CUDA(cudaMemcpyAsync(pos.data(), g.d_snap_pos, pos.size() * 2, cudaMemcpyDeviceToHost, s.stream));
        INC_CUDA(cudaMemcpyAsync(fl.data(), g.d_snap_flags, fl.size(), cudaMemcpyDeviceToHost, s.stream));
        INC_CUDA(cudaStreamSynchronize(s.stream));
        int32_t lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
        bool any = false;
        for (uint64_t r = 0; r < g.snap_n; ++r)
        {
            if ((fl[r] & rf.need_set) != rf.need_set || !(fl[r] & 1u)) continue;
            for (uint32_t d = 0; d < g.dim; ++d)
            {
                const int32_t v = pos[r * stride + d]; // lanes are read as i16, as the index kernels read them
                lo[d] = any ? std::min(lo[d], v) : v;
                hi[d] = any ? std::max(hi[d], v) : v;
            }
            any = true;
        }
        uint64_t cells = 1;
        for (uint32_t d = 0; d < g.dim; ++d) cells *= static_cast<uint64_t>(hi[d] - lo[d] + 1);
        if (cells > (1ull << 28))
        {
            set_error("unbounded spatial grid: the entities span %llu cells, more than the on-demand index holds (2^28)",
                      static_cast<unsigned long long>(cells));
            return INCERTO_ERR_UNSUPPORTED;
        }
        if (cells != g.n_cells || !g.d_cell_start)
        {
            dev_free(g.d_cell_start);
            dev_free(g.d_cursor);
            dev_free(g.d_sorted_rows);
            g.d_cell_start = g.d_cursor = g.d_sorted_rows = nullptr;
            g.n_cells = cells;
        }
        for (uint32_t d = 0; d < 3; ++d)
        {
            g.bmin[d] = d < g.dim ? lo[d] : 0;
            g.bmax[d] = d < g.dim ? hi[d] : 0;
        }
        g.cap_rows = std::max<uint64_t>(g.cap_rows, g.snap_n);
    }
    INC_TRY(grid_alloc_dense(g));
    INC_TRY(ensure_scratch(s, scan_scratch_bytes(g.n_cells) + 256));
    KTimer kt(s, "grid_build", 16.0 * g.snap_n + 12.0 * g.n_cells);
    launch_grid_build(grid_desc(g), static_cast<const int16_t*>(g.d_snap_pos), s.comps[g.position].stride(),
                      g.d_snap_flags, rf.need_set, g.snap_n, g.d_cell_start, g.d_cursor, g.d_sorted_rows,
                      reinterpret_cast<uint32_t*>(s.d_scratch), s.d_error, s.stream);
    INC_CUDA(cudaGetLastError());
    INC_CUDA(cudaStreamSynchronize(s.stream));
    g.snap_built = true;
    return INCERTO_OK;
}

int32_t grids_rebuild_for_step(Sim& s, uint64_t step, bool last_of_run)
{
    // PreUpdate maintenance (spatial_grid.rs:418-425): the index seen during
    // step k holds the positions as of the end of step k-1.
    for (auto& g : s.grids) INC_TRY(grid_rebuild(s, g, step, last_of_run));
    return INCERTO_OK;
}

} // namespace incerto

using namespace incerto;

#define INC_BUILDER(b)                                    \
    Builder* B = reinterpret_cast<Builder*>(b);           \
    if (!B || !B->sim)                                    \
    {                                                     \
        set_error("null or consumed builder handle");     \
        return INCERTO_ERR_INVALID_ARGUMENT;              \
    }                                                     \
    Sim& S = *B->sim;
#define INC_SIM(s)                                \
    Sim* SP = reinterpret_cast<Sim*>(s);          \
    if (!SP)                                      \
    {                                             \
        set_error("null simulation handle");      \
        return INCERTO_ERR_INVALID_ARGUMENT;      \
    }                                             \
    Sim& S = *SP;                                 \
    if (cudaSetDevice(S.device) != cudaSuccess) return INCERTO_ERR_CUDA;

extern "C"
{

int32_t incerto_builder_add_spatial_grid(incerto_builder* b, uint32_t dim, const int32_t* bounds,
                                         incerto_component position, incerto_component component, incerto_grid* out)
{
    INC_BUILDER(b);
    if ((dim != 2 && dim != 3) || position < 0 || component < 0 || position >= static_cast<int>(S.comps.size()) ||
        component >= static_cast<int>(S.comps.size()))
        return INCERTO_ERR_INVALID_ARGUMENT;
    GridInst g;
    g.dim = dim;
    g.position = position;
    g.component = component;
    if (bounds)
    {
        g.bounded = true;
        for (uint32_t d = 0; d < dim; ++d)
        {
            g.bmin[d] = bounds[d];
            g.bmax[d] = bounds[dim + d];
            if (g.bmin[d] > g.bmax[d])
            {
                set_error("grid bounds: min > max"); // GridBounds::contains asserts, spatial_grid.rs:219-220
                return INCERTO_ERR_INVALID_ARGUMENT;
            }
        }
    }
    S.grids.push_back(g);
    if (out) *out = static_cast<incerto_grid>(S.grids.size() - 1);
    return INCERTO_OK;
}

int32_t incerto_builder_record_aggregate_time_series(incerto_builder* b, const incerto_aggregate_desc* agg,
                                                     uint64_t sample_interval)
{
    INC_BUILDER(b);
    if (sample_interval == 0)
    {
        set_error("sample_interval must be > 0"); // assert!(sample_interval > 0), simulation_builder.rs:271
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    AggSpec a;
    INC_TRY(parse_agg(S, agg, a));
    for (auto& se : S.series)
        if (!se.per_entity && se.agg.same_series_key(a)) return INCERTO_ERR_TIME_SERIES_RECORDING_CONFLICT;
    SeriesInst se;
    se.agg = a;
    se.interval = sample_interval;
    S.series.push_back(se);
    if (a.kind == INCERTO_AGG_FIRE_STATS && a.filter.with.empty() && a.filter.without.empty())
        S.mods->forest.series = static_cast<int>(S.series.size()) - 1;
    return INCERTO_OK;
}

int32_t incerto_builder_record_time_series(incerto_builder* b, incerto_component component,
                                           incerto_component identifier, uint64_t sample_interval)
{
    INC_BUILDER(b);
    if (sample_interval == 0)
    {
        set_error("sample_interval must be > 0"); // simulation_builder.rs:194
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    if (component < 0 || identifier < 0 || component >= static_cast<int>(S.comps.size()) ||
        identifier >= static_cast<int>(S.comps.size()))
        return INCERTO_ERR_INVALID_ARGUMENT;
    for (auto& se : S.series)
        if (se.per_entity && se.component == component && se.identifier == identifier)
            return INCERTO_ERR_TIME_SERIES_RECORDING_CONFLICT;
    SeriesInst se;
    se.per_entity = true;
    se.component = component;
    se.identifier = identifier;
    se.interval = sample_interval;
    S.series.push_back(se);
    return INCERTO_OK;
}

int32_t incerto_sim_count(incerto_sim* s, const incerto_filter* filter, uint64_t* out)
{
    INC_SIM(s);
    if (!out) return INCERTO_ERR_INVALID_ARGUMENT;
    FilterSpec f;
    INC_TRY(parse_filter(filter, f));
    INC_TRY(validate_filter(S, f));
    INC_TRY(ensure_scratch(S, 8 * kNumSMs * 8 + 8 * 64));
    uint64_t total = 0;
    uint64_t* d_out = reinterpret_cast<uint64_t*>(S.d_scratch + 8 * kNumSMs * 8);
    size_t slot = 0;
    std::vector<size_t> pending;
    for (auto& t : S.tables)
    {
        RowFilter rf = row_filter(S, t, f);
        if (rf.impossible || !t.n) continue;
        if (rf.trivial)
        {
            total += t.n;
            continue;
        }
        if (slot >= 64) return INCERTO_ERR_UNSUPPORTED;
        KTimer kt(S, "count_rows", 1.0 * t.n);
        launch_count_rows(t.d_flags, rf.need_set, rf.need_clear, t.n, d_out + slot, S.d_scratch, S.d_ticket, S.stream);
        pending.push_back(slot++);
    }
    if (!pending.empty())
    {
        uint64_t h[64];
        INC_CUDA(cudaMemcpyAsync(h, d_out, slot * 8, cudaMemcpyDeviceToHost, S.stream));
        INC_CUDA(cudaStreamSynchronize(S.stream));
        for (size_t i : pending) total += h[i];
    }
    *out = total;
    return INCERTO_OK;
}

static int32_t sample_aggregate_impl(incerto_sim* s, const incerto_aggregate_desc* agg, void* out, size_t out_size, bool global);

int32_t incerto_sim_sample_aggregate(incerto_sim* s, const incerto_aggregate_desc* agg, void* out, size_t out_size)
{
    return sample_aggregate_impl(s, agg, out, out_size, false);
}

int32_t incerto_sim_sample_aggregate_global(incerto_sim* s, const incerto_aggregate_desc* agg, void* out, size_t out_size)
{
    INC_SIM(s);
    if (!S.comm)
    {
        set_error("sample_aggregate_global: no communicator attached (incerto_sim_attach_comm)");
        return INCERTO_ERR_COMM;
    }
    return sample_aggregate_impl(s, agg, out, out_size, true);
}

static int32_t sample_aggregate_impl(incerto_sim* s, const incerto_aggregate_desc* agg, void* out, size_t out_size, bool global)
{
    INC_SIM(s);
    AggSpec a;
    INC_TRY(parse_agg(S, agg, a));
    if (!S.comps[a.component].exists) return INCERTO_ERR_COMPONENT_DOES_NOT_EXIST;
    INC_TRY(validate_filter(S, a.filter));
    const size_t fs = final_size(S, a);
    if (!out || out_size < fs)
    {
        set_error("sample_aggregate: output buffer too small (%zu < %zu)", out_size, fs);
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    std::vector<int> tabs = agg_tables(S, a);
    if (global && tabs.size() != 1)
    {
        // every rank must take part in the same collectives: one archetype table holding the component on each
        set_error("sample_aggregate_global: the component must live in exactly one archetype table on every rank");
        return INCERTO_ERR_UNSUPPORTED;
    }
    if (tabs.empty()) return INCERTO_ERR_AGGREGATE_NO_ENTITIES;
    if (tabs.size() > 16) return INCERTO_ERR_UNSUPPORTED;
    const size_t rs = raw_size(a.kind);
    // raw records live past the reduction scratch area
    INC_TRY(ensure_scratch(S, 64 * kNumSMs * 8 + kSelectStateWords * 8 + 256 + 16 * 64));
    uint8_t* d_raw = S.d_scratch + 64 * kNumSMs * 8 + kSelectStateWords * 8 + 256;
    std::vector<uint8_t> h(rs * tabs.size());
    for (size_t i = 0; i < tabs.size(); ++i)
    {
        INC_TRY(aggregate_launch(S, a, tabs[i], d_raw + i * 64, global));
        INC_CUDA(cudaMemcpyAsync(h.data() + i * rs, d_raw + i * 64, rs, cudaMemcpyDeviceToHost, S.stream));
    }
    INC_CUDA(cudaStreamSynchronize(S.stream));
    INC_TRY(incerto_sim_synchronize(s));
    if ((a.kind == INCERTO_AGG_MEDIAN || a.kind == INCERTO_AGG_PERCENTILE) && tabs.size() > 1)
    {
        // the query of the reference spans archetypes (simulation.rs:107-121): one selection over the columns of all
        // tables.  The per-table records above give the combined count (which fixes the index) and the NaN flag.
        GenericRaw tot;
        std::memset(&tot, 0, sizeof(tot));
        for (size_t i = 0; i < tabs.size(); ++i)
        {
            GenericRaw g;
            std::memcpy(&g, h.data() + i * rs, sizeof(g));
            tot.rec.count += g.rec.count;
            tot.rec.nan_seen |= g.rec.nan_seen;
        }
        if (tot.rec.count == 0) return INCERTO_ERR_AGGREGATE_NO_ENTITIES;
        const ComponentInfo& ci = S.comps[a.component];
        std::vector<SelectSource> srcs;
        for (int ti : tabs)
        {
            Table& t = S.tables[ti];
            RowFilter rf = row_filter(S, t, a.filter);
            if (rf.impossible || !t.n) continue;
            srcs.push_back(SelectSource{t.col(a.component)->d, ci.stride(), rf.trivial ? nullptr : t.d_flags, rf.need_set,
                                        rf.need_clear, t.n});
        }
        GenericRaw* d_tot = reinterpret_cast<GenericRaw*>(d_raw);
        INC_CUDA(cudaMemcpyAsync(d_tot, &tot, sizeof(tot), cudaMemcpyHostToDevice, S.stream));
        uint64_t* state = reinterpret_cast<uint64_t*>(S.d_scratch + 64 * kNumSMs * 8);
        {
            KTimer kt(S, "radix_select", 1.0 * tot.rec.count * dtype_size(ci.dtype) * dtype_size(ci.dtype));
            launch_radix_select_multi(srcs.data(), srcs.size(), ci.dtype, &d_tot->rec, a.kind == INCERTO_AGG_MEDIAN ? 1u : 2u,
                                      a.percentile, state, S.stream);
        }
        INC_CUDA(cudaMemcpyAsync(&tot.sel_value, &state[0], 8, cudaMemcpyDeviceToHost, S.stream));
        INC_CUDA(cudaMemcpyAsync(&tot.sel_error, &state[3], 8, cudaMemcpyDeviceToHost, S.stream));
        INC_CUDA(cudaStreamSynchronize(S.stream));
        INC_TRY(incerto_sim_synchronize(s));
        return aggregate_finalize(S, a, reinterpret_cast<const uint8_t*>(&tot), 1, out, nullptr);
    }
    return aggregate_finalize(S, a, h.data(), tabs.size(), out, nullptr, global);
}

static int32_t find_by_id(Sim& S, int component, int identifier, const void* id_value, int& table_out, uint64_t& row_out)
{
    if (component < 0 || identifier < 0 || component >= static_cast<int>(S.comps.size()) ||
        identifier >= static_cast<int>(S.comps.size()) || !id_value)
        return INCERTO_ERR_INVALID_ARGUMENT;
    if (!S.comps[component].exists || !S.comps[identifier].exists) return INCERTO_ERR_COMPONENT_DOES_NOT_EXIST;
    const ComponentInfo& ic = S.comps[identifier];
    if (ic.dtype == INCERTO_MARKER || ic.lanes != 1 || dtype_is_float(ic.dtype))
    {
        set_error("identifier must be a scalar integer component (Eq + Hash)");
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    uint64_t id_bits = 0;
    {
        uint8_t tmp[8] = {0};
        std::memcpy(tmp, id_value, dtype_size(ic.dtype));
        switch (ic.dtype)
        {
        case INCERTO_I8: id_bits = static_cast<uint64_t>(static_cast<int64_t>(*reinterpret_cast<int8_t*>(tmp))); break;
        case INCERTO_I16: id_bits = static_cast<uint64_t>(static_cast<int64_t>(*reinterpret_cast<int16_t*>(tmp))); break;
        case INCERTO_I32: id_bits = static_cast<uint64_t>(static_cast<int64_t>(*reinterpret_cast<int32_t*>(tmp))); break;
        default: std::memcpy(&id_bits, tmp, 8); break;
        }
    }
    INC_TRY(ensure_scratch(S, 4096));
    uint64_t* d_out = reinterpret_cast<uint64_t*>(S.d_scratch);
    uint64_t matches = 0;
    table_out = -1;
    for (size_t ti = 0; ti < S.tables.size(); ++ti)
    {
        Table& t = S.tables[ti];
        if (!t.has(component) || !t.has(identifier) || !t.n) continue;
        const uint64_t init[2] = {0, ~0ull};
        INC_CUDA(cudaMemcpyAsync(d_out, init, 16, cudaMemcpyHostToDevice, S.stream));
        {
            KTimer kt(S, "find_id", 1.0 * t.n * dtype_size(ic.dtype));
            launch_find_id(t.col(identifier)->d, ic.dtype, 1, t.all_alive ? nullptr : t.d_flags, t.n, id_bits, d_out,
                           S.stream);
        }
        uint64_t h[2];
        INC_CUDA(cudaMemcpyAsync(h, d_out, 16, cudaMemcpyDeviceToHost, S.stream));
        INC_CUDA(cudaStreamSynchronize(S.stream));
        if (h[0] && table_out < 0)
        {
            table_out = static_cast<int>(ti);
            row_out = h[1];
        }
        matches += h[0];
    }
    if (matches == 0) return INCERTO_ERR_ENTITY_IDENTIFIER_NOT_FOUND;
    if (matches > 1) return INCERTO_ERR_ENTITY_IDENTIFIER_NOT_UNIQUE;
    return INCERTO_OK;
}

static int32_t copy_row_out(Sim& S, Table& t, int component, uint64_t row, void* out, size_t out_size)
{
    const ComponentInfo& ci = S.comps[component];
    const size_t bytes = dtype_size(ci.dtype) * ci.lanes;
    if (!out || out_size < bytes)
    {
        set_error("output buffer too small (%zu < %zu)", out_size, bytes);
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    Column* c = t.col(component);
    INC_CUDA(cudaMemcpyAsync(out, static_cast<uint8_t*>(c->d) + row * c->row_bytes, bytes, cudaMemcpyDeviceToHost, S.stream));
    INC_CUDA(cudaStreamSynchronize(S.stream));
    return INCERTO_OK;
}

int32_t incerto_sim_sample(incerto_sim* s, incerto_component component, incerto_component identifier,
                           const void* id_value, void* out, size_t out_size)
{
    INC_SIM(s);
    int ti = -1;
    uint64_t row = 0;
    INC_TRY(find_by_id(S, component, identifier, id_value, ti, row));
    return copy_row_out(S, S.tables[ti], component, row, out, out_size);
}

int32_t incerto_sim_sample_single(incerto_sim* s, incerto_component component, void* out, size_t out_size)
{
    INC_SIM(s);
    if (component < 0 || component >= static_cast<int>(S.comps.size())) return INCERTO_ERR_INVALID_ARGUMENT;
    if (!S.comps[component].exists) return INCERTO_ERR_COMPONENT_DOES_NOT_EXIST;
    if (S.comps[component].dtype == INCERTO_MARKER) return INCERTO_ERR_INVALID_ARGUMENT;
    uint64_t found = 0, row = 0;
    int tab = -1;
    for (size_t ti = 0; ti < S.tables.size(); ++ti)
    {
        Table& t = S.tables[ti];
        if (!t.has(component) || !t.n) continue;
        if (S.mods->forest.active && S.mods->forest.table == static_cast<int>(ti))
        {
            found += t.n;
            continue;
        }
        if (t.all_alive)
        {
            if (!found)
            {
                tab = static_cast<int>(ti);
                row = 0;
            }
            found += t.n;
        }
        else
        {
            std::vector<uint8_t> fl(t.n);
            INC_CUDA(cudaMemcpyAsync(fl.data(), t.d_flags, t.n, cudaMemcpyDeviceToHost, S.stream));
            INC_CUDA(cudaStreamSynchronize(S.stream));
            for (uint64_t r = 0; r < t.n; ++r)
                if (fl[r] & kAliveBit)
                {
                    if (!found)
                    {
                        tab = static_cast<int>(ti);
                        row = r;
                    }
                    ++found;
                }
        }
    }
    if (found == 0) return INCERTO_ERR_SINGLE_NO_ENTITIES;       // QuerySingleError::NoEntities, simulation.rs:88
    if (found > 1) return INCERTO_ERR_SINGLE_MULTIPLE_ENTITIES;  // :89
    if (tab < 0) return INCERTO_ERR_UNSUPPORTED;
    return copy_row_out(S, S.tables[tab], component, row, out, out_size);
}

int32_t incerto_sim_get_aggregate_time_series(incerto_sim* s, const incerto_aggregate_desc* agg, const uint64_t** time,
                                              const void** values, uint64_t* len, uint64_t* record_size,
                                              uint64_t* sample_interval)
{
    INC_SIM(s);
    AggSpec a;
    INC_TRY(parse_agg(S, agg, a));
    for (auto& se : S.series)
    {
        if (se.per_entity || !se.agg.same_series_key(a)) continue;
        INC_TRY(incerto_sim_synchronize(s));
        INC_TRY(series_materialize(S, se));
        if (time) *time = se.h_time.data();
        if (values) *values = se.h_values.data();
        if (len) *len = se.h_time.size();
        if (record_size) *record_size = se.record_size;
        if (sample_interval) *sample_interval = se.interval;
        return INCERTO_OK;
    }
    return INCERTO_ERR_TIME_SERIES_NOT_RECORDED; // simulation.rs:253
}

int32_t incerto_sim_get_time_series(incerto_sim* s, incerto_component component, incerto_component identifier,
                                    const void* id_value, const uint64_t** time, const void** values, uint64_t* len,
                                    uint64_t* record_size, uint64_t* sample_interval)
{
    INC_SIM(s);
    for (auto& se : S.series)
    {
        if (!se.per_entity || se.component != component || se.identifier != identifier) continue;
        INC_TRY(incerto_sim_synchronize(s));
        int ti = -1;
        uint64_t row = 0;
        INC_TRY(find_by_id(S, component, identifier, id_value, ti, row));
        const ComponentInfo& ci = S.comps[component];
        const size_t bytes = dtype_size(ci.dtype) * ci.lanes;
        se.h_time.clear();
        se.h_values.assign(se.n_slots * bytes, 0);
        if (se.n_slots)
        {
            INC_CUDA(cudaMemcpy2DAsync(se.h_values.data(), bytes, se.d_values + row * se.record_size,
                                       se.n_rows * se.record_size, bytes, se.n_slots, cudaMemcpyDeviceToHost, S.stream));
            INC_CUDA(cudaStreamSynchronize(S.stream));
        }
        for (uint64_t k = 0; k < se.n_slots; ++k) se.h_time.push_back((k + 1) * se.interval);
        if (time) *time = se.h_time.data();
        if (values) *values = se.h_values.data();
        if (len) *len = se.n_slots;
        if (record_size) *record_size = bytes;
        if (sample_interval) *sample_interval = se.interval;
        return INCERTO_OK;
    }
    // the reference finds no TimeSeriesData<C, Id, O> component: try_query fails (simulation.rs:196-199)
    return INCERTO_ERR_COMPONENT_DOES_NOT_EXIST;
}

// -------------------------------------------------------------- introspection
int32_t incerto_sim_num_entities(incerto_sim* s, incerto_component component, uint64_t* out)
{
    INC_SIM(s);
    if (component < 0 || component >= static_cast<int>(S.comps.size()) || !out) return INCERTO_ERR_INVALID_ARGUMENT;
    incerto_filter f;
    f.with = &component;
    f.n_with = 1;
    f.without = nullptr;
    f.n_without = 0;
    if (!S.comps[component].exists)
    {
        *out = 0;
        return INCERTO_OK;
    }
    return incerto_sim_count(s, &f, out);
}

int32_t incerto_sim_read_component(incerto_sim* s, incerto_component component, void* out, size_t out_size,
                                   uint64_t* uids_out)
{
    INC_SIM(s);
    if (component < 0 || component >= static_cast<int>(S.comps.size()) || !out) return INCERTO_ERR_INVALID_ARGUMENT;
    INC_TRY(incerto_sim_synchronize(s));
    const ComponentInfo& ci = S.comps[component];
    const size_t es = dtype_size(ci.dtype);
    const size_t abi_bytes = ci.dtype == INCERTO_MARKER ? 1 : es * ci.lanes;
    uint8_t* o = static_cast<uint8_t*>(out);
    size_t written = 0;
    uint64_t k = 0;
    ForestModule& fm = S.mods->forest;
    for (size_t ti = 0; ti < S.tables.size(); ++ti)
    {
        Table& t = S.tables[ti];
        if (ci.dtype != INCERTO_MARKER && !t.has(component)) continue;
        if (!t.n) continue;
        if (fm.active && fm.table == static_cast<int>(ti))
        {
            const uint64_t rows = static_cast<uint64_t>(fm.x_end - fm.x_begin) * fm.H + fm.n_overlay;
            if (written + rows * abi_bytes > out_size) return INCERTO_ERR_INVALID_ARGUMENT;
            if (component == fm.comp_cell)
            {
                INC_TRY(forest_read_cells(S, o + written));
            }
            else if (component == fm.comp_pos)
            {
                // positions of the dense prefix are implicit: uid = x*H + y
                uint8_t* p = o + written;
                for (int32_t x = fm.x_begin; x < fm.x_end; ++x)
                    for (int32_t y = 0; y < fm.H; ++y)
                    {
                        if (es == 2)
                        {
                            int16_t v[2] = {static_cast<int16_t>(x), static_cast<int16_t>(y)};
                            std::memcpy(p, v, 4);
                        }
                        else
                        {
                            int32_t v[2] = {x, y};
                            std::memcpy(p, v, 8);
                        }
                        p += abi_bytes;
                    }
                for (uint32_t q = 0; q < fm.n_overlay; ++q)
                {
                    if (es == 2)
                        std::memcpy(p, &fm.h_ov_pos[2 * q], 4);
                    else
                    {
                        int32_t v[2] = {fm.h_ov_pos[2 * q], fm.h_ov_pos[2 * q + 1]};
                        std::memcpy(p, v, 8);
                    }
                    p += abi_bytes;
                }
            }
            else
                return INCERTO_ERR_UNSUPPORTED;
            if (uids_out)
            {
                for (int32_t x = fm.x_begin; x < fm.x_end; ++x)
                    for (int32_t y = 0; y < fm.H; ++y) uids_out[k++] = static_cast<uint64_t>(x) * fm.H + y;
                for (uint32_t q = 0; q < fm.n_overlay; ++q) uids_out[k++] = static_cast<uint64_t>(fm.W) * fm.H + q;
            }
            written += rows * abi_bytes;
            continue;
        }
        if (ci.managed_bytes)
        {
            size_t w = 0;
            if (S.mods->traders.active && S.mods->traders.p.portfolio == component)
                INC_TRY(traders_read_portfolio(S, static_cast<int>(ti), o + written, out_size - written, &w));
            else if (S.mods->spatial.active && S.mods->spatial.p.contact_history == component)
                INC_TRY(spatial_read_history(S, static_cast<int>(ti), o + written, out_size - written, &w));
            else
                return INCERTO_ERR_UNSUPPORTED;
            written += w;
            continue;
        }
        std::vector<uint32_t> rows, uids;
        INC_TRY(live_rows_by_uid(S, t, rows, uids));
        std::vector<uint8_t> fl;
        if (ci.dtype == INCERTO_MARKER)
        {
            fl.resize(t.n);
            INC_CUDA(cudaMemcpyAsync(fl.data(), t.d_flags, t.n, cudaMemcpyDeviceToHost, S.stream));
        }
        std::vector<uint8_t> data;
        size_t rb = 0;
        if (ci.dtype != INCERTO_MARKER)
        {
            Column* c = t.col(component);
            rb = c->row_bytes;
            data.resize(t.n * rb);
            INC_CUDA(cudaMemcpyAsync(data.data(), c->d, data.size(), cudaMemcpyDeviceToHost, S.stream));
        }
        INC_CUDA(cudaStreamSynchronize(S.stream));
        for (size_t j = 0; j < rows.size(); ++j)
        {
            const uint32_t r = rows[j];
            if (written + abi_bytes > out_size)
            {
                set_error("read_component: output buffer too small");
                return INCERTO_ERR_INVALID_ARGUMENT;
            }
            if (ci.dtype == INCERTO_MARKER)
                o[written] = (fl[r] >> ci.marker_bit) & 1u;
            else
                std::memcpy(o + written, &data[r * rb], abi_bytes);
            written += abi_bytes;
            if (uids_out) uids_out[k++] = uids[j];
        }
    }
    return INCERTO_OK;
}

int32_t incerto_sim_read_marker(incerto_sim* s, incerto_component marker, incerto_component rows_of, uint8_t* out,
                                size_t out_size)
{
    INC_SIM(s);
    if (marker < 0 || rows_of < 0 || marker >= static_cast<int>(S.comps.size()) ||
        rows_of >= static_cast<int>(S.comps.size()) || !out || S.comps[marker].dtype != INCERTO_MARKER ||
        S.comps[rows_of].dtype == INCERTO_MARKER)
        return INCERTO_ERR_INVALID_ARGUMENT;
    INC_TRY(incerto_sim_synchronize(s));
    size_t written = 0;
    for (auto& t : S.tables)
    {
        if (!t.has(rows_of) || !t.n || !t.d_flags) continue;
        std::vector<uint32_t> rows, uids;
        INC_TRY(live_rows_by_uid(S, t, rows, uids));
        std::vector<uint8_t> fl(t.n);
        INC_CUDA(cudaMemcpyAsync(fl.data(), t.d_flags, t.n, cudaMemcpyDeviceToHost, S.stream));
        INC_CUDA(cudaStreamSynchronize(S.stream));
        for (uint32_t r : rows)
        {
            if (written >= out_size)
            {
                set_error("read_marker: output buffer too small");
                return INCERTO_ERR_INVALID_ARGUMENT;
            }
            out[written++] = (fl[r] >> S.comps[marker].marker_bit) & 1u;
        }
    }
    return INCERTO_OK;
}

// ------------------------------------------------------------ grid queries
static int32_t grid_lookup(Sim& S, incerto_grid g, GridInst*& out)
{
    if (g < 0 || g >= static_cast<int>(S.grids.size()))
    {
        set_error("unknown spatial grid handle %d", g);
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    out = &S.grids[g];
    return INCERTO_OK;
}

static bool grid_contains(const GridInst& g, const int32_t* pos)
{
    if (!g.bounded) return true; // is_none_or, spatial_grid.rs:268-270
    for (uint32_t d = 0; d < g.dim; ++d)
        if (pos[d] < g.bmin[d] || pos[d] > g.bmax[d]) return false;
    return true;
}

static int32_t grid_cell_uids(Sim& S, GridInst& g, const int32_t* pos, std::vector<uint64_t>& out)
{
    if (!grid_contains(g, pos)) return INCERTO_OK; // outside the table: no entities can be there
    ForestModule& fm = S.mods->forest;
    if (g.table >= 0 && fm.active && fm.table == g.table)
    {
        if (pos[0] >= 0 && pos[0] < fm.W && pos[1] >= 0 && pos[1] < fm.H)
            out.push_back(static_cast<uint64_t>(pos[0]) * fm.H + pos[1]);
        for (uint32_t q = 0; q < fm.n_overlay; ++q)
            if (fm.h_ov_pos[2 * q] == pos[0] && fm.h_ov_pos[2 * q + 1] == pos[1])
                out.push_back(static_cast<uint64_t>(fm.W) * fm.H + q);
        return INCERTO_OK;
    }
    INC_TRY(grid_ensure_built(S, g));
    if (g.table < 0 || !g.d_cell_start || g.built_step == ~0ull) return INCERTO_OK;
    if (!g.bounded) // the on-demand table of an unbounded grid covers the entities' bounding box: nobody lives outside it
        for (uint32_t k = 0; k < g.dim; ++k)
            if (pos[k] < g.bmin[k] || pos[k] > g.bmax[k]) return INCERTO_OK;
    GridDesc d = grid_desc(g);
    const uint64_t key = grid_cell_key(d, pos[0], pos[1], g.dim == 3 ? pos[2] : 0);
    uint32_t se[2];
    INC_CUDA(cudaMemcpyAsync(se, g.d_cell_start + key, 8, cudaMemcpyDeviceToHost, S.stream));
    INC_CUDA(cudaStreamSynchronize(S.stream));
    const uint32_t cnt = se[1] - se[0];
    if (!cnt) return INCERTO_OK;
    std::vector<uint32_t> rows(cnt);
    INC_CUDA(cudaMemcpyAsync(rows.data(), g.d_sorted_rows + se[0], cnt * 4, cudaMemcpyDeviceToHost, S.stream));
    INC_CUDA(cudaStreamSynchronize(S.stream));
    Table& t = S.tables[g.table];
    for (uint32_t r : rows)
    {
        uint64_t uid = r;
        if (t.d_uid)
        {
            uint32_t u;
            INC_CUDA(cudaMemcpy(&u, t.d_uid + r, 4, cudaMemcpyDeviceToHost));
            uid = u;
        }
        out.push_back(uid);
    }
    return INCERTO_OK;
}

static int32_t emit_uids(const std::vector<uint64_t>& v, uint64_t* out_uids, uint64_t capacity, uint64_t* n_out)
{
    if (n_out) *n_out = v.size();
    if (out_uids)
        for (size_t i = 0; i < v.size() && i < capacity; ++i) out_uids[i] = v[i];
    return INCERTO_OK;
}

int32_t incerto_sim_grid_entities_at(incerto_sim* s, incerto_grid g, const int32_t* pos, uint64_t* out_uids,
                                     uint64_t capacity, uint64_t* n_out)
{
    INC_SIM(s);
    GridInst* G;
    INC_TRY(grid_lookup(S, g, G));
    if (!pos) return INCERTO_ERR_INVALID_ARGUMENT;
    INC_TRY(incerto_sim_synchronize(s));
    std::vector<uint64_t> v;
    INC_TRY(grid_cell_uids(S, *G, pos, v));
    return emit_uids(v, out_uids, capacity, n_out);
}

int32_t incerto_sim_grid_neighbors_of(incerto_sim* s, incerto_grid g, const int32_t* pos, uint32_t orthogonal,
                                      uint64_t* out_uids, uint64_t capacity, uint64_t* n_out)
{
    INC_SIM(s);
    GridInst* G;
    INC_TRY(grid_lookup(S, g, G));
    if (!pos) return INCERTO_ERR_INVALID_ARGUMENT;
    INC_TRY(incerto_sim_synchronize(s));
    std::vector<uint64_t> v;
    // neighbour enumeration orders of spatial_grid.rs:55-109
    std::vector<std::array<int32_t, 3>> offs;
    if (G->dim == 2)
    {
        if (orthogonal)
            offs = {{0, -1, 0}, {-1, 0, 0}, {1, 0, 0}, {0, 1, 0}}; // N, W, E, S (:67)
        else
            offs = {{-1, -1, 0}, {0, -1, 0}, {1, -1, 0}, {-1, 0, 0}, {1, 0, 0}, {-1, 1, 0}, {0, 1, 0}, {1, 1, 0}}; // :59-63
    }
    else
    {
        if (orthogonal)
            offs = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}}; // W, E, N, S, Down, Up (:101)
        else
            for (int dx = -1; dx <= 1; ++dx)
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dz = -1; dz <= 1; ++dz)
                        if (dx || dy || dz) offs.push_back({dx, dy, dz}); // :79-96
    }
    for (auto& o : offs)
    {
        int32_t q[3] = {pos[0] + o[0], pos[1] + o[1], G->dim == 3 ? pos[2] + o[2] : 0};
        if (!grid_contains(*G, q)) continue; // bounds filter of neighbor queries (:337-340, :361-364)
        INC_TRY(grid_cell_uids(S, *G, q, v));
    }
    return emit_uids(v, out_uids, capacity, n_out);
}

int32_t incerto_sim_grid_is_empty(incerto_sim* s, incerto_grid g, const int32_t* pos, uint32_t* out)
{
    INC_SIM(s);
    GridInst* G;
    INC_TRY(grid_lookup(S, g, G));
    if (!pos || !out) return INCERTO_ERR_INVALID_ARGUMENT;
    INC_TRY(incerto_sim_synchronize(s));
    std::vector<uint64_t> v;
    INC_TRY(grid_cell_uids(S, *G, pos, v));
    *out = v.empty() ? 1u : 0u;
    return INCERTO_OK;
}

int32_t incerto_sim_grid_num_entities(incerto_sim* s, incerto_grid g, uint64_t* out)
{
    INC_SIM(s);
    GridInst* G;
    INC_TRY(grid_lookup(S, g, G));
    if (!out) return INCERTO_ERR_INVALID_ARGUMENT;
    INC_TRY(incerto_sim_synchronize(s));
    *out = 0;
    ForestModule& fm = S.mods->forest;
    if (G->table >= 0 && fm.active && fm.table == G->table)
    {
        *out = S.step > 1 ? S.tables[G->table].n : 0;
        return INCERTO_OK;
    }
    INC_TRY(grid_ensure_built(S, *G));
    if (G->table < 0 || !G->d_cell_start || G->built_step == ~0ull) return INCERTO_OK;
    uint32_t total = 0;
    INC_CUDA(cudaMemcpy(&total, G->d_cell_start + G->n_cells, 4, cudaMemcpyDeviceToHost));
    *out = total;
    return INCERTO_OK;
}

int32_t incerto_sim_grid_position_of(incerto_sim* s, incerto_grid g, uint64_t uid, int32_t* out_pos)
{
    INC_SIM(s);
    GridInst* G;
    INC_TRY(grid_lookup(S, g, G));
    if (!out_pos) return INCERTO_ERR_INVALID_ARGUMENT;
    if (G->table < 0) return INCERTO_ERR_ENTITY_IDENTIFIER_NOT_FOUND;
    Table& t = S.tables[G->table];
    ForestModule& fm = S.mods->forest;
    if (fm.active && fm.table == G->table)
    {
        const uint64_t cells = static_cast<uint64_t>(fm.W) * fm.H;
        if (uid < cells)
        {
            out_pos[0] = static_cast<int32_t>(uid / fm.H);
            out_pos[1] = static_cast<int32_t>(uid % fm.H);
            return INCERTO_OK;
        }
        if (uid - cells < fm.n_overlay)
        {
            out_pos[0] = fm.h_ov_pos[2 * (uid - cells)];
            out_pos[1] = fm.h_ov_pos[2 * (uid - cells) + 1];
            return INCERTO_OK;
        }
        return INCERTO_ERR_ENTITY_IDENTIFIER_NOT_FOUND;
    }
    uint64_t row = uid;
    if (t.d_uid)
    {
        std::vector<uint32_t> rows, uids;
        INC_TRY(live_rows_by_uid(S, t, rows, uids));
        auto it = std::lower_bound(uids.begin(), uids.end(), static_cast<uint32_t>(uid));
        if (uid >= (1ull << 32) || it == uids.end() || *it != uid) return INCERTO_ERR_ENTITY_IDENTIFIER_NOT_FOUND;
        row = rows[it - uids.begin()];
    }
    else if (uid >= t.n)
        return INCERTO_ERR_ENTITY_IDENTIFIER_NOT_FOUND;
    int16_t p[4];
    Column* c = t.col(G->position);
    INC_CUDA(cudaMemcpy(p, static_cast<uint8_t*>(c->d) + row * c->row_bytes, c->row_bytes, cudaMemcpyDeviceToHost));
    for (uint32_t d = 0; d < G->dim; ++d) out_pos[d] = p[d];
    return INCERTO_OK;
}

int32_t incerto_sim_grid_in_bounds(incerto_sim* s, incerto_grid g, const int32_t* pos, uint32_t* out)
{
    INC_SIM(s);
    GridInst* G;
    INC_TRY(grid_lookup(S, g, G));
    if (!pos || !out) return INCERTO_ERR_INVALID_ARGUMENT;
    *out = grid_contains(*G, pos) ? 1u : 0u;
    return INCERTO_OK;
}

} // extern "C"
