// Internal engine structures (host side). Not part of the ABI.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/incerto_b200.h"

namespace incerto
{

constexpr int kMaxMarkers = 7;     // bits 1..7 of the per-row flags byte; bit 0 = alive
constexpr uint8_t kAliveBit = 1u;

void set_error(const char* fmt, ...);
int32_t cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define INC_CUDA(call)                                                     \
    do                                                                     \
    {                                                                      \
        cudaError_t e__ = (call);                                          \
        if (e__ != cudaSuccess) return ::incerto::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define INC_TRY(call)                        \
    do                                       \
    {                                        \
        int32_t s__ = (call);                \
        if (s__ != INCERTO_OK) return s__;   \
    } while (0)

inline size_t dtype_size(int dt)
{
    switch (dt)
    {
    case INCERTO_U8:
    case INCERTO_I8: return 1;
    case INCERTO_U16:
    case INCERTO_I16: return 2;
    case INCERTO_U32:
    case INCERTO_I32:
    case INCERTO_F32: return 4;
    case INCERTO_U64:
    case INCERTO_I64:
    case INCERTO_F64: return 8;
    default: return 0;
    }
}
inline bool dtype_is_float(int dt) { return dt == INCERTO_F32 || dt == INCERTO_F64; }
inline bool dtype_is_signed(int dt) { return dt >= INCERTO_I8 && dt <= INCERTO_I64; }

inline uint64_t bernoulli_threshold(double p)
{
    if (!(p > 0.0)) return 0;
    if (p >= 1.0) return 1ull << 32;
    return static_cast<uint64_t>(p * 4294967296.0); // floor for positive values
}

struct ComponentInfo
{
    std::string name;
    int dtype = INCERTO_MARKER;
    uint32_t lanes = 1;
    int marker_bit = -1;   // 1..7 for markers
    bool exists = false;   // "registered in the world": spawned, or referenced by a system/grid
    // engine-managed wide components (ContactHistory, portfolio): bytes per row override
    uint32_t managed_bytes = 0;
    uint32_t stride() const { return lanes == 3 ? 4u : lanes; } // device stride in elements
    size_t row_bytes() const { return managed_bytes ? managed_bytes : dtype_size(dtype) * stride(); }
};

struct Column
{
    int comp = -1;
    void* d = nullptr;     // n * row_bytes
    void* d_alt = nullptr; // second buffer for double-buffered kernels (allocated on demand)
    size_t row_bytes = 0;
};

struct Table
{
    uint64_t n = 0;           // rows, including dead ones until compaction
    uint64_t n_spawned = 0;   // uid high-water mark
    std::vector<int> signature; // sorted data components
    std::vector<Column> cols;
    uint8_t* d_flags = nullptr;
    uint8_t* d_flags_alt = nullptr;
    uint32_t* d_uid = nullptr; // nullptr: uid == row index
    bool all_alive = true;
    // dense 2-D grid prefix (forest fire): rows [0, dense_w*dense_h) are cells in x-major order
    int32_t dense_w = 0, dense_h = 0;
    // row partition of the dense prefix: this process holds x in [x_begin, x_end)
    int32_t x_begin = 0, x_end = 0;

    Column* col(int comp)
    {
        for (auto& c : cols)
            if (c.comp == comp) return &c;
        return nullptr;
    }
    bool has(int comp) const
    {
        for (auto& c : cols)
            if (c.comp == comp) return true;
        return false;
    }
};

struct HostSpawn
{
    bool device_kind = false;
    uint32_t kind = 0;
    std::vector<uint8_t> params;
    uint64_t n = 0;
    struct Col
    {
        int comp;
        bool has_data;
        std::vector<uint8_t> data;      // owned copy (incerto_builder_add_entity_spawner)
        const uint8_t* ext = nullptr;   // or the caller's buffer, valid until build returns (..._borrowed)
        const uint8_t* bytes() const { return ext ? ext : data.data(); }
    };
    std::vector<Col> cols;
    bool borrowed = false; // the data is gone after build: the simulation cannot be reset
};

struct SystemInst
{
    uint32_t kind = 0;
    std::vector<uint8_t> params;
    uint32_t order = 0; // insertion ordinal (stable tie-break)
};

struct FilterSpec
{
    std::vector<int> with, without;
    bool operator==(const FilterSpec& o) const { return with == o.with && without == o.without; }
};

struct AggSpec
{
    int component = -1, component2 = -1;
    uint32_t kind = 0, percentile = 0;
    uint64_t param = 0;
    FilterSpec filter;
    bool same_series_key(const AggSpec& o) const
    {
        // (C, F, O) of TimeSeriesData<C, F, O>, src/simulation_builder.rs:274-278
        return component == o.component && kind == o.kind && percentile == o.percentile && filter == o.filter;
    }
};

struct SeriesInst
{
    bool per_entity = false;
    AggSpec agg;            // aggregate series
    int component = -1, identifier = -1; // per-entity series
    uint64_t interval = 1;
    uint64_t record_size = 0;
    // device ring: records written by the step program; slot = number of sampling instants so far
    uint8_t* d_values = nullptr;   // capacity * record_size (per-entity: capacity * n_rows * record_size)
    uint64_t* d_counts = nullptr;  // matching entity count per slot (0 => sample skipped, time_series.rs:79)
    uint64_t capacity = 0;
    uint64_t n_slots = 0;          // sampling instants recorded so far
    uint64_t first_step = 0;
    std::vector<uint64_t> slot_steps;
    uint64_t n_rows = 0;           // per-entity: rows at first sample
    // host mirror handed out by get_*_time_series
    std::vector<uint64_t> h_time;
    std::vector<uint8_t> h_values;
    bool host_valid = false;
};

struct GridInst
{
    uint32_t dim = 2;
    bool bounded = false;
    int32_t bmin[3] = {0, 0, 0}, bmax[3] = {0, 0, 0};
    int position = -1, component = -1;
    int table = -1;
    uint64_t n_cells = 0;
    uint32_t* d_cell_start = nullptr; // n_cells + 1
    uint32_t* d_cursor = nullptr;     // n_cells
    uint32_t* d_sorted_rows = nullptr; // n rows
    uint64_t cap_rows = 0;
    uint64_t built_step = ~0ull;
    uint64_t n_indexed = 0;
    // lazy: the dense table is not rebuilt every step (huge / sparse grids whose consumers keep their own
    // index); the positions seen by the last PreUpdate are snapshotted and the table is built when queried
    bool lazy = false;
    void* d_snap_pos = nullptr;
    uint8_t* d_snap_flags = nullptr;
    uint64_t snap_n = 0;
    bool snap_valid = false, snap_built = false;
};

struct KernelStat
{
    const char* name;
    uint64_t launches = 0;
    double device_ms = 0.0;
    double algorithmic_bytes = 0.0;
};

struct Comm; // NCCL, comm.cpp

struct PendingTimer
{
    cudaEvent_t a, b;
    size_t stat;
};

struct Sim
{
    int device = 0;
    cudaStream_t stream = nullptr;
    uint64_t seed = 0x1CE27002025ull;
    uint32_t replica = 0;
    std::vector<ComponentInfo> comps;
    std::vector<Table> tables;
    std::vector<HostSpawn> spawns;
    std::vector<SystemInst> systems;
    std::vector<SeriesInst> series;
    std::vector<GridInst> grids;
    uint64_t step = 1; // StepNumber observed by the next user step (simulation_builder.rs:50)
    uint32_t part_rank = 0, part_world = 1;

    // scratch
    uint8_t* d_scratch = nullptr;
    size_t scratch_bytes = 0;
    uint32_t* d_ticket = nullptr; // zero-initialised; kernels leave it zero
    uint32_t* d_error = nullptr;  // device-raised error flags
    uint8_t* d_flush = nullptr;
    size_t flush_bytes = 0;
    void* h_pinned = nullptr;
    size_t pinned_bytes = 0;

    // instrumentation
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    bool run_pending = false;
    double last_ms = 0.0;
    uint64_t last_launches = 0, launches = 0;
    bool profiling = false;
    std::vector<KernelStat> kstats;
    std::vector<PendingTimer> pending_timers; // profiling event pairs of this handle, resolved when it synchronises
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;

    // workload state owned by modules (forest, pandemic, ...), see workloads.hpp
    struct Modules;
    std::unique_ptr<Modules> mods;
    Comm* comm = nullptr;

    Sim();
    ~Sim();
};

struct Builder
{
    std::unique_ptr<Sim> sim;
    bool device_set = false;
};

// ---- helpers implemented in engine.cu
int32_t ensure_scratch(Sim& s, size_t bytes);
// Device allocations of a handle's life come from a small process-wide cache of freed blocks (exact size match, per
// device): building simulation after simulation - the end-to-end job of bench.py, replica after replica - then costs no
// cudaMalloc / cudaFree (each a device synchronisation and a page-table update; their cost also grows with what the
// process freed before).  dev_release may only be called for memory no stream is using any more (the destroy path,
// after the handle's stream was synchronised); `cache` = false frees for real (memory exported through CUDA IPC).
cudaError_t dev_alloc_bytes(void** p, size_t bytes);
template <typename T>
inline cudaError_t dev_alloc(T** p, size_t bytes) { return dev_alloc_bytes(reinterpret_cast<void**>(p), bytes); }
void dev_release(void* p, bool cache = true);
cudaError_t dev_free(void* p); // cudaFree of a dev_alloc block that may still be in use (synchronises like cudaFree)
int32_t ensure_pinned(Sim& s, size_t bytes);
int table_for_components(Sim& s, const std::vector<int>& comps); // first table holding all data comps, -1 if none
struct RowFilter
{
    uint8_t need_set = 0, need_clear = 0;
    bool impossible = false; // a With<data comp> the table does not have
    bool trivial = false;    // no flags read needed
};
RowFilter row_filter(const Sim& s, const Table& t, const FilterSpec& f);
int32_t validate_filter(const Sim& s, const FilterSpec& f);
KernelStat& kstat(Sim& s, const char* name);
// live rows of a table in ascending uid order (the order every host-side reader reports entities in), with their uids
int32_t live_rows_by_uid(Sim& s, Table& t, std::vector<uint32_t>& rows, std::vector<uint32_t>& uids);

// scoped kernel timing (only when profiling is on)
// developer aid: with INCERTO_B200_TIMING=1 the host-side stages of build() report their wall time on stderr
void stage_mark(const char* what);

struct KTimer
{
    Sim& s;
    size_t stat = 0; // index into s.kstats
    cudaEvent_t a = nullptr, b = nullptr;
    KTimer(Sim& s_, const char* name, double bytes);
    ~KTimer();
};

} // namespace incerto
