// Engine core: handles, builder, spawners, the per-step program and run loop.
//
// Reference surface restated here (host side of the boundary):
//   SimulationBuilder::new/add_systems/add_entity_spawner/build   src/simulation_builder.rs:42-70,154-158,295-305
//   Simulation::run                                                src/simulation.rs:25-31
//   StepNumberPlugin                                               src/plugins/step_number.rs:6-23
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdarg>
#include <cstdio>

#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "kernels_grid.hpp"
#include "modules.hpp"

namespace incerto
{

// ------------------------------------------------------------------- errors
static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int32_t cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
    set_error("CUDA error %d (%s) in %s at %s:%d", static_cast<int>(e), cudaGetErrorString(e), what, file, line);
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? INCERTO_ERR_NO_DEVICE : INCERTO_ERR_CUDA;
}

// ------------------------------------------------------- device block cache
namespace
{
struct BlockCache
{
    std::mutex mu;
    std::map<std::pair<int, size_t>, std::vector<void*>> free_blocks; // (device, bytes) -> blocks
    std::map<void*, std::pair<int, size_t>> owner;                    // blocks handed out by dev_alloc_bytes
    size_t cached_bytes = 0;
    static constexpr size_t kMaxCached = 3ull << 30;
};
BlockCache& block_cache()
{
    static BlockCache* c = new BlockCache; // never destroyed: the CUDA context may be gone at exit
    return *c;
}
} // namespace

cudaError_t dev_alloc_bytes(void** p, size_t bytes)
{
    static const bool off = [] { const char* e = std::getenv("INCERTO_B200_NO_BLOCK_CACHE"); return e && e[0] == '1'; }();
    int dev = 0;
    cudaGetDevice(&dev);
    BlockCache& c = block_cache();
    if (!off)
    {
        std::lock_guard<std::mutex> lk(c.mu);
        auto it = c.free_blocks.find({dev, bytes});
        if (it != c.free_blocks.end() && !it->second.empty())
        {
            *p = it->second.back();
            it->second.pop_back();
            c.cached_bytes -= bytes;
            c.owner[*p] = {dev, bytes};
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess && !off)
    {
        // out of memory with blocks parked in the cache: give them back and try again
        std::lock_guard<std::mutex> lk(c.mu);
        for (auto& kv : c.free_blocks)
            if (kv.first.first == dev)
            {
                for (void* q : kv.second) cudaFree(q);
                c.cached_bytes -= kv.first.second * kv.second.size();
                kv.second.clear();
            }
        (void)cudaGetLastError();
        e = cudaMalloc(p, bytes);
    }
    if (e == cudaSuccess && !off)
    {
        std::lock_guard<std::mutex> lk(c.mu);
        c.owner[*p] = {dev, bytes};
    }
    return e;
}

void dev_release(void* p, bool cache)
{
    if (!p) return;
    BlockCache& c = block_cache();
    {
        std::lock_guard<std::mutex> lk(c.mu);
        auto it = c.owner.find(p);
        if (it != c.owner.end())
        {
            const std::pair<int, size_t> key = it->second;
            c.owner.erase(it);
            if (cache && c.cached_bytes + key.second <= BlockCache::kMaxCached)
            {
                c.free_blocks[key].push_back(p);
                c.cached_bytes += key.second;
                return;
            }
        }
    }
    cudaFree(p);
}

cudaError_t dev_free(void* p)
{
    if (!p) return cudaSuccess;
    {
        BlockCache& c = block_cache();
        std::lock_guard<std::mutex> lk(c.mu);
        c.owner.erase(p);
    }
    return cudaFree(p);
}

// ---------------------------------------------------------------------- Sim
Sim::Sim() : mods(new Modules) {}

Sim::~Sim()
{
    if (stream == nullptr && d_scratch == nullptr && tables.empty()) return;
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    forest_free(*this);
    pandemic_free(*this);
    air_free(*this);
    spatial_free(*this);
    comm_free(*this);
    for (auto& t : tables)
    {
        for (auto& c : t.cols)
        {
            dev_release(c.d);
            dev_release(c.d_alt);
        }
        dev_release(t.d_flags);
        dev_release(t.d_flags_alt);
        dev_release(t.d_uid);
    }
    for (auto& se : series)
    {
        dev_release(se.d_values);
        dev_release(se.d_counts);
    }
    for (auto& g : grids)
    {
        dev_release(g.d_cell_start);
        dev_release(g.d_cursor);
        dev_release(g.d_sorted_rows);
        dev_release(g.d_snap_pos);
        dev_release(g.d_snap_flags);
    }
    dev_release(d_scratch);
    dev_release(d_ticket);
    dev_release(d_error);
    dev_release(d_flush);
    dev_release(mods->d_step);
    if (h_pinned) cudaFreeHost(h_pinned);
    if (ev_begin) cudaEventDestroy(ev_begin);
    if (ev_end) cudaEventDestroy(ev_end);
    for (auto& p : ev_pool)
    {
        cudaEventDestroy(p.first);
        cudaEventDestroy(p.second);
    }
    if (stream) cudaStreamDestroy(stream);
}

int32_t ensure_scratch(Sim& s, size_t bytes)
{
    if (bytes <= s.scratch_bytes) return INCERTO_OK;
    if (s.d_scratch)
    {
        INC_CUDA(cudaStreamSynchronize(s.stream));
        INC_CUDA(dev_free(s.d_scratch));
        s.d_scratch = nullptr;
    }
    bytes = (bytes + 0xfffff) & ~static_cast<size_t>(0xfffff);
    INC_CUDA(dev_alloc(&s.d_scratch, bytes));
    s.scratch_bytes = bytes;
    return INCERTO_OK;
}

int32_t ensure_pinned(Sim& s, size_t bytes)
{
    if (bytes <= s.pinned_bytes) return INCERTO_OK;
    if (s.h_pinned)
    {
        INC_CUDA(cudaStreamSynchronize(s.stream));
        INC_CUDA(cudaFreeHost(s.h_pinned));
        s.h_pinned = nullptr;
    }
    bytes = (bytes + 0xffff) & ~static_cast<size_t>(0xffff);
    INC_CUDA(cudaMallocHost(&s.h_pinned, bytes));
    s.pinned_bytes = bytes;
    return INCERTO_OK;
}

void stage_mark(const char* what)
{
    static const bool on = std::getenv("INCERTO_B200_TIMING") != nullptr;
    if (!on) return;
    static auto last = std::chrono::steady_clock::now();
    const auto now = std::chrono::steady_clock::now();
    if (what) std::fprintf(stderr, "[incerto_b200 timing] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
}

KernelStat& kstat(Sim& s, const char* name)
{
    for (auto& k : s.kstats)
        if (k.name == name || std::strcmp(k.name, name) == 0) return k;
    KernelStat k;
    k.name = name;
    s.kstats.push_back(k);
    return s.kstats.back();
}

// Profiling timers wait in the handle that launched them (Sim::pending_timers) until that handle synchronises: two
// handles driven from one thread (replicas in flight) never see each other's events.
KTimer::KTimer(Sim& s_, const char* name, double bytes) : s(s_)
{
    KernelStat& st = kstat(s_, name);
    stat = static_cast<size_t>(&st - s.kstats.data()); // an index: kstats may grow while this timer is alive
    st.launches += 1;
    st.algorithmic_bytes += bytes;
    s.launches += 1;
    if (s.profiling)
    {
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess || cudaEventRecord(a, s.stream) != cudaSuccess)
        {
            if (a) cudaEventDestroy(a);
            if (b) cudaEventDestroy(b);
            a = b = nullptr;
            cudaGetLastError();
        }
    }
}
KTimer::~KTimer()
{
    if (a)
    {
        cudaEventRecord(b, s.stream);
        s.pending_timers.push_back(PendingTimer{a, b, stat});
    }
}

static void resolve_timers(Sim& s)
{
    for (auto& p : s.pending_timers)
    {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess && p.stat < s.kstats.size())
            s.kstats[p.stat].device_ms += ms;
        else
            cudaGetLastError();
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    s.pending_timers.clear();
}

int32_t live_rows_by_uid(Sim& s, Table& t, std::vector<uint32_t>& rows, std::vector<uint32_t>& uids)
{
    rows.clear();
    uids.clear();
    if (!t.n || !t.d_flags) return INCERTO_OK;
    std::vector<uint8_t> fl(t.n);
    std::vector<uint32_t> uid;
    INC_CUDA(cudaMemcpyAsync(fl.data(), t.d_flags, t.n, cudaMemcpyDeviceToHost, s.stream));
    if (t.d_uid)
    {
        uid.resize(t.n);
        INC_CUDA(cudaMemcpyAsync(uid.data(), t.d_uid, t.n * 4, cudaMemcpyDeviceToHost, s.stream));
    }
    INC_CUDA(cudaStreamSynchronize(s.stream));
    for (uint64_t r = 0; r < t.n; ++r)
        if (fl[r] & kAliveBit) rows.push_back(static_cast<uint32_t>(r));
    if (t.d_uid)
        std::sort(rows.begin(), rows.end(), [&](uint32_t a, uint32_t b) { return uid[a] < uid[b]; });
    uids.reserve(rows.size());
    for (uint32_t r : rows) uids.push_back(t.d_uid ? uid[r] : r);
    return INCERTO_OK;
}

int table_for_components(Sim& s, const std::vector<int>& comps)
{
    for (size_t ti = 0; ti < s.tables.size(); ++ti)
    {
        bool ok = true;
        for (int c : comps)
        {
            if (c < 0 || c >= static_cast<int>(s.comps.size())) return -1;
            if (s.comps[c].dtype == INCERTO_MARKER) continue;
            if (!s.tables[ti].has(c)) ok = false;
        }
        if (ok) return static_cast<int>(ti);
    }
    return -1;
}

int32_t validate_filter(const Sim& s, const FilterSpec& f)
{
    for (int c : f.with)
    {
        if (c < 0 || c >= static_cast<int>(s.comps.size()))
        {
            set_error("filter names an unregistered component handle %d", c);
            return INCERTO_ERR_INVALID_ARGUMENT;
        }
        if (!s.comps[c].exists) return INCERTO_ERR_COMPONENT_DOES_NOT_EXIST;
    }
    for (int c : f.without)
    {
        if (c < 0 || c >= static_cast<int>(s.comps.size()))
        {
            set_error("filter names an unregistered component handle %d", c);
            return INCERTO_ERR_INVALID_ARGUMENT;
        }
        if (!s.comps[c].exists) return INCERTO_ERR_COMPONENT_DOES_NOT_EXIST;
    }
    return INCERTO_OK;
}

RowFilter row_filter(const Sim& s, const Table& t, const FilterSpec& f)
{
    RowFilter r;
    r.need_set = kAliveBit;
    for (int c : f.with)
    {
        const ComponentInfo& ci = s.comps[c];
        if (ci.dtype == INCERTO_MARKER)
            r.need_set |= static_cast<uint8_t>(1u << ci.marker_bit);
        else if (!t.has(c))
            r.impossible = true;
    }
    for (int c : f.without)
    {
        const ComponentInfo& ci = s.comps[c];
        if (ci.dtype == INCERTO_MARKER)
            r.need_clear |= static_cast<uint8_t>(1u << ci.marker_bit);
        else if (t.has(c))
            r.impossible = true;
    }
    r.trivial = (r.need_set == kAliveBit) && r.need_clear == 0 && t.all_alive;
    return r;
}

// ------------------------------------------------------------------ spawning
static std::vector<int> signature_of(const Sim& s, const HostSpawn& sp)
{
    std::vector<int> sig;
    for (auto& c : sp.cols)
        if (s.comps[c.comp].dtype != INCERTO_MARKER) sig.push_back(c.comp);
    std::sort(sig.begin(), sig.end());
    return sig;
}

static int find_or_add_table(Sim& s, const std::vector<int>& sig)
{
    for (size_t i = 0; i < s.tables.size(); ++i)
        if (s.tables[i].signature == sig) return static_cast<int>(i);
    Table t;
    t.signature = sig;
    for (int c : sig)
    {
        Column col;
        col.comp = c;
        col.row_bytes = s.comps[c].row_bytes();
        t.cols.push_back(col);
    }
    s.tables.push_back(std::move(t));
    return static_cast<int>(s.tables.size()) - 1;
}

// Device spawners that create generic tables (coin tossers ...) are expanded here.
static int32_t device_spawn_signature(Sim& s, const HostSpawn& sp, std::vector<int>& sig, uint64_t& n)
{
    switch (sp.kind)
    {
    case INCERTO_SPAWN_COIN_TOSSERS:
    {
        if (sp.params.size() != sizeof(incerto_spawn_coin_tossers_params)) return INCERTO_ERR_INVALID_ARGUMENT;
        incerto_spawn_coin_tossers_params p;
        std::memcpy(&p, sp.params.data(), sizeof(p));
        sig = {p.num_heads, p.num_tosses};
        std::sort(sig.begin(), sig.end());
        n = p.num_entities;
        return INCERTO_OK;
    }
    case INCERTO_SPAWN_FOREST_GRID:
    {
        if (sp.params.size() != sizeof(incerto_spawn_forest_params)) return INCERTO_ERR_INVALID_ARGUMENT;
        incerto_spawn_forest_params p;
        std::memcpy(&p, sp.params.data(), sizeof(p));
        sig = {p.position, p.cell};
        std::sort(sig.begin(), sig.end());
        n = static_cast<uint64_t>(p.width) * p.height + p.initial_fire_count;
        return INCERTO_OK;
    }
    case INCERTO_SPAWN_PANDEMIC_PEOPLE:
    {
        if (sp.params.size() != sizeof(incerto_spawn_pandemic_params)) return INCERTO_ERR_INVALID_ARGUMENT;
        incerto_spawn_pandemic_params p;
        std::memcpy(&p, sp.params.data(), sizeof(p));
        auto bad = [&](int c) { return c < 0 || c >= static_cast<int>(s.comps.size()); };
        if (bad(p.person) || bad(p.position) || bad(p.infected)) return INCERTO_ERR_INVALID_ARGUMENT;
        s.comps[p.person].exists = true;
        s.comps[p.infected].exists = true;
        sig = {p.position};
        n = p.initial_population;
        return INCERTO_OK;
    }
    case INCERTO_SPAWN_SPATIAL_POPULATION:
    {
        if (sp.params.size() != sizeof(incerto_spawn_spatial_params)) return INCERTO_ERR_INVALID_ARGUMENT;
        incerto_spawn_spatial_params p;
        std::memcpy(&p, sp.params.data(), sizeof(p));
        auto bad = [&](int c) { return c < 0 || c >= static_cast<int>(s.comps.size()); };
        if (bad(p.position) || bad(p.disease_state) || bad(p.social_distancing) || bad(p.quarantined) || bad(p.contact_history))
            return INCERTO_ERR_INVALID_ARGUMENT;
        s.comps[p.social_distancing].exists = true;
        sig = {p.position, p.disease_state, p.quarantined, p.contact_history};
        std::sort(sig.begin(), sig.end());
        n = p.initial_population;
        return INCERTO_OK;
    }
    case INCERTO_SPAWN_TRADERS:
    {
        if (sp.params.size() != sizeof(incerto_spawn_traders_params)) return INCERTO_ERR_INVALID_ARGUMENT;
        incerto_spawn_traders_params p;
        std::memcpy(&p, sp.params.data(), sizeof(p));
        auto bad = [&](int c) { return c < 0 || c >= static_cast<int>(s.comps.size()); };
        if (bad(p.trader_id) || bad(p.cash) || bad(p.portfolio) || bad(p.net_worth)) return INCERTO_ERR_INVALID_ARGUMENT;
        sig = {p.trader_id, p.cash, p.portfolio, p.net_worth};
        std::sort(sig.begin(), sig.end());
        n = p.num_traders;
        return INCERTO_OK;
    }
    case INCERTO_SPAWN_STOCKS:
    {
        if (sp.params.size() != sizeof(incerto_spawn_stocks_params)) return INCERTO_ERR_INVALID_ARGUMENT;
        incerto_spawn_stocks_params p;
        std::memcpy(&p, sp.params.data(), sizeof(p));
        auto bad = [&](int c) { return c < 0 || c >= static_cast<int>(s.comps.size()); };
        if (bad(p.stock_id) || bad(p.stock_price) || bad(p.stock_delisted)) return INCERTO_ERR_INVALID_ARGUMENT;
        sig = {p.stock_id, p.stock_price};
        std::sort(sig.begin(), sig.end());
        n = p.num_stocks;
        return INCERTO_OK;
    }
    case INCERTO_SPAWN_AIRCRAFT:
    {
        if (sp.params.size() != sizeof(incerto_spawn_aircraft_params)) return INCERTO_ERR_INVALID_ARGUMENT;
        incerto_spawn_aircraft_params p;
        std::memcpy(&p, sp.params.data(), sizeof(p));
        auto bad = [&](int c) { return c < 0 || c >= static_cast<int>(s.comps.size()); };
        if (bad(p.position) || bad(p.target_altitude)) return INCERTO_ERR_INVALID_ARGUMENT;
        sig = {p.position, p.target_altitude};
        std::sort(sig.begin(), sig.end());
        n = p.num_aircraft;
        return INCERTO_OK;
    }
    default:
        set_error("device spawner kind %u is not implemented", sp.kind);
        return INCERTO_ERR_UNSUPPORTED;
    }
}

static int32_t run_spawners(Sim& s, bool first_time)
{
    // pass 1: rows per table
    std::vector<uint64_t> rows(s.tables.size(), 0);
    std::vector<int> spawn_table(s.spawns.size(), -1);
    for (size_t i = 0; i < s.spawns.size(); ++i)
    {
        const HostSpawn& sp = s.spawns[i];
        std::vector<int> sig;
        uint64_t n = sp.n;
        if (sp.device_kind)
            INC_TRY(device_spawn_signature(s, sp, sig, n));
        else
            sig = signature_of(s, sp);
        for (int c : sig)
        {
            if (c < 0 || c >= static_cast<int>(s.comps.size()))
            {
                set_error("spawner names an unregistered component handle %d", c);
                return INCERTO_ERR_INVALID_ARGUMENT;
            }
            s.comps[c].exists = true;
        }
        for (auto& c : sp.cols) s.comps[c.comp].exists = true;
        const int ti = find_or_add_table(s, sig);
        if (static_cast<size_t>(ti) >= rows.size()) rows.resize(ti + 1, 0);
        rows[ti] += n;
        spawn_table[i] = ti;
    }
    INC_TRY(traders_layout(s, rows)); // the managed portfolio's width depends on the number of stocks
    INC_TRY(spatial_layout(s));
    // allocate
    for (size_t ti = 0; ti < s.tables.size(); ++ti)
    {
        Table& t = s.tables[ti];
        const bool dense_device =
            std::any_of(s.spawns.begin(), s.spawns.end(), [&](const HostSpawn& sp) {
                return sp.device_kind && sp.kind == INCERTO_SPAWN_FOREST_GRID;
            }) &&
            s.mods->forest.table == static_cast<int>(ti);
        if (first_time)
        {
            t.n = rows[ti];
            t.n_spawned = rows[ti];
            if (rows[ti] >= (1ull << 32))
            {
                set_error("a table holds at most 2^32-1 entities per replica (uid is 32 bit)");
                return INCERTO_ERR_INVALID_ARGUMENT;
            }
        }
        else if (t.n_spawned != rows[ti])
        {
            set_error("reset: spawner row count changed");
            return INCERTO_ERR_INVALID_ARGUMENT;
        }
        (void)dense_device;
    }
    // tables the forest module adopts from the host spawn data never need their generic columns on the device
    int forest_pos = -1, forest_cell = -1;
    for (auto& sys : s.systems)
        if (sys.kind >= INCERTO_SYS_FOREST_SPREAD && sys.kind <= INCERTO_SYS_FOREST_REGROWTH &&
            sys.params.size() == sizeof(incerto_forest_params))
        {
            incerto_forest_params fp;
            std::memcpy(&fp, sys.params.data(), sizeof(fp));
            forest_pos = fp.position;
            forest_cell = fp.cell;
        }
    // pass 2: fill
    std::vector<uint64_t> cursor(s.tables.size(), 0);
    for (size_t i = 0; i < s.spawns.size(); ++i)
    {
        const HostSpawn& sp = s.spawns[i];
        const int ti = spawn_table[i];
        Table& t = s.tables[ti];
        if (sp.device_kind && sp.kind == INCERTO_SPAWN_FOREST_GRID)
        {
            if (cursor[ti] != 0)
            {
                set_error("the forest grid spawner must be the first spawner of its archetype");
                return INCERTO_ERR_INVALID_ARGUMENT;
            }
            s.mods->forest.table = ti;
            INC_TRY(forest_spawn_device(s, sp));
            std::vector<int> sig;
            uint64_t n = 0;
            INC_TRY(device_spawn_signature(s, sp, sig, n));
            cursor[ti] += n;
            continue;
        }
        if (!sp.device_kind && forest_pos >= 0 && t.signature.size() == 2 && t.has(forest_pos) && t.has(forest_cell))
        {
            cursor[ti] += sp.n; // adopted by forest_bind / forest_reset_after_spawn straight from the host buffers
            continue;
        }
        // generic columns
        if (first_time || t.d_flags == nullptr)
        {
            if (t.d_flags == nullptr && t.n)
            {
                INC_CUDA(dev_alloc(&t.d_flags, t.n));
                for (auto& c : t.cols)
                    if (c.d == nullptr) INC_CUDA(dev_alloc(&c.d, t.n * c.row_bytes));
            }
        }
        // engine-managed columns start empty (Trader.portfolio: HashMap::new(), traders.rs:123)
        for (auto& c : t.cols)
            if (s.comps[c.comp].managed_bytes && cursor[ti] == 0 && c.d)
                INC_CUDA(cudaMemsetAsync(c.d, 0, t.n_spawned * c.row_bytes, s.stream));
        uint64_t n = sp.n;
        const uint64_t off = cursor[ti];
        if (sp.device_kind)
        {
            std::vector<int> sig;
            INC_TRY(device_spawn_signature(s, sp, sig, n));
            if (sp.kind == INCERTO_SPAWN_PANDEMIC_PEOPLE)
            {
                INC_TRY(pandemic_spawn_device(s, sp, t, off));
                cursor[ti] += n;
                continue;
            }
            if (sp.kind == INCERTO_SPAWN_SPATIAL_POPULATION)
            {
                INC_TRY(spatial_spawn_device(s, sp, t, off));
                cursor[ti] += n;
                continue;
            }
            if (sp.kind == INCERTO_SPAWN_TRADERS || sp.kind == INCERTO_SPAWN_STOCKS)
            {
                INC_TRY(traders_spawn_device(s, sp, t, off));
                cursor[ti] += n;
                continue;
            }
            if (sp.kind == INCERTO_SPAWN_AIRCRAFT)
            {
                INC_TRY(air_spawn_device(s, sp, t, off));
                cursor[ti] += n;
                continue;
            }
            // coin tossers: CoinTosser::default() (coin_toss.rs:60-62) -> zeros
            for (auto& c : t.cols) INC_CUDA(cudaMemsetAsync(static_cast<uint8_t*>(c.d) + off * c.row_bytes, 0, n * c.row_bytes, s.stream));
            INC_CUDA(cudaMemsetAsync(t.d_flags + off, kAliveBit, n, s.stream));
            cursor[ti] += n;
            continue;
        }
        // flags: alive + markers
        std::vector<uint8_t> flags(n, kAliveBit);
        for (auto& c : sp.cols)
        {
            const ComponentInfo& ci = s.comps[c.comp];
            if (ci.dtype != INCERTO_MARKER) continue;
            const uint8_t bit = static_cast<uint8_t>(1u << ci.marker_bit);
            if (!c.has_data)
                for (auto& f : flags) f |= bit;
            else
                for (uint64_t r = 0; r < n; ++r)
                    if (c.bytes()[r]) flags[r] |= bit;
        }
        INC_CUDA(cudaMemcpyAsync(t.d_flags + off, flags.data(), n, cudaMemcpyHostToDevice, s.stream));
        for (auto& c : sp.cols)
        {
            const ComponentInfo& ci = s.comps[c.comp];
            if (ci.dtype == INCERTO_MARKER) continue;
            Column* col = t.col(c.comp);
            const size_t es = dtype_size(ci.dtype);
            if (ci.managed_bytes) continue; // zeroed above; host values are not accepted for managed layouts
            if (!c.has_data)
            {
                // no values: Default::default()
                INC_CUDA(cudaMemsetAsync(static_cast<uint8_t*>(col->d) + off * col->row_bytes, 0, n * col->row_bytes, s.stream));
                continue;
            }
            if (ci.lanes == 3)
            {
                // 3 lanes are padded to a stride of 4 on the device
                std::vector<uint8_t> padded(n * 4 * es, 0);
                for (uint64_t r = 0; r < n; ++r) std::memcpy(&padded[r * 4 * es], c.bytes() + r * 3 * es, 3 * es);
                INC_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(col->d) + off * col->row_bytes, padded.data(),
                                         padded.size(), cudaMemcpyHostToDevice, s.stream));
                INC_CUDA(cudaStreamSynchronize(s.stream));
            }
            else
            {
                INC_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(col->d) + off * col->row_bytes, c.bytes(),
                                         n * col->row_bytes, cudaMemcpyHostToDevice, s.stream));
            }
        }
        INC_CUDA(cudaStreamSynchronize(s.stream)); // `flags` is a stack-owned staging buffer
        cursor[ti] += n;
    }
    for (auto& t : s.tables)
    {
        t.all_alive = true;
        t.n = t.n_spawned;
        if (t.d_uid)
        {
            INC_CUDA(dev_free(t.d_uid));
            t.d_uid = nullptr;
        }
    }
    return INCERTO_OK;
}

// ------------------------------------------------------------ row order
// Store the rows of a table in cell order (counting sort on the linearised cell key of `pos_comp`): entities that
// are close in space become close in memory, so that the per-cell tables the spatial workloads scatter into and
// gather from are touched sector by sector instead of at random.  Entities keep their uid (= spawn ordinal): every
// draw is keyed by it, so results do not depend on the row order.  Columns in `zero_comps` are known to be all
// zero (freshly spawned engine-managed containers) and are left alone.
int32_t table_sort_by_cell(Sim& s, int ti, int pos_comp, const GridDesc& g, const std::vector<int>& zero_comps)
{
    if (const char* e = std::getenv("INCERTO_B200_NO_ROW_SORT"))
        if (e[0] == '1') return INCERTO_OK;
    Table& t = s.tables[ti];
    if (t.n < 2 || t.n != t.n_spawned || !t.all_alive) return INCERTO_OK;
    Column* pc = t.col(pos_comp);
    if (!pc) return INCERTO_OK;
    const uint64_t n_cells = static_cast<uint64_t>(g.ext[0]) * g.ext[1] * g.ext[2];
    if (n_cells == 0 || n_cells >= (1ull << 32)) return INCERTO_OK;
    uint32_t *d_start = nullptr, *d_cursor = nullptr, *d_perm = nullptr, *d_scan = nullptr;
    auto cleanup = [&]() {
        dev_free(d_start);
        dev_free(d_cursor);
        dev_free(d_scan);
        dev_free(d_perm);
    };
    if (dev_alloc(&d_start, (n_cells + 1) * 4) != cudaSuccess || dev_alloc(&d_cursor, n_cells * 4) != cudaSuccess ||
        dev_alloc(&d_perm, t.n * 4) != cudaSuccess || dev_alloc(&d_scan, scan_scratch_bytes(n_cells) + 256) != cudaSuccess)
    {
        cudaGetLastError();
        cleanup();
        return INCERTO_OK; // not enough memory for the temporary index: keep the spawn order
    }
    // rows outside the bounds raise the grid's error flag: remember it, sort nothing (the first run reports it)
    launch_grid_build(g, static_cast<const int16_t*>(pc->d), s.comps[pos_comp].stride(), nullptr, 0, t.n, d_start, d_cursor,
                      d_perm, d_scan, s.d_error, s.stream);
    uint32_t total = 0;
    cudaError_t ce = cudaMemcpyAsync(&total, d_start + n_cells, 4, cudaMemcpyDeviceToHost, s.stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s.stream);
    if (ce != cudaSuccess || total != t.n)
    {
        cudaMemsetAsync(s.d_error, 0, 64, s.stream); // the workload's own index kernel raises it again, at run time
        cleanup();
        return ce == cudaSuccess ? INCERTO_OK : cuda_fail(ce, "table_sort_by_cell", __FILE__, __LINE__);
    }
    KTimer kt(s, "table_sort_by_cell", 0.0);
    auto permute = [&](void*& buf, size_t row_bytes) -> int32_t {
        void* nb = nullptr;
        INC_CUDA(dev_alloc(&nb, t.n * row_bytes));
        launch_gather_rows(buf, nb, static_cast<uint32_t>(row_bytes), d_perm, t.n, s.stream);
        INC_CUDA(cudaStreamSynchronize(s.stream));
        INC_CUDA(dev_free(buf));
        buf = nb;
        return INCERTO_OK;
    };
    int32_t st = INCERTO_OK;
    for (auto& c : t.cols)
    {
        if (std::find(zero_comps.begin(), zero_comps.end(), c.comp) != zero_comps.end()) continue;
        if (s.comps[c.comp].managed_bytes)
        {
            st = INCERTO_ERR_UNSUPPORTED;
            set_error("table_sort_by_cell: cannot reorder a populated engine-managed column");
            break;
        }
        if ((st = permute(c.d, c.row_bytes)) != INCERTO_OK) break;
    }
    if (st == INCERTO_OK)
    {
        void* f = t.d_flags;
        st = permute(f, 1);
        t.d_flags = static_cast<uint8_t*>(f);
    }
    if (st == INCERTO_OK)
    {
        if (t.d_uid)
        {
            void* u = t.d_uid;
            st = permute(u, 4);
            t.d_uid = static_cast<uint32_t*>(u);
        }
        else
        {
            t.d_uid = d_perm; // uid of row i = its spawn-order row
            d_perm = nullptr;
        }
    }
    cleanup();
    return st;
}

// ------------------------------------------------------------ step program
static int32_t bind_generic_systems(Sim& s)
{
    auto& prog = s.mods->program;
    std::vector<SystemInst> sorted = s.systems;
    std::stable_sort(sorted.begin(), sorted.end(),
                     [](const SystemInst& a, const SystemInst& b) { return a.kind < b.kind; });
    bool forest_added = false;
    for (const SystemInst& sys : sorted)
    {
        switch (sys.kind)
        {
        case INCERTO_SYS_ADD_CONST:
        {
            if (sys.params.size() != sizeof(incerto_add_const_params)) return INCERTO_ERR_INVALID_ARGUMENT;
            incerto_add_const_params p;
            std::memcpy(&p, sys.params.data(), sizeof(p));
            if (p.target < 0 || p.target >= static_cast<int>(s.comps.size()) || p.n_with > 4)
                return INCERTO_ERR_INVALID_ARGUMENT;
            s.comps[p.target].exists = true;
            FilterSpec f;
            for (uint32_t i = 0; i < p.n_with; ++i)
            {
                if (p.with[i] < 0 || p.with[i] >= static_cast<int>(s.comps.size())) return INCERTO_ERR_INVALID_ARGUMENT;
                f.with.push_back(p.with[i]);
                s.comps[p.with[i]].exists = true;
            }
            StepOp op;
            op.name = "add_const";
            op.fn = [p, f](Sim& sim, uint64_t, uint32_t nsteps) -> int32_t {
                for (auto& t : sim.tables)
                {
                    Column* c = t.col(p.target);
                    if (!c || !t.n) continue;
                    RowFilter rf = row_filter(sim, t, f);
                    if (rf.impossible) continue;
                    const ComponentInfo& ci = sim.comps[p.target];
                    KTimer kt(sim, "add_const", 2.0 * t.n * dtype_size(ci.dtype));
                    launch_add_const(c->d, ci.dtype, ci.stride(), rf.trivial ? nullptr : t.d_flags, rf.need_set, t.n,
                                     p.amount * static_cast<int64_t>(nsteps), sim.stream);
                }
                return INCERTO_OK;
            };
            op.multi_step_ok = true;
            prog.push_back(op);
            break;
        }
        case INCERTO_SYS_ADD_COLUMN:
        {
            if (sys.params.size() != sizeof(incerto_add_column_params)) return INCERTO_ERR_INVALID_ARGUMENT;
            incerto_add_column_params p;
            std::memcpy(&p, sys.params.data(), sizeof(p));
            if (p.target < 0 || p.source < 0 || p.target >= static_cast<int>(s.comps.size()) ||
                p.source >= static_cast<int>(s.comps.size()))
                return INCERTO_ERR_INVALID_ARGUMENT;
            s.comps[p.target].exists = true;
            s.comps[p.source].exists = true;
            StepOp op;
            op.name = "add_column";
            op.fn = [p](Sim& sim, uint64_t, uint32_t) -> int32_t {
                for (auto& t : sim.tables)
                {
                    Column* d = t.col(p.target);
                    Column* sc = t.col(p.source);
                    if (!d || !sc || !t.n) continue;
                    KTimer kt(sim, "add_column", 3.0 * t.n * dtype_size(sim.comps[p.target].dtype));
                    launch_add_column(d->d, sim.comps[p.target].dtype, sc->d, sim.comps[p.source].dtype,
                                      t.all_alive ? nullptr : t.d_flags, t.n, sim.stream);
                }
                return INCERTO_OK;
            };
            prog.push_back(op);
            break;
        }
        case INCERTO_SYS_COIN_TOSS:
        {
            if (sys.params.size() != sizeof(incerto_coin_toss_params)) return INCERTO_ERR_INVALID_ARGUMENT;
            incerto_coin_toss_params p;
            std::memcpy(&p, sys.params.data(), sizeof(p));
            if (p.num_heads < 0 || p.num_tosses < 0 || p.num_heads >= static_cast<int>(s.comps.size()) ||
                p.num_tosses >= static_cast<int>(s.comps.size()))
                return INCERTO_ERR_INVALID_ARGUMENT;
            const int dt = s.comps[p.num_heads].dtype;
            if ((dt != INCERTO_U32 && dt != INCERTO_U64) || s.comps[p.num_tosses].dtype != dt ||
                s.comps[p.num_heads].lanes != 1 || s.comps[p.num_tosses].lanes != 1)
            {
                set_error("coin_toss: num_heads/num_tosses must both be U32 or both U64 scalar columns");
                return INCERTO_ERR_INVALID_ARGUMENT;
            }
            s.comps[p.num_heads].exists = true;
            s.comps[p.num_tosses].exists = true;
            s.mods->coin.active = true;
            s.mods->coin.comp_heads = p.num_heads;
            s.mods->coin.comp_tosses = p.num_tosses;
            s.mods->coin.odds = p.odds_heads;
            StepOp op;
            op.name = "coin_toss";
            op.multi_step_ok = true;
            op.fn = [p, dt](Sim& sim, uint64_t first_step, uint32_t nsteps) -> int32_t {
                const PhiloxKey key = make_key(sim.seed, sim.replica, TAG_COIN_TOSS);
                const uint64_t thr = bernoulli_threshold(p.odds_heads);
                for (auto& t : sim.tables)
                {
                    Column* h = t.col(p.num_heads);
                    Column* c = t.col(p.num_tosses);
                    if (!h || !c || !t.n) continue;
                    const uint8_t* flags = t.all_alive ? nullptr : t.d_flags;
                    // 16 B per entity-update; a multi-step launch still moves 16 B per entity
                    KTimer kt(sim, nsteps == 1 ? "coin_toss_step" : "coin_toss_multistep",
                              (dt == INCERTO_U32 ? 16.0 : 32.0) * t.n);
                    if (dt == INCERTO_U32)
                        launch_coin_toss(static_cast<uint32_t*>(h->d), static_cast<uint32_t*>(c->d), t.d_uid, flags,
                                         t.n, key, thr, static_cast<uint32_t>(first_step), nsteps, sim.stream);
                    else
                        launch_coin_toss_u64(static_cast<uint64_t*>(h->d), static_cast<uint64_t*>(c->d), t.d_uid,
                                             flags, t.n, key, thr, static_cast<uint32_t>(first_step), nsteps,
                                             sim.stream);
                }
                return INCERTO_OK;
            };
            prog.push_back(op);
            break;
        }
        case INCERTO_SYS_FOREST_SPREAD:
        case INCERTO_SYS_FOREST_BURN_PROGRESSION:
        case INCERTO_SYS_FOREST_REGROWTH:
        {
            if (!forest_added)
            {
                INC_TRY(forest_bind(s));
                StepOp op;
                op.name = "forest_step";
                op.bumps_device_step = true;
                op.whole_run_ok = true;
                op.prepare = [](Sim& sim, uint64_t nsteps) { return forest_prepare(sim, nsteps); };
                op.fn = [](Sim& sim, uint64_t first_step, uint32_t nsteps) { return forest_step(sim, first_step, nsteps); };
                prog.push_back(op);
                forest_added = true;
            }
            break;
        }
        case INCERTO_SYS_PANDEMIC_MOVE:
        case INCERTO_SYS_PANDEMIC_INFECT:
        case INCERTO_SYS_PANDEMIC_DIE:
        case INCERTO_SYS_PANDEMIC_RECOVER:
        {
            if (!s.mods->pandemic.active)
            {
                INC_TRY(pandemic_bind(s));
                StepOp op;
                op.name = "pandemic_step";
                op.fn = [](Sim& sim, uint64_t first_step, uint32_t nsteps) { return pandemic_step(sim, first_step, nsteps); };
                prog.push_back(op);
            }
            break;
        }
        case INCERTO_SYS_SPATIAL_INCUBATION:
        case INCERTO_SYS_SPATIAL_TRANSMISSION:
        case INCERTO_SYS_SPATIAL_RECOVERY_DEATH:
        case INCERTO_SYS_SPATIAL_CONTACT_HISTORY:
        case INCERTO_SYS_SPATIAL_CONTACT_TRACING:
        case INCERTO_SYS_SPATIAL_QUARANTINE:
        case INCERTO_SYS_SPATIAL_MOVE:
        {
            if (!s.mods->spatial.active)
            {
                INC_TRY(spatial_bind(s));
                StepOp op;
                op.name = "spatial_step";
                op.fn = [](Sim& sim, uint64_t first_step, uint32_t nsteps) { return spatial_step(sim, first_step, nsteps); };
                prog.push_back(op);
            }
            break;
        }
        case INCERTO_SYS_TRADERS_BUY:
        case INCERTO_SYS_TRADERS_SELL:
        case INCERTO_SYS_TRADERS_NET_WORTH:
        case INCERTO_SYS_STOCKS_PRICE_CHANGE:
        {
            if (!s.mods->traders.active)
            {
                INC_TRY(traders_bind(s));
                StepOp op;
                op.name = "traders_step";
                op.fn = [](Sim& sim, uint64_t first_step, uint32_t nsteps) { return traders_step(sim, first_step, nsteps); };
                prog.push_back(op);
            }
            break;
        }
        case INCERTO_SYS_AIR_CHECK_CONFLICTS:
        case INCERTO_SYS_AIR_MOVE:
        {
            if (!s.mods->air.active)
            {
                INC_TRY(air_bind(s));
                StepOp op;
                op.name = "air_step";
                op.fn = [](Sim& sim, uint64_t first_step, uint32_t nsteps) { return air_step(sim, first_step, nsteps); };
                prog.push_back(op);
            }
            break;
        }
        default:
            set_error("system kind %u is not implemented in this build", sys.kind);
            return INCERTO_ERR_UNSUPPORTED;
        }
    }
    return INCERTO_OK;
}

// implemented in sampling.cu
int32_t series_prepare(Sim& s, uint64_t first_step, uint64_t nsteps);
int32_t series_sample_step(Sim& s, uint64_t step);
void series_after_run(Sim& s);
void series_clear(Sim& s);
int32_t grids_allocate(Sim& s);
int32_t grids_rebuild_for_step(Sim& s, uint64_t step, bool last_of_run);

static int32_t do_run_enqueue(Sim& s, uint64_t num_steps);

// StepNumber and the Philox counter word that carries it are 32 bit on the device: a run that would take the step
// counter past 2^32 - 1 is refused instead of silently repeating its draws
static int32_t check_step_range(const Sim& s, uint64_t num_steps)
{
    if (num_steps > 0xffffffffull || s.step + num_steps > 0xffffffffull)
    {
        set_error("run: StepNumber would exceed 2^32 - 1 (step %llu + %llu); the draw counter is 32 bit",
                  static_cast<unsigned long long>(s.step), static_cast<unsigned long long>(num_steps));
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    return INCERTO_OK;
}

static int32_t do_run(Sim& s, uint64_t num_steps)
{
    INC_TRY(check_step_range(s, num_steps));
    INC_CUDA(cudaSetDevice(s.device));
    if (s.run_pending) INC_TRY(incerto_sim_synchronize(reinterpret_cast<incerto_sim*>(&s)));
    s.last_launches = s.launches;
    INC_TRY(series_prepare(s, s.step, num_steps));
    for (auto& op : s.mods->program)
        if (op.prepare) INC_TRY(op.prepare(s, num_steps));
    INC_CUDA(cudaEventRecord(s.ev_begin, s.stream));
    INC_TRY(do_run_enqueue(s, num_steps));
    INC_TRY(forest_finalize(s)); // the last step's FireStats record (every earlier one was written by its successor)
    INC_CUDA(cudaEventRecord(s.ev_end, s.stream));
    s.run_pending = true;
    series_after_run(s);
    return INCERTO_OK;
}

// the steps themselves, enqueued on the engine's stream
static int32_t do_run_enqueue(Sim& s, uint64_t num_steps)
{
    auto& prog = s.mods->program;
    bool all_multi = !prog.empty();
    bool device_bump = false;
    for (auto& op : prog)
    {
        all_multi = all_multi && op.multi_step_ok;
        device_bump = device_bump || op.bumps_device_step;
    }
    uint64_t remaining = num_steps;
    // one op that owns the whole step (and samples its own fused series): hand it the run in one go,
    // so that it can replay a captured graph of steps
    bool whole = prog.size() == 1 && prog[0].whole_run_ok && remaining > 0;
    for (size_t i = 0; i < s.series.size() && whole; ++i) whole = static_cast<int>(i) == s.mods->forest.series;
    for (auto& g : s.grids) whole = whole && (g.d_cell_start == nullptr);
    if (whole)
    {
        uint64_t left = remaining;
        while (left)
        {
            const uint32_t k = static_cast<uint32_t>(std::min<uint64_t>(left, 1u << 20));
            INC_TRY(prog[0].fn(s, s.step, k));
            s.step += k;
            left -= k;
        }
        for (auto& se : s.series) se.n_slots = std::max<uint64_t>(se.n_slots, std::min<uint64_t>((s.step - 1) / se.interval, se.capacity));
        remaining = 0;
    }
    while (remaining)
    {
        uint64_t k = 1;
        if (all_multi)
        {
            k = std::min<uint64_t>(remaining, 1u << 20);
            for (auto& se : s.series)
            {
                const uint64_t next = ((s.step + se.interval - 1) / se.interval) * se.interval; // next sampling step
                k = std::min<uint64_t>(k, next - s.step + 1);
            }
        }
        INC_TRY(grids_rebuild_for_step(s, s.step, remaining == k));
        for (auto& op : prog) INC_TRY(op.fn(s, s.step, static_cast<uint32_t>(k)));
        s.step += k - 1;
        INC_TRY(series_sample_step(s, s.step)); // PostUpdate (time_series.rs:100,175-178)
        s.step += 1;                            // Last (step_number.rs:16)
        if (!device_bump)
        {
            // keep the device mirror in lock step
            const uint32_t v = static_cast<uint32_t>(s.step);
            INC_CUDA(cudaMemcpyAsync(s.mods->d_step, &v, sizeof(v), cudaMemcpyHostToDevice, s.stream));
        }
        remaining -= k;
    }
    return INCERTO_OK;
}

} // namespace incerto

using namespace incerto;

// =========================================================================
//                                  C ABI
// =========================================================================
extern "C"
{

const char* incerto_last_error_message(void) { return g_error; }
uint32_t incerto_abi_version(void) { return INCERTO_B200_ABI_VERSION; }

int32_t incerto_builder_new(incerto_builder** out)
{
    if (!out) return INCERTO_ERR_INVALID_ARGUMENT;
    Builder* b = new Builder;
    b->sim.reset(new Sim);
    *out = reinterpret_cast<incerto_builder*>(b);
    return INCERTO_OK;
}

void incerto_builder_free(incerto_builder* b) { delete reinterpret_cast<Builder*>(b); }

#define INC_BUILDER(b)                                    \
    Builder* B = reinterpret_cast<Builder*>(b);           \
    if (!B || !B->sim)                                    \
    {                                                     \
        set_error("null or consumed builder handle");     \
        return INCERTO_ERR_INVALID_ARGUMENT;              \
    }                                                     \
    Sim& S = *B->sim;

int32_t incerto_builder_set_seed(incerto_builder* b, uint64_t seed, uint32_t replica)
{
    INC_BUILDER(b);
    S.seed = seed;
    S.replica = replica;
    return INCERTO_OK;
}

int32_t incerto_builder_set_device(incerto_builder* b, int32_t ordinal)
{
    INC_BUILDER(b);
    S.device = ordinal;
    B->device_set = true;
    return INCERTO_OK;
}

int32_t incerto_builder_set_row_partition(incerto_builder* b, uint32_t rank, uint32_t world)
{
    INC_BUILDER(b);
    if (world == 0 || rank >= world) return INCERTO_ERR_INVALID_ARGUMENT;
    S.part_rank = rank;
    S.part_world = world;
    return INCERTO_OK;
}

int32_t incerto_builder_register_component(incerto_builder* b, const char* name, incerto_dtype dtype, uint32_t lanes,
                                           incerto_component* out)
{
    INC_BUILDER(b);
    if (!out || dtype < INCERTO_MARKER || dtype > INCERTO_F64 || lanes < 1 || lanes > 3 ||
        (dtype == INCERTO_MARKER && lanes != 1))
    {
        set_error("register_component: bad dtype/lanes");
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    ComponentInfo ci;
    ci.name = name ? name : "";
    ci.dtype = dtype;
    ci.lanes = lanes;
    if (dtype == INCERTO_MARKER)
    {
        int used = 0;
        for (auto& c : S.comps)
            if (c.dtype == INCERTO_MARKER) ++used;
        if (used >= kMaxMarkers)
        {
            set_error("at most %d marker components per simulation", kMaxMarkers);
            return INCERTO_ERR_UNSUPPORTED;
        }
        ci.marker_bit = used + 1;
    }
    S.comps.push_back(ci);
    *out = static_cast<incerto_component>(S.comps.size() - 1);
    return INCERTO_OK;
}

static int32_t add_entity_spawner(incerto_builder* b, uint64_t n, const incerto_column_data* columns, uint32_t n_columns,
                                  bool borrow)
{
    INC_BUILDER(b);
    if (n_columns && !columns) return INCERTO_ERR_INVALID_ARGUMENT;
    HostSpawn sp;
    sp.n = n;
    sp.borrowed = borrow;
    for (uint32_t i = 0; i < n_columns; ++i)
    {
        const int c = columns[i].component;
        if (c < 0 || c >= static_cast<int>(S.comps.size()))
        {
            set_error("spawner names an unregistered component handle %d", c);
            return INCERTO_ERR_INVALID_ARGUMENT;
        }
        for (auto& prev : sp.cols)
            if (prev.comp == c)
            {
                set_error("component listed twice in one bundle");
                return INCERTO_ERR_INVALID_ARGUMENT;
            }
        const ComponentInfo& ci = S.comps[c];
        HostSpawn::Col col;
        col.comp = c;
        col.has_data = columns[i].data != nullptr;
        if (ci.dtype == INCERTO_MARKER)
        {
            if (col.has_data && borrow)
                col.ext = static_cast<const uint8_t*>(columns[i].data);
            else if (col.has_data)
                col.data.assign(static_cast<const uint8_t*>(columns[i].data), static_cast<const uint8_t*>(columns[i].data) + n);
        }
        else
        {
            const size_t bytes = n * dtype_size(ci.dtype) * ci.lanes;
            if (n && col.has_data && borrow)
                col.ext = static_cast<const uint8_t*>(columns[i].data);
            else if (n && col.has_data)
                col.data.assign(static_cast<const uint8_t*>(columns[i].data), static_cast<const uint8_t*>(columns[i].data) + bytes);
        }
        sp.cols.push_back(std::move(col));
    }
    S.spawns.push_back(std::move(sp));
    return INCERTO_OK;
}

int32_t incerto_builder_add_entity_spawner(incerto_builder* b, uint64_t n, const incerto_column_data* columns,
                                           uint32_t n_columns)
{
    return add_entity_spawner(b, n, columns, n_columns, false);
}

int32_t incerto_builder_add_entity_spawner_borrowed(incerto_builder* b, uint64_t n, const incerto_column_data* columns,
                                                    uint32_t n_columns)
{
    return add_entity_spawner(b, n, columns, n_columns, true);
}

int32_t incerto_builder_add_device_spawner(incerto_builder* b, uint32_t kind, const void* params, uint32_t params_size)
{
    INC_BUILDER(b);
    if (!params || !params_size) return INCERTO_ERR_INVALID_ARGUMENT;
    HostSpawn sp;
    sp.device_kind = true;
    sp.kind = kind;
    sp.params.assign(static_cast<const uint8_t*>(params), static_cast<const uint8_t*>(params) + params_size);
    S.spawns.push_back(std::move(sp));
    return INCERTO_OK;
}

int32_t incerto_builder_add_systems(incerto_builder* b, const incerto_system_desc* systems, uint32_t n)
{
    INC_BUILDER(b);
    if (n && !systems) return INCERTO_ERR_INVALID_ARGUMENT;
    for (uint32_t i = 0; i < n; ++i)
    {
        SystemInst si;
        si.kind = systems[i].kind;
        if (systems[i].params && systems[i].params_size)
            si.params.assign(static_cast<const uint8_t*>(systems[i].params),
                             static_cast<const uint8_t*>(systems[i].params) + systems[i].params_size);
        si.order = static_cast<uint32_t>(S.systems.size());
        S.systems.push_back(std::move(si));
    }
    return INCERTO_OK;
}

int32_t incerto_builder_build(incerto_builder* b, incerto_sim** out)
{
    Builder* B = reinterpret_cast<Builder*>(b);
    if (!B || !B->sim || !out)
    {
        set_error("null or consumed builder handle");
        delete B;
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    std::unique_ptr<Builder> owner(B); // consumed, also on failure (simulation_builder.rs:295 takes self)
    Sim& S = *B->sim;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
    {
        set_error("no CUDA device available (%s); this engine has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return INCERTO_ERR_NO_DEVICE;
    }
    if (S.device < 0 || S.device >= ndev)
    {
        set_error("device ordinal %d out of range (%d devices)", S.device, ndev);
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    stage_mark(nullptr);
    INC_CUDA(cudaSetDevice(S.device));
    INC_CUDA(cudaStreamCreateWithFlags(&S.stream, cudaStreamNonBlocking));
    INC_CUDA(cudaEventCreate(&S.ev_begin));
    INC_CUDA(cudaEventCreate(&S.ev_end));
    INC_CUDA(dev_alloc(&S.d_ticket, 64));
    INC_CUDA(cudaMemsetAsync(S.d_ticket, 0, 64, S.stream));
    INC_CUDA(dev_alloc(&S.d_error, 64));
    INC_CUDA(cudaMemsetAsync(S.d_error, 0, 64, S.stream));
    INC_CUDA(dev_alloc(&S.mods->d_step, 64));
    const uint32_t one = 1;
    INC_CUDA(cudaMemcpyAsync(S.mods->d_step, &one, sizeof(one), cudaMemcpyHostToDevice, S.stream));
    INC_TRY(ensure_scratch(S, 1u << 20));
    S.step = 1; // SimulationBuilder::new runs one empty update (simulation_builder.rs:50)

    // components named by grids and series "exist" like the query state of a bevy system
    for (auto& g : S.grids)
    {
        S.comps[g.position].exists = true;
        S.comps[g.component].exists = true;
    }
    for (auto& se : S.series)
    {
        // the sampling systems hold Query<&C, F> / Query<(&C, &Id)>: their components get registered
        if (se.per_entity)
        {
            S.comps[se.component].exists = true;
            S.comps[se.identifier].exists = true;
            continue;
        }
        S.comps[se.agg.component].exists = true;
        if (se.agg.component2 >= 0) S.comps[se.agg.component2].exists = true;
        for (int c : se.agg.filter.with) S.comps[c].exists = true;
        for (int c : se.agg.filter.without) S.comps[c].exists = true;
    }
    stage_mark("build: stream, scratch");
    INC_TRY(run_spawners(S, true));
    stage_mark("build: run_spawners");
    INC_TRY(bind_generic_systems(S));
    INC_TRY(grids_allocate(S));
    INC_CUDA(cudaStreamSynchronize(S.stream));
    stage_mark("build: bind, grids, sync");
    for (auto& sp : S.spawns)
        if (sp.borrowed)
            for (auto& c : sp.cols) c.ext = nullptr; // the caller's buffers are only promised until here
    *out = reinterpret_cast<incerto_sim*>(B->sim.release());
    return INCERTO_OK;
}

#define INC_SIM(s)                                \
    Sim* SP = reinterpret_cast<Sim*>(s);          \
    if (!SP)                                      \
    {                                             \
        set_error("null simulation handle");      \
        return INCERTO_ERR_INVALID_ARGUMENT;      \
    }                                             \
    Sim& S = *SP;

int32_t incerto_sim_run(incerto_sim* s, uint64_t num_steps)
{
    INC_SIM(s);
    INC_TRY(do_run(S, num_steps));
    return incerto_sim_synchronize(s);
}

int32_t incerto_sim_run_async(incerto_sim* s, uint64_t num_steps)
{
    INC_SIM(s);
    return do_run(S, num_steps);
}

// Bench hygiene in one call: `num_steps` steps, each started from a flushed L2, everything enqueued back to back (no
// host round trip between the steps, so the ranks of a partitioned grid stay in lock step on the device); the device
// time reported by incerto_sim_last_run_stats is the sum over the steps of the CUDA-event time around each step alone.
int32_t incerto_sim_run_cold(incerto_sim* s, uint64_t num_steps)
{
    INC_SIM(s);
    INC_CUDA(cudaSetDevice(S.device));
    if (num_steps == 0 || num_steps > 65536) return INCERTO_ERR_INVALID_ARGUMENT;
    INC_TRY(check_step_range(S, num_steps));
    if (S.run_pending) INC_TRY(incerto_sim_synchronize(s));
    INC_TRY(incerto_sim_flush_l2(s)); // allocates the flush buffer on first use
    std::vector<cudaEvent_t> ev(2 * num_steps, nullptr);
    int32_t rc = INCERTO_OK;
    auto fail = [&](cudaError_t e) {
        if (e != cudaSuccess && rc == INCERTO_OK)
        {
            set_error("run_cold: %s", cudaGetErrorString(e));
            rc = INCERTO_ERR_CUDA;
        }
    };
    for (auto& e : ev) fail(cudaEventCreate(&e));
    if (S.comm && rc == INCERTO_OK) rc = ensure_scratch(S, 64);
    const uint64_t launches0 = S.launches;
    if (rc == INCERTO_OK) rc = series_prepare(S, S.step, num_steps);
    for (auto& op : S.mods->program)
        if (rc == INCERTO_OK && op.prepare) rc = op.prepare(S, 1);
    for (uint64_t i = 0; i < num_steps && rc == INCERTO_OK; ++i)
    {
        rc = incerto_sim_flush_l2(s);
        if (rc != INCERTO_OK) break;
        if (S.comm)
        {
            // ranks of a partitioned grid: line them up again after the (untimed) flush, whose duration differs from
            // GPU to GPU — otherwise that skew would be charged to the step as time spent waiting for the neighbour
            rc = comm_allreduce(S, S.d_scratch, 1, 0, 0);
            if (rc != INCERTO_OK) break;
            // ... and the neighbours once more over peer memory (they leave within one NVLink latency of each other;
            // the collective's kernels end a few microseconds apart, and a step waits for its neighbours' launch)
            rc = forest_cold_rendezvous(S);
            if (rc != INCERTO_OK) break;
        }
        fail(cudaEventRecord(ev[2 * i], S.stream));
        if (rc == INCERTO_OK) rc = do_run_enqueue(S, 1);
        fail(cudaEventRecord(ev[2 * i + 1], S.stream));
    }
    // the record of the last step (every earlier one was published by its successor's grid, inside that step's event
    // pair): timed by a pair of its own and charged to the run like the steps
    cudaEvent_t ev_fin[2] = {nullptr, nullptr};
    for (auto& e : ev_fin) fail(cudaEventCreate(&e));
    fail(cudaEventRecord(ev_fin[0], S.stream));
    if (rc == INCERTO_OK) rc = forest_finalize(S);
    fail(cudaEventRecord(ev_fin[1], S.stream));
    fail(cudaStreamSynchronize(S.stream));
    double total = 0.0;
    if (rc == INCERTO_OK)
    {
        float ms = 0.f;
        fail(cudaEventElapsedTime(&ms, ev_fin[0], ev_fin[1]));
        total += ms;
    }
    for (auto& e : ev_fin)
        if (e) cudaEventDestroy(e);
    for (uint64_t i = 0; i < num_steps && rc == INCERTO_OK; ++i)
    {
        float ms = 0.f;
        fail(cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]));
        total += ms;
    }
    for (auto& e : ev)
        if (e) cudaEventDestroy(e);
    series_after_run(S);
    if (rc != INCERTO_OK) return rc;
    S.last_ms = total;
    S.last_launches = S.launches - launches0;
    return incerto_sim_synchronize(s); // resolves the per-kernel timers and checks the device error flag
}

// Device time from the start of `first`'s last run to the end of `last`'s last run (CUDA events on the two handles'
// streams, same device): the span of several handles driven concurrently with run_async.  Both must be synchronised.
int32_t incerto_sim_span_ms(incerto_sim* first, incerto_sim* last, double* ms_out)
{
    Sim* A = reinterpret_cast<Sim*>(first);
    Sim* B = reinterpret_cast<Sim*>(last);
    if (!A || !B || !ms_out || A->device != B->device) return INCERTO_ERR_INVALID_ARGUMENT;
    if (A->run_pending || B->run_pending)
    {
        set_error("span_ms: synchronize both handles first");
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    INC_CUDA(cudaSetDevice(A->device));
    float ms = 0.f;
    INC_CUDA(cudaEventElapsedTime(&ms, A->ev_begin, B->ev_end));
    *ms_out = ms;
    return INCERTO_OK;
}

int32_t incerto_sim_synchronize(incerto_sim* s)
{
    INC_SIM(s);
    INC_CUDA(cudaSetDevice(S.device));
    INC_CUDA(cudaStreamSynchronize(S.stream));
    if (S.run_pending)
    {
        float ms = 0.f;
        INC_CUDA(cudaEventElapsedTime(&ms, S.ev_begin, S.ev_end));
        S.last_ms = ms;
        S.last_launches = S.launches - S.last_launches;
        S.run_pending = false;
    }
    resolve_timers(S);
    uint32_t err[2] = {0, 0};
    INC_CUDA(cudaMemcpy(err, S.d_error, sizeof(err), cudaMemcpyDeviceToHost));
    if (err[0])
    {
        INC_CUDA(cudaMemset(S.d_error, 0, 64));
        set_error("device raised error flag %u (detail %u)", err[0], err[1]);
        return static_cast<int32_t>(err[0]);
    }
    return INCERTO_OK;
}

void incerto_sim_destroy(incerto_sim* s) { delete reinterpret_cast<Sim*>(s); }

int32_t incerto_sim_reset(incerto_sim* s, uint32_t replica)
{
    INC_SIM(s);
    INC_CUDA(cudaSetDevice(S.device));
    for (auto& sp : S.spawns)
        if (sp.borrowed)
        {
            set_error("reset: the entities were spawned from borrowed buffers (add_entity_spawner_borrowed), which are gone");
            return INCERTO_ERR_UNSUPPORTED;
        }
    INC_CUDA(cudaStreamSynchronize(S.stream));
    S.replica = replica;
    S.step = 1;
    const uint32_t one = 1;
    INC_CUDA(cudaMemcpyAsync(S.mods->d_step, &one, sizeof(one), cudaMemcpyHostToDevice, S.stream));
    series_clear(S);
    if (S.mods->forest.active) INC_TRY(forest_peer_quiesce(S));
    INC_TRY(run_spawners(S, false));
    if (S.mods->forest.active) INC_TRY(forest_reset_after_spawn(S));
    if (S.mods->pandemic.active) INC_TRY(pandemic_reset(S));
    if (S.mods->air.active) INC_TRY(air_reset(S));
    if (S.mods->spatial.active) INC_TRY(spatial_reset(S));
    for (auto& g : S.grids)
    {
        g.built_step = ~0ull;
        g.snap_valid = g.snap_built = false;
    }
    INC_CUDA(cudaStreamSynchronize(S.stream));
    return INCERTO_OK;
}

int32_t incerto_sim_step_number(incerto_sim* s, uint64_t* out)
{
    INC_SIM(s);
    if (!out) return INCERTO_ERR_INVALID_ARGUMENT;
    *out = S.step;
    return INCERTO_OK;
}

int32_t incerto_sim_last_run_stats(incerto_sim* s, double* device_ms, uint64_t* kernel_launches)
{
    INC_SIM(s);
    if (S.run_pending) INC_TRY(incerto_sim_synchronize(s));
    if (device_ms) *device_ms = S.last_ms;
    if (kernel_launches) *kernel_launches = S.last_launches;
    return INCERTO_OK;
}

int32_t incerto_sim_set_profiling(incerto_sim* s, uint32_t enabled)
{
    INC_SIM(s);
    S.profiling = enabled != 0;
    return INCERTO_OK;
}

int32_t incerto_sim_kernel_stats(incerto_sim* s, incerto_kernel_stat* out, uint32_t capacity, uint32_t* n_out)
{
    INC_SIM(s);
    if (S.run_pending) INC_TRY(incerto_sim_synchronize(s));
    uint32_t n = 0;
    for (auto& k : S.kstats)
    {
        if (out && n < capacity)
        {
            out[n].name = k.name;
            out[n].launches = k.launches;
            out[n].device_ms = k.device_ms;
            out[n].algorithmic_bytes = k.algorithmic_bytes;
        }
        ++n;
    }
    if (n_out) *n_out = n;
    return INCERTO_OK;
}

int32_t incerto_sim_flush_l2(incerto_sim* s)
{
    INC_SIM(s);
    INC_CUDA(cudaSetDevice(S.device));
    if (!S.d_flush)
    {
        S.flush_bytes = 512ull << 20; // 4x the 126 MB L2
        if (const char* e = std::getenv("INCERTO_B200_FLUSH_MB")) // experiments
            S.flush_bytes = static_cast<uint64_t>(std::max(1l, std::atol(e))) << 20;
        INC_CUDA(dev_alloc(&S.d_flush, S.flush_bytes));
    }
    static uint32_t v = 0;
    launch_flush(S.d_flush, S.flush_bytes, ++v, S.stream);
    INC_CUDA(cudaGetLastError());
    return INCERTO_OK;
}

} // extern "C"
