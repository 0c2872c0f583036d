// pandemic_spatial module: host side of kernels_spatial.cu (examples/pandemic_spatial.rs).
#include <algorithm>

#include "kernels_grid.hpp"
#include "kernels_spatial.hpp"
#include "modules.hpp"

namespace incerto
{

static bool valid_comp(const Sim& s, int c) { return c >= 0 && c < static_cast<int>(s.comps.size()); }

// ContactHistory is engine-managed (20 packed positions per person, contiguous: one 80-byte row); called by run_spawners
// between counting and allocation.
int32_t spatial_layout(Sim& s)
{
    int hist = -1;
    for (auto& sys : s.systems)
        if (sys.kind >= INCERTO_SYS_SPATIAL_INCUBATION && sys.kind <= INCERTO_SYS_SPATIAL_MOVE &&
            sys.params.size() == sizeof(incerto_spatial_params))
        {
            incerto_spatial_params p;
            std::memcpy(&p, sys.params.data(), sizeof(p));
            hist = p.contact_history;
        }
    for (auto& sp : s.spawns)
        if (sp.device_kind && sp.kind == INCERTO_SPAWN_SPATIAL_POPULATION && sp.params.size() == sizeof(incerto_spawn_spatial_params))
        {
            incerto_spawn_spatial_params p;
            std::memcpy(&p, sp.params.data(), sizeof(p));
            hist = p.contact_history;
        }
    if (!valid_comp(s, hist)) return INCERTO_OK;
    s.comps[hist].managed_bytes = 4 * kSpatialHistorySlots;
    for (auto& t : s.tables)
        if (Column* c = t.col(hist)) c->row_bytes = s.comps[hist].managed_bytes;
    return INCERTO_OK;
}

static int32_t spatial_check_components(Sim& s, int position, int disease, int social, int quarantined, int history)
{
    for (int c : {position, disease, social, quarantined, history})
        if (!valid_comp(s, c)) return INCERTO_ERR_INVALID_ARGUMENT;
    const ComponentInfo& pc = s.comps[position];
    if ((pc.dtype != INCERTO_I16 && pc.dtype != INCERTO_U16) || pc.lanes != 2 || s.comps[disease].dtype != INCERTO_U8 ||
        s.comps[disease].lanes != 1 || s.comps[quarantined].dtype != INCERTO_U8 || s.comps[quarantined].lanes != 1 ||
        s.comps[social].dtype != INCERTO_MARKER)
    {
        set_error("pandemic_spatial: GridPosition2D must be 2-lane I16, disease_state and Quarantined U8 columns, "
                  "social_distancing a marker (device layout)");
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    return INCERTO_OK;
}

int32_t spatial_bind(Sim& s)
{
    SpatialModule& sm = s.mods->spatial;
    bool first = true;
    for (auto& sys : s.systems)
    {
        if (sys.kind < INCERTO_SYS_SPATIAL_INCUBATION || sys.kind > INCERTO_SYS_SPATIAL_MOVE) continue;
        if (sys.params.size() != sizeof(incerto_spatial_params)) return INCERTO_ERR_INVALID_ARGUMENT;
        incerto_spatial_params p;
        std::memcpy(&p, sys.params.data(), sizeof(p));
        INC_TRY(spatial_check_components(s, p.position, p.disease_state, p.social_distancing, p.quarantined, p.contact_history));
        if (!first && std::memcmp(&sm.p, &p, sizeof(p)) != 0)
        {
            set_error("pandemic_spatial systems must be given identical parameter blocks");
            return INCERTO_ERR_INVALID_ARGUMENT;
        }
        sm.p = p;
        first = false;
        sm.systems |= 1u << (sys.kind - INCERTO_SYS_SPATIAL_INCUBATION);
        for (int c : {p.position, p.disease_state, p.social_distancing, p.quarantined, p.contact_history}) s.comps[c].exists = true;
    }
    const incerto_spatial_params& p = sm.p;
    if (p.grid_size < 1 || p.grid_size > 32767 || p.incubation_period < 1 || p.incubation_period > 127 ||
        p.contact_quarantine_duration < 1 || p.contact_quarantine_duration > 255 ||
        p.contact_history_clear_above > kSpatialHistorySlots || p.infection_radius < 0)
    {
        set_error("pandemic_spatial: grid_size 1..=32767, incubation 1..=127, quarantine duration 1..=255, history clear "
                  "threshold <= %u", kSpatialHistorySlots);
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    sm.active = true;
    sm.table = table_for_components(s, {p.position, p.disease_state, p.quarantined, p.contact_history});
    for (size_t gi = 0; gi < s.grids.size(); ++gi)
        if (s.grids[gi].dim == 2 && s.grids[gi].position == p.position)
        {
            // the kernels keep their own occupancy / source / mark tables: the dense entity index is only built
            // when the host queries it
            s.grids[gi].lazy = true;
            sm.grid = static_cast<int>(gi);
        }
    if (sm.table < 0) return INCERTO_OK;
    Table& t = s.tables[sm.table];
    if (t.n >= (1ull << kSpatialRowBits) - 1)
    {
        set_error("pandemic_spatial: at most 2^26-2 people per world (source lists link rows with 26 bits)");
        return INCERTO_ERR_UNSUPPORTED;
    }
    const uint64_t cells = static_cast<uint64_t>(p.grid_size) * p.grid_size;
    const uint64_t n = std::max<uint64_t>(t.n, 1);
    INC_CUDA(dev_alloc(&sm.d_aux, n));
    INC_CUDA(dev_alloc(&sm.d_count, cells * 4));
    INC_CUDA(dev_alloc(&sm.d_src_bits, (cells / 32 + 1) * 4));
    INC_CUDA(dev_alloc(&sm.d_src_head, cells * 4));
    INC_CUDA(dev_alloc(&sm.d_src_next, n * 4));
    INC_CUDA(dev_alloc(&sm.d_mark, cells * 8));
    INC_CUDA(dev_alloc(&sm.d_marked, (n / 32 + 2) * 4));
    INC_CUDA(dev_alloc(&sm.d_stats_acc, 64));
    return spatial_reset(s);
}

// after (re)spawning: nobody is counted yet, nobody remembers anything
int32_t spatial_reset(Sim& s)
{
    SpatialModule& sm = s.mods->spatial;
    if (sm.table < 0 || !sm.d_aux) return INCERTO_OK;
    Table& t = s.tables[sm.table];
    const uint64_t cells = static_cast<uint64_t>(sm.p.grid_size) * sm.p.grid_size;
    INC_CUDA(cudaMemsetAsync(sm.d_aux, 6u << 5, std::max<uint64_t>(t.n_spawned, 1), s.stream)); // OP_NEW, empty history
    INC_CUDA(cudaMemsetAsync(sm.d_count, 0, cells * 4, s.stream));
    INC_CUDA(cudaMemsetAsync(sm.d_src_head, 0, cells * 4, s.stream));
    INC_CUDA(cudaMemsetAsync(sm.d_mark, 0, cells * 8, s.stream)); // kept incrementally from here on
    INC_CUDA(cudaMemsetAsync(sm.d_marked, 0, (std::max<uint64_t>(t.n_spawned, 1) / 32 + 2) * 4, s.stream));
    sm.stats_valid = false;
    // rows in cell order: the per-cell tables are then touched sector by sector (people move one cell at a time and
    // mostly sit in quarantine, so the order stays good for hundreds of steps)
    GridDesc g;
    g.dim = 2;
    g.bmin[0] = g.bmin[1] = g.bmin[2] = 0;
    g.ext[0] = g.ext[1] = sm.p.grid_size;
    g.ext[2] = 1;
    return table_sort_by_cell(s, sm.table, sm.p.position, g, {sm.p.contact_history});
}

void spatial_free(Sim& s)
{
    SpatialModule& sm = s.mods->spatial;
    dev_free(sm.d_aux);
    dev_free(sm.d_count);
    dev_free(sm.d_src_bits);
    dev_free(sm.d_src_head);
    dev_free(sm.d_src_next);
    dev_free(sm.d_mark);
    dev_free(sm.d_stats_acc);
    dev_free(sm.d_marked);
    sm.d_marked = nullptr;
    sm.d_stats_acc = nullptr;
    sm.d_aux = nullptr;
    sm.d_count = sm.d_src_bits = sm.d_src_head = sm.d_src_next = nullptr;
    sm.d_mark = nullptr;
}

static void spatial_fill(Sim& s, SpatialModule& sm, SpatialArgs& a, Table& t, uint64_t step)
{
    const incerto_spatial_params& p = sm.p;
    std::memset(&a, 0, sizeof(a));
    a.pos = static_cast<uint32_t*>(t.col(p.position)->d);
    a.disease = static_cast<uint8_t*>(t.col(p.disease_state)->d);
    a.flags = t.d_flags;
    a.quar = static_cast<uint8_t*>(t.col(p.quarantined)->d);
    a.aux = sm.d_aux;
    a.hist = static_cast<uint32_t*>(t.col(p.contact_history)->d);
    a.uid = t.d_uid;
    a.n = t.n;
    a.n_cap = t.n_spawned;
    a.G = p.grid_size;
    a.step = static_cast<uint32_t>(step);
    a.systems = sm.systems;
    a.bit_social = 1u << s.comps[p.social_distancing].marker_bit;
    a.radius = static_cast<uint32_t>(p.infection_radius);
    a.incubation = p.incubation_period;
    a.crowding = p.crowding_threshold;
    a.quarantine_duration = p.contact_quarantine_duration;
    a.history_cap = std::max<uint32_t>(14u, p.contact_history_clear_above); // `len > 14` guards `len > 20` (:549-555)
    a.zone_x = p.quarantine_center_x;
    a.zone_y = p.quarantine_center_y;
    a.zone_radius = p.quarantine_radius;
    a.thr_infect1 = bernoulli_threshold(p.chance_infect_distance_1);
    a.thr_infect2 = bernoulli_threshold(p.chance_infect_distance_2);
    a.thr_recover = bernoulli_threshold(p.chance_recover);
    a.thr_die = bernoulli_threshold(p.chance_die);
    a.thr_attempt = bernoulli_threshold(p.chance_attempt_move);
    a.rk_move = make_round_keys(make_key(s.seed, s.replica, TAG_SPATIAL_MOVE));
    a.rk_transmit = make_round_keys(make_key(s.seed, s.replica, TAG_SPATIAL_TRANSMIT));
    a.rk_fate = make_round_keys(make_key(s.seed, s.replica, TAG_SPATIAL_FATE));
    a.count = sm.d_count;
    a.src_bits = sm.d_src_bits;
    a.src_head = sm.d_src_head;
    a.src_next = sm.d_src_next;
    a.tag = static_cast<uint32_t>(step % 63) + 1;
    a.mark = sm.d_mark;
    a.marked = sm.d_marked;
    a.stats_acc = sm.d_stats_acc;
    a.error = s.d_error;
}

int32_t spatial_step(Sim& s, uint64_t first_step, uint32_t nsteps)
{
    SpatialModule& sm = s.mods->spatial;
    if (sm.table < 0) return INCERTO_OK;
    Table& t = s.tables[sm.table];
    if (!t.n) return INCERTO_OK;
    const uint64_t cells = static_cast<uint64_t>(sm.p.grid_size) * sm.p.grid_size;
    for (uint32_t i = 0; i < nsteps; ++i)
    {
        const uint64_t step = first_step + i;
        SpatialArgs a;
        spatial_fill(s, sm, a, t, step);
        if (sm.systems & 2u)
        {
            INC_CUDA(cudaMemsetAsync(sm.d_src_bits, 0, (cells / 32 + 1) * 4, s.stream));
            if (step % 63 == 0) INC_CUDA(cudaMemsetAsync(sm.d_src_head, 0, cells * 4, s.stream)); // the tag comes round again
        }
        {
            // pos 4 + disease 1 + flags 1 + quarantine 1 + aux 1 read, aux 1 written; history and tables on top
            KTimer kt(s, "spatial_prepare", 9.0 * t.n);
            launch_spatial_prepare(a, s.stream);
        }
        {
            KTimer kt(s, "spatial_interact", 8.0 * t.n);
            launch_spatial_interact(a, s.stream);
        }
        INC_CUDA(cudaGetLastError());
        if (sm.systems & 4u) t.all_alive = false;
        sm.stats_valid = true; // the interact kernel left the PandemicStats counts of the post-flush state
    }
    return INCERTO_OK;
}

// INCERTO_SPAWN_SPATIAL_POPULATION: examples/pandemic_spatial.rs:279-313 with the Philox contract
int32_t spatial_spawn_device(Sim& s, const HostSpawn& sp, Table& t, uint64_t row_offset)
{
    incerto_spawn_spatial_params p;
    std::memcpy(&p, sp.params.data(), sizeof(p));
    if (row_offset != 0)
    {
        set_error("the population device spawner must be the only spawner of its archetype");
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    INC_TRY(spatial_check_components(s, p.position, p.disease_state, p.social_distancing, p.quarantined, p.contact_history));
    if (p.grid_size < 1 || p.grid_size > 32767 || p.incubation_period < 1 || p.incubation_period > 127) return INCERTO_ERR_INVALID_ARGUMENT;
    SpatialArgs a;
    std::memset(&a, 0, sizeof(a));
    a.pos = static_cast<uint32_t*>(t.col(p.position)->d);
    a.disease = static_cast<uint8_t*>(t.col(p.disease_state)->d);
    a.flags = t.d_flags;
    a.quar = static_cast<uint8_t*>(t.col(p.quarantined)->d);
    a.n = p.initial_population;
    a.G = p.grid_size;
    a.incubation = p.incubation_period;
    a.bit_social = 1u << s.comps[p.social_distancing].marker_bit;
    a.thr_social = bernoulli_threshold(p.social_distance_compliance);
    a.thr_start = bernoulli_threshold(p.chance_start_infected);
    a.rk_spawn = make_round_keys(make_key(s.seed, s.replica, TAG_SPATIAL_SPAWN));
    KTimer kt(s, "spatial_spawn", 7.0 * a.n);
    launch_spatial_spawn(a, s.stream);
    INC_CUDA(cudaGetLastError());
    return INCERTO_OK;
}

// read_component of the managed ContactHistory: per live person 21 u32 = length, then the remembered
// positions (x | y << 16) in insertion order, zero padded
int32_t spatial_read_history(Sim& s, int table, uint8_t* out, size_t out_size, size_t* written)
{
    SpatialModule& sm = s.mods->spatial;
    Table& t = s.tables[table];
    Column* c = t.col(sm.p.contact_history);
    std::vector<uint32_t> h(static_cast<size_t>(kSpatialHistorySlots) * t.n_spawned);
    std::vector<uint8_t> aux(t.n);
    std::vector<uint32_t> rows, uids;
    INC_TRY(live_rows_by_uid(s, t, rows, uids));
    INC_CUDA(cudaMemcpyAsync(h.data(), c->d, h.size() * 4, cudaMemcpyDeviceToHost, s.stream));
    INC_CUDA(cudaMemcpyAsync(aux.data(), sm.d_aux, t.n, cudaMemcpyDeviceToHost, s.stream));
    INC_CUDA(cudaStreamSynchronize(s.stream));
    size_t w = 0;
    const size_t rec = (kSpatialHistorySlots + 1) * 4;
    for (uint32_t r : rows)
    {
        if (w + rec > out_size)
        {
            set_error("read_component: output buffer too small");
            return INCERTO_ERR_INVALID_ARGUMENT;
        }
        uint32_t row[kSpatialHistorySlots + 1] = {0};
        row[0] = aux[r] & 31u;
        for (uint32_t k = 0; k < row[0]; ++k) row[1 + k] = h[static_cast<size_t>(r) * kSpatialHistorySlots + k];
        std::memcpy(out + w, row, rec);
        w += rec;
    }
    *written = w;
    return INCERTO_OK;
}

int32_t spatial_stats_now(Sim& s, uint64_t initial_population, uint8_t* d_raw)
{
    SpatialModule& sm = s.mods->spatial;
    if (sm.table < 0) return INCERTO_ERR_AGGREGATE_NO_ENTITIES;
    Table& t = s.tables[sm.table];
    if (sm.stats_valid)
    {
        KTimer kt(s, "spatial_stats_fused", 0.0);
        launch_spatial_stats_finalize(sm.d_stats_acc, initial_population, reinterpret_cast<uint64_t*>(d_raw), s.stream);
        INC_CUDA(cudaGetLastError());
        return INCERTO_OK;
    }
    INC_TRY(ensure_scratch(s, 64 * kNumSMs * 8 + 4096));
    KTimer kt(s, "spatial_stats", 2.0 * t.n);
    launch_spatial_stats(static_cast<const uint8_t*>(t.col(sm.p.disease_state)->d), t.d_flags, t.n, initial_population,
                         reinterpret_cast<uint64_t*>(d_raw), s.d_scratch, s.d_ticket, s.stream);
    INC_CUDA(cudaGetLastError());
    return INCERTO_OK;
}

} // namespace incerto
