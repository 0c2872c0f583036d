// Air-traffic module: host side of kernels_air.cu (examples/air_traffic_3d.rs).
#include <algorithm>

#include "kernels_air.hpp"
#include "kernels_grid.hpp"
#include "modules.hpp"

namespace incerto
{

static int32_t air_check(Sim& s, AirModule& am, const incerto_air_params& p)
{
    auto bad = [&](int c) { return c < 0 || c >= static_cast<int>(s.comps.size()); };
    if (bad(p.position) || bad(p.target_altitude)) return INCERTO_ERR_INVALID_ARGUMENT;
    const ComponentInfo& pc = s.comps[p.position];
    const ComponentInfo& tc = s.comps[p.target_altitude];
    if (pc.dtype != INCERTO_I16 || pc.lanes != 3 || tc.dtype != INCERTO_I16 || tc.lanes != 1)
    {
        set_error("air_traffic: GridPosition3D must be a 3-lane I16 component and target_altitude an I16 column (device layout)");
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    if (p.airspace_size < 1 || p.airspace_size > 32767 || p.airspace_height < 1 || p.airspace_height > 32767)
    {
        set_error("air_traffic: AIRSPACE_SIZE / AIRSPACE_HEIGHT must be in 1..=32767");
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    if (am.comp_pos >= 0 && (am.comp_pos != p.position || am.comp_target != p.target_altitude ||
                             am.size != p.airspace_size || am.height != p.airspace_height))
    {
        set_error("air_traffic systems must agree on their components and airspace");
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    am.comp_pos = p.position;
    am.comp_target = p.target_altitude;
    am.size = p.airspace_size;
    am.height = p.airspace_height;
    s.comps[p.position].exists = true;
    s.comps[p.target_altitude].exists = true;
    return INCERTO_OK;
}

static int32_t air_sort_rows(Sim& s);

int32_t air_bind(Sim& s)
{
    AirModule& am = s.mods->air;
    for (auto& sys : s.systems)
    {
        if (sys.kind != INCERTO_SYS_AIR_CHECK_CONFLICTS && sys.kind != INCERTO_SYS_AIR_MOVE) continue;
        if (sys.params.size() != sizeof(incerto_air_params)) return INCERTO_ERR_INVALID_ARGUMENT;
        incerto_air_params p;
        std::memcpy(&p, sys.params.data(), sizeof(p));
        INC_TRY(air_check(s, am, p));
        if (sys.kind == INCERTO_SYS_AIR_MOVE)
        {
            am.do_move = true;
            am.p_move = p.chance_horizontal_move;
        }
        else
            am.do_check = true;
    }
    am.active = true;
    am.table = table_for_components(s, {am.comp_pos, am.comp_target});
    if (am.do_check)
    {
        // check_conflicts takes Res<SpatialGrid3D<Aircraft>> (:125): the grid must have been added
        for (size_t gi = 0; gi < s.grids.size(); ++gi)
            if (s.grids[gi].dim == 3 && s.grids[gi].position == am.comp_pos) am.grid = static_cast<int>(gi);
        if (am.grid < 0 || !s.grids[am.grid].bounded)
        {
            set_error("check_conflicts needs a bounded 3-D spatial grid on its position component (add_spatial_grid_3d)");
            return INCERTO_ERR_INVALID_ARGUMENT;
        }
        GridInst& g = s.grids[am.grid];
        uint64_t cells = 1;
        for (int d = 0; d < 3; ++d) cells *= static_cast<uint64_t>(g.bmax[d] - g.bmin[d] + 1);
        if (cells >= (1ull << 32))
        {
            set_error("air_traffic: the airspace grid has too many cells (%llu >= 2^32)", static_cast<unsigned long long>(cells));
            return INCERTO_ERR_UNSUPPORTED;
        }
        g.lazy = true; // the kernels use their own occupancy index; the dense table is only built for host queries
    }
    INC_CUDA(cudaMalloc(&am.d_acc, 64));
    INC_CUDA(cudaMemsetAsync(am.d_acc, 0, 64, s.stream));
    if (am.table < 0) return INCERTO_OK;
    Table& t = s.tables[am.table];
    if (am.do_check)
    {
        if (t.n >= (1ull << 24))
        {
            // per-cell counts are 24 bit; more aircraft than that could in principle share one cell
            if (t.n >= (1ull << 31)) return INCERTO_ERR_UNSUPPORTED;
        }
        uint64_t want = std::max<uint64_t>(1024, 2 * t.n);
        uint32_t bits = 10;
        while ((1ull << bits) < want) ++bits;
        am.n_slots = 1u << bits;
        am.slot_bits = bits;
        INC_CUDA(cudaMalloc(&am.d_slots, static_cast<size_t>(am.n_slots) * 8));
        INC_CUDA(cudaMemsetAsync(am.d_slots, 0, static_cast<size_t>(am.n_slots) * 8, s.stream));
        // pre-filter bitmap: >= 32 bits per aircraft (at most ~3 % of the bits set)
        uint32_t fb = 16;
        while ((1ull << fb) < 32 * t.n && fb < 32) ++fb;
        am.filter_bits = fb;
        INC_CUDA(cudaMalloc(&am.d_bits, (1ull << fb) / 8));
        INC_CUDA(cudaMalloc(&am.d_bits_next, (1ull << fb) / 8));
        INC_CUDA(cudaMalloc(&am.d_cand, std::max<uint64_t>(t.n, 2)));
    }
    return air_sort_rows(s);
}

// rows in (x, y) column order: neighbouring aircraft then probe neighbouring sectors of the pre-filter
static int32_t air_sort_rows(Sim& s)
{
    AirModule& am = s.mods->air;
    if (am.table < 0 || am.grid < 0 || !am.do_check) return INCERTO_OK;
    const GridInst& gi = s.grids[am.grid];
    GridDesc g;
    g.dim = 2; // z (lane 2) is ignored: one sort bin per column
    for (int d = 0; d < 3; ++d)
    {
        g.bmin[d] = d < 2 ? gi.bmin[d] : 0;
        g.ext[d] = d < 2 ? gi.bmax[d] - gi.bmin[d] + 1 : 1;
    }
    return table_sort_by_cell(s, am.table, am.comp_pos, g, {});
}

int32_t air_reset(Sim& s)
{
    AirModule& am = s.mods->air;
    if (am.d_slots) INC_CUDA(cudaMemsetAsync(am.d_slots, 0, static_cast<size_t>(am.n_slots) * 8, s.stream));
    if (am.d_acc) INC_CUDA(cudaMemsetAsync(am.d_acc, 0, 64, s.stream));
    am.bits_ready = false;
    return air_sort_rows(s);
}

void air_free(Sim& s)
{
    AirModule& am = s.mods->air;
    cudaFree(am.d_slots);
    cudaFree(am.d_acc);
    cudaFree(am.d_bits);
    cudaFree(am.d_bits_next);
    am.d_bits_next = nullptr;
    cudaFree(am.d_cand);
    am.d_bits = nullptr;
    am.d_cand = nullptr;
    am.d_slots = nullptr;
    am.d_acc = nullptr;
}

static void air_fill(Sim& s, AirModule& am, AirArgs& a, Table& t, uint64_t step)
{
    std::memset(&a, 0, sizeof(a));
    a.pos = static_cast<uint2*>(t.col(am.comp_pos)->d);
    a.target = static_cast<const int16_t*>(t.col(am.comp_target)->d);
    a.flags = t.all_alive ? nullptr : t.d_flags;
    a.uid = t.d_uid;
    a.n = t.n;
    a.size = am.size;
    a.height = am.height;
    a.step = static_cast<uint32_t>(step);
    a.do_check = am.do_check;
    a.do_move = am.do_move;
    a.thr_move = bernoulli_threshold(am.p_move);
    a.rk_move = make_round_keys(make_key(s.seed, s.replica, TAG_AIR_MOVE));
    if (am.grid >= 0)
    {
        const GridInst& g = s.grids[am.grid];
        for (int d = 0; d < 3; ++d)
        {
            a.bmin[d] = g.bmin[d];
            a.ext[d] = g.bmax[d] - g.bmin[d] + 1;
        }
    }
    a.slots = am.d_slots;
    a.slot_mask = am.n_slots ? am.n_slots - 1 : 0;
    a.slot_shift = 32 - am.slot_bits;
    a.tag = static_cast<uint32_t>(step % 255) + 1;
    a.acc = am.d_acc;
    a.bits = am.d_bits;
    a.bits_next = am.do_check ? am.d_bits_next : nullptr;
    a.bits_shift = 32 - (am.filter_bits - 8); // 256-bit sectors are hashed, the low 8 key bits index inside
    a.cand = am.d_cand;
    a.error = s.d_error;
}

int32_t air_step(Sim& s, uint64_t first_step, uint32_t nsteps)
{
    AirModule& am = s.mods->air;
    if (am.table < 0) return INCERTO_OK;
    Table& t = s.tables[am.table];
    if (!t.n) return INCERTO_OK;
    for (uint32_t i = 0; i < nsteps; ++i)
    {
        const uint64_t step = first_step + i;
        AirArgs a;
        air_fill(s, am, a, t, step);
        if (am.do_check)
        {
            // the tag of 255 steps ago comes round again: drop those entries first
            if (step % 255 == 0) INC_CUDA(cudaMemsetAsync(am.d_slots, 0, static_cast<size_t>(am.n_slots) * 8, s.stream));
            if (!am.bits_ready)
            {
                INC_CUDA(cudaMemsetAsync(am.d_bits, 0, (1ull << am.filter_bits) / 8, s.stream));
                KTimer kt(s, "air_bits", 8.0 * t.n);
                launch_air_bits(a, s.stream);
            }
            INC_CUDA(cudaMemsetAsync(am.d_bits_next, 0, (1ull << am.filter_bits) / 8, s.stream));
            {
                KTimer kt(s, "air_candidates", 9.0 * t.n);
                launch_air_candidates(a, s.stream);
            }
        }
        {
            KTimer kt(s, "air_step", (am.do_move ? 18.0 : 8.0) * t.n + (am.do_check ? 1.0 : 0.0) * t.n);
            launch_air_step(a, s.stream);
        }
        INC_CUDA(cudaGetLastError());
        if (am.do_check)
        {
            std::swap(am.d_bits, am.d_bits_next); // the step kernel left the next PreUpdate's bitmap behind
            am.bits_ready = true;
        }
    }
    return INCERTO_OK;
}

// INCERTO_SPAWN_AIRCRAFT: examples/air_traffic_3d.rs:61-90 with the Philox contract
int32_t air_spawn_device(Sim& s, const HostSpawn& sp, Table& t, uint64_t row_offset)
{
    incerto_spawn_aircraft_params p;
    std::memcpy(&p, sp.params.data(), sizeof(p));
    if (row_offset != 0)
    {
        set_error("the aircraft device spawner must be the only spawner of its archetype");
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    const ComponentInfo& pc = s.comps[p.position];
    const ComponentInfo& tc = s.comps[p.target_altitude];
    if (pc.dtype != INCERTO_I16 || pc.lanes != 3 || tc.dtype != INCERTO_I16 || tc.lanes != 1 || p.airspace_size < 1 ||
        p.airspace_size > 32767 || p.airspace_height < 1 || p.airspace_height > 32767)
    {
        set_error("aircraft spawner: GridPosition3D must be 3-lane I16, target_altitude I16, airspace within 1..=32767");
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    AirArgs a;
    std::memset(&a, 0, sizeof(a));
    a.pos = static_cast<uint2*>(t.col(p.position)->d);
    a.n = p.num_aircraft;
    a.size = p.airspace_size;
    a.height = p.airspace_height;
    a.rk_spawn = make_round_keys(make_key(s.seed, s.replica, TAG_AIR_SPAWN));
    KTimer kt(s, "air_spawn", 11.0 * a.n);
    launch_air_spawn(a, static_cast<int16_t*>(t.col(p.target_altitude)->d), t.d_flags, s.stream);
    INC_CUDA(cudaGetLastError());
    return INCERTO_OK;
}

} // namespace incerto
