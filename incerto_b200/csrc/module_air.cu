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
static int32_t air_partition_setup(Sim& s);
static void air_fill(Sim& s, AirModule& am, AirArgs& a, Table& t, uint64_t step);

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
    INC_CUDA(dev_alloc(&am.d_acc, 64));
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
        INC_CUDA(dev_alloc(&am.d_slots, static_cast<size_t>(am.n_slots) * 8));
        INC_CUDA(cudaMemsetAsync(am.d_slots, 0, static_cast<size_t>(am.n_slots) * 8, s.stream));
        // pre-filter bitmap: >= 32 bits per aircraft (at most ~3 % of the bits set)
        // (one airspace split over the ranks: this rank sets bits for its own share of the aircraft only; twice the
        // average share keeps the false-positive rate if the split is uneven - the filter has no capacity to overflow)
        const uint64_t filter_n = s.part_world > 1 ? std::max<uint64_t>(2 * t.n / s.part_world, 1024) : t.n;
        uint32_t fb = 16;
        while ((1ull << fb) < 32 * filter_n && fb < 32) ++fb;
        am.filter_bits = fb;
        INC_CUDA(dev_alloc(&am.d_bits, (1ull << fb) / 8));
        INC_CUDA(dev_alloc(&am.d_bits_next, (1ull << fb) / 8));
        INC_CUDA(dev_alloc(&am.d_cand, std::max<uint64_t>(t.n, 2)));
    }
    INC_TRY(air_sort_rows(s));
    return air_partition_setup(s);
}

// ---- domain decomposition: ONE airspace over the GPUs of a box, split by x.
// Every rank spawns the whole population (same uids, same draws) and keeps the full table; the alive bit of a row means
// "owned by this rank" (grid x inside [x_lo, x_hi)).  Rows are sorted by (x, y) column at spawn, so a rank's aircraft
// are one contiguous run of rows plus, later, the few that wandered in.  After every step the ranks exchange one
// fixed-size mailbox per neighbour (AirMailbox, kernels_air.hpp): the aircraft that crossed the cut - uid and new
// position; the receiver revives its copy of the row, the sender marks its copy dead - and the positions of the
// aircraft standing in the boundary column, which the neighbour needs as ghosts for the +-x probes of check_conflicts.
// Draws are keyed by uid, so the union of the slabs is bit for bit the undivided world (tests/test_gpu_air_partition.py).
static AirMailbox* air_mailbox(AirModule& am, int which) // 0 out_lo, 1 out_hi, 2 in_lo, 3 in_hi
{
    return reinterpret_cast<AirMailbox*>(am.d_mail + static_cast<size_t>(which) * ((AirMailbox::bytes(am.mail_cap) + 255) & ~size_t(255)));
}

static int32_t air_partition_setup(Sim& s)
{
    AirModule& am = s.mods->air;
    am.part = false;
    if (s.part_world <= 1 || am.table < 0) return INCERTO_OK;
    if (am.grid < 0)
    {
        set_error("a row-partitioned air_traffic world needs the bounded 3-D spatial grid (its x extent is what gets split)");
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    const GridInst& g = s.grids[am.grid];
    const int32_t ext_x = g.bmax[0] - g.bmin[0] + 1;
    incerto_partition_range(ext_x, s.part_rank, s.part_world, &am.x_lo, &am.x_hi);
    if (am.x_hi - am.x_lo < 2)
    {
        set_error("air_traffic partition: every rank needs at least two columns of x (extent %d over %u ranks)", ext_x, s.part_world);
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    Table& t = s.tables[am.table];
    // a boundary column holds n / ext_x aircraft on average; room for 8 x that + 1024 per list
    const uint32_t cap = static_cast<uint32_t>(std::min<uint64_t>(8 * (t.n / static_cast<uint64_t>(ext_x) + 1) + 1024, std::max<uint64_t>(t.n, 1)));
    if (!am.d_mail || cap != am.mail_cap)
    {
        dev_free(am.d_mail);
        am.d_mail = nullptr;
        am.mail_cap = cap;
        const size_t one = (AirMailbox::bytes(cap) + 255) & ~size_t(255);
        INC_CUDA(dev_alloc(&am.d_mail, 4 * one));
    }
    const size_t one = (AirMailbox::bytes(am.mail_cap) + 255) & ~size_t(255);
    INC_CUDA(cudaMemsetAsync(am.d_mail, 0, 4 * one, s.stream));
    for (int k = 0; k < 4; ++k)
    {
        const uint32_t hdr[4] = {0u, 0u, am.mail_cap, 0u};
        INC_CUDA(cudaMemcpyAsync(air_mailbox(am, k), hdr, sizeof(hdr), cudaMemcpyHostToDevice, s.stream));
    }
    dev_free(am.d_row_of_uid);
    am.d_row_of_uid = nullptr;
    if (t.d_uid)
    {
        INC_CUDA(dev_alloc(&am.d_row_of_uid, std::max<uint64_t>(t.n_spawned, 1) * 4));
        launch_invert_permutation(t.d_uid, am.d_row_of_uid, t.n, s.stream);
    }
    am.part = true;
    t.all_alive = false;
    AirArgs a;
    air_fill(s, am, a, t, 0);
    launch_air_own_init(a, s.part_rank > 0 ? air_mailbox(am, 2) : nullptr,
                        s.part_rank + 1 < s.part_world ? air_mailbox(am, 3) : nullptr, s.stream);
    INC_CUDA(cudaGetLastError());
    am.bits_ready = false;
    return INCERTO_OK;
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
    INC_TRY(air_sort_rows(s));
    return air_partition_setup(s);
}

void air_free(Sim& s)
{
    AirModule& am = s.mods->air;
    dev_free(am.d_slots);
    dev_free(am.d_acc);
    dev_free(am.d_bits);
    dev_free(am.d_bits_next);
    am.d_bits_next = nullptr;
    dev_free(am.d_cand);
    dev_free(am.d_mail);
    dev_free(am.d_row_of_uid);
    am.d_mail = nullptr;
    am.d_row_of_uid = nullptr;
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
    a.part = am.part ? 1u : 0u;
    if (am.part)
    {
        a.x_lo = am.x_lo;
        a.x_hi = am.x_hi;
        a.flags = t.d_flags;
        a.flags_rw = t.d_flags;
        a.out_lo = s.part_rank > 0 ? air_mailbox(am, 0) : nullptr;
        a.out_hi = s.part_rank + 1 < s.part_world ? air_mailbox(am, 1) : nullptr;
        a.in_lo = s.part_rank > 0 ? air_mailbox(am, 2) : nullptr;
        a.in_hi = s.part_rank + 1 < s.part_world ? air_mailbox(am, 3) : nullptr;
        a.row_of_uid = am.d_row_of_uid;
    }
}

// the per-step exchange over NCCL (one fixed-size message per neighbour and direction), then the arrivals come alive
static int32_t air_exchange(Sim& s, AirModule& am, Table& t, uint64_t step)
{
    if (!am.part) return INCERTO_OK;
    AirArgs a;
    air_fill(s, am, a, t, step);
    if (s.comm)
    {
        const void* send_lo[1] = {air_mailbox(am, 0)};
        void* recv_lo[1] = {air_mailbox(am, 2)};
        const void* send_hi[1] = {air_mailbox(am, 1)};
        void* recv_hi[1] = {air_mailbox(am, 3)};
        const size_t bytes[1] = {AirMailbox::bytes(am.mail_cap)};
        KTimer kt(s, "air_exchange_nccl", 2.0 * bytes[0]);
        INC_TRY(comm_halo(s, send_lo, recv_lo, send_hi, recv_hi, bytes, 1));
        launch_air_unpack(a, const_cast<AirMailbox*>(a.in_lo), const_cast<AirMailbox*>(a.in_hi), s.stream);
        INC_CUDA(cudaGetLastError());
    }
    // without a communicator the host moves the mailboxes (incerto_sim_halo_export / _import; the import unpacks)
    return INCERTO_OK;
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
            if (am.part)
            {
                // the ghosts: what the neighbours sent after the previous step + what this rank handed over itself
                KTimer kt(s, "air_ghosts", 0.0);
                launch_air_ghosts(a, a.out_lo, a.out_hi, s.stream);
            }
            {
                KTimer kt(s, "air_candidates", 9.0 * t.n);
                launch_air_candidates(a, s.stream);
            }
        }
        if (am.part) launch_air_mail_reset(a.out_lo, a.out_hi, nullptr, nullptr, s.stream); // the step kernel refills them
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
        INC_TRY(air_exchange(s, am, t, step));
    }
    return INCERTO_OK;
}

// Manual transport of the mailboxes (single-process tests of the decomposition; with a communicator attached the
// exchange runs over NCCL inside run): export = this rank's outbox for `side` (0 lower, 1 upper neighbour), import =
// the neighbour's outbox into this rank's inbox, followed by the unpack of the arrivals.
int32_t air_mail_export(Sim& s, uint32_t side, void* out, size_t capacity, size_t* n_out)
{
    AirModule& am = s.mods->air;
    if (!am.part) return INCERTO_ERR_UNSUPPORTED;
    const size_t n = AirMailbox::bytes(am.mail_cap);
    if (n_out) *n_out = n;
    if (capacity < n) return INCERTO_ERR_INVALID_ARGUMENT;
    INC_CUDA(cudaMemcpyAsync(out, air_mailbox(am, side == 0 ? 0 : 1), n, cudaMemcpyDeviceToHost, s.stream));
    INC_CUDA(cudaStreamSynchronize(s.stream));
    return INCERTO_OK;
}
int32_t air_mail_import(Sim& s, uint32_t side, const void* in, size_t n)
{
    AirModule& am = s.mods->air;
    if (!am.part) return INCERTO_ERR_UNSUPPORTED;
    if (n != AirMailbox::bytes(am.mail_cap)) return INCERTO_ERR_INVALID_ARGUMENT;
    Table& t = s.tables[am.table];
    AirMailbox* box = air_mailbox(am, side == 0 ? 2 : 3);
    INC_CUDA(cudaMemcpyAsync(box, in, n, cudaMemcpyHostToDevice, s.stream));
    AirArgs a;
    air_fill(s, am, a, t, s.step);
    launch_air_unpack(a, side == 0 ? box : nullptr, side == 1 ? box : nullptr, s.stream);
    INC_CUDA(cudaGetLastError());
    INC_CUDA(cudaStreamSynchronize(s.stream));
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
