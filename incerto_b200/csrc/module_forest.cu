// Forest-fire module: dense pitched slab + overlay entities + row partition.
// Host side of kernels_forest.cu; see that file for the layout and draw contract.
//
// Reference: examples/forest_fire.rs (spawn :234-279, systems :282-365, stats :104-134).
#include <algorithm>
#include <cstdlib>
#include <thread>

#include "modules.hpp"

namespace incerto
{

static constexpr size_t kOvSlot = 4096; // bytes per overlay inbox slot (at most 4096 overlay entities)
static constexpr size_t kLead = 256; // bytes before column 0 / after the last column (reads of y-1 / y+1 bytes)

static void forest_partition(const Sim& s, ForestModule& fm)
{
    incerto_partition_range(fm.W, s.part_rank, s.part_world, &fm.x_begin, &fm.x_end);
}

static int owner_rank_of_column(const Sim& s, const ForestModule& fm, int32_t x)
{
    for (uint32_t r = 0; r < s.part_world; ++r)
    {
        int32_t b, e;
        incerto_partition_range(fm.W, r, s.part_world, &b, &e);
        if (x >= b && x < e) return static_cast<int>(r);
    }
    return -1;
}

static int32_t forest_alloc(Sim& s, ForestModule& fm)
{
    fm.pitch = forest_pitch(fm.H);
    const uint32_t slab_w = static_cast<uint32_t>(fm.x_end - fm.x_begin);
    const size_t bytes = static_cast<size_t>(forest_alloc_cols(static_cast<int32_t>(slab_w))) * fm.pitch;
    if (bytes >= (1ull << 32))
    {
        set_error("forest slab larger than 4 GiB per GPU is not supported");
        return INCERTO_ERR_UNSUPPORTED;
    }
    if (fm.d_alloc[0] && bytes == fm.slab_bytes) return INCERTO_OK;
    forest_free(s);
    fm.slab_bytes = bytes;
    for (int i = 0; i < 2; ++i)
    {
        INC_CUDA(dev_alloc(&fm.d_alloc[i], bytes + 2 * kLead));
        fm.d_slab[i] = fm.d_alloc[i] + kLead;
    }
    INC_CUDA(dev_alloc(&fm.d_stats_run, 16 * sizeof(uint64_t)));
    INC_CUDA(cudaMemsetAsync(fm.d_stats_run, 0, 16 * sizeof(uint64_t), s.stream));
    INC_CUDA(dev_alloc(&fm.d_stats, 8 * sizeof(uint64_t)));
    const size_t nov = std::max<uint32_t>(fm.n_overlay, 1);
    INC_CUDA(dev_alloc(&fm.d_ov_pos, nov * 2 * sizeof(int16_t)));
    for (int i = 0; i < 2; ++i) INC_CUDA(dev_alloc(&fm.d_ov[i], nov));
    // inboxes of the neighbours' overlay states: two slots each (the peer-memory transports alternate by step parity,
    // so that a neighbour one step ahead never overwrites what this rank still reads); NCCL only uses slot 0
    INC_CUDA(dev_alloc(&fm.d_ov_lo, 2 * kOvSlot));
    INC_CUDA(dev_alloc(&fm.d_ov_hi, 2 * kOvSlot));
    INC_CUDA(dev_alloc(&fm.d_ov_owner, nov));
    {
        uint32_t tc, tr;
        forest_tile_shape(&tc, &tr);
        const size_t tiles = static_cast<size_t>(fm.pitch / tr) * ((forest_alloc_cols(static_cast<int32_t>(slab_w)) - 2) / tc);
        INC_CUDA(dev_alloc(&fm.d_near_bits, (tiles / 32 + 1) * sizeof(uint32_t)));
    }
    return INCERTO_OK;
}

void forest_graph_free(ForestModule& fm);

void forest_peer_free(Sim& s)
{
    ForestModule& fm = s.mods->forest;
    for (void*& p : fm.ipc_opened)
    {
        if (p) cudaIpcCloseMemHandle(p);
        p = nullptr;
    }
    dev_release(fm.d_mail, false);
    fm.d_mail = nullptr;
    fm.peer = false;
    fm.peer_alloc_lo[0] = fm.peer_alloc_lo[1] = fm.peer_alloc_hi[0] = fm.peer_alloc_hi[1] = nullptr;
    fm.peer_ov_lo = fm.peer_ov_hi = nullptr;
    fm.peer_ovs_lo[0] = fm.peer_ovs_lo[1] = fm.peer_ovs_hi[0] = fm.peer_ovs_hi[1] = nullptr;
    fm.peer_mail_lo = fm.peer_mail_hi = nullptr;
    fm.pushes = 0;
    fm.rendezvous = 0;
}

// Map the neighbours' slabs through CUDA IPC (handles exchanged with one NCCL all-gather).  Any failure leaves the NCCL
// send/recv transport in place: `fm.peer` tells which one runs.
int32_t forest_peer_setup(Sim& s)
{
    ForestModule& fm = s.mods->forest;
    if (!fm.active || s.part_world <= 1 || !s.comm || !fm.d_alloc[0]) return INCERTO_OK;
    forest_peer_free(s);
    struct Pack
    {
        cudaIpcMemHandle_t alloc[2], ov[2], ov_lo, ov_hi, mail;
        uint32_t ok, slab_w, pad[2];
    };
    // INCERTO_B200_HALO = nccl | peer (separate wait / push kernels) | fused (default: inside the step kernel)
    const char* env = std::getenv("INCERTO_B200_HALO");
    const bool want = !(env && std::strcmp(env, "nccl") == 0);
    fm.fused = !(env && std::strcmp(env, "peer") == 0);
    Pack mine;
    std::memset(&mine, 0, sizeof(mine));
    INC_CUDA(dev_alloc(&fm.d_mail, 64));
    INC_CUDA(cudaMemsetAsync(fm.d_mail, 0, 64, s.stream));
    mine.ok = want && cudaIpcGetMemHandle(&mine.alloc[0], fm.d_alloc[0]) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.alloc[1], fm.d_alloc[1]) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.ov[0], fm.d_ov[0]) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.ov[1], fm.d_ov[1]) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.ov_lo, fm.d_ov_lo) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.ov_hi, fm.d_ov_hi) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.mail, fm.d_mail) == cudaSuccess;
    cudaGetLastError();
    mine.slab_w = static_cast<uint32_t>(fm.x_end - fm.x_begin);
    // all-gather the packs (device buffers), read them back
    const size_t all = sizeof(Pack) * s.part_world;
    INC_TRY(ensure_scratch(s, sizeof(Pack) + all + 256));
    uint8_t* d_send = s.d_scratch;
    uint8_t* d_recv = s.d_scratch + ((sizeof(Pack) + 255) & ~static_cast<size_t>(255));
    INC_CUDA(cudaMemcpyAsync(d_send, &mine, sizeof(Pack), cudaMemcpyHostToDevice, s.stream));
    INC_TRY(comm_allgather_bytes(s, d_send, d_recv, sizeof(Pack)));
    std::vector<Pack> packs(s.part_world);
    INC_CUDA(cudaMemcpyAsync(packs.data(), d_recv, all, cudaMemcpyDeviceToHost, s.stream));
    INC_CUDA(cudaStreamSynchronize(s.stream));
    bool ok = true;
    for (auto& p : packs) ok = ok && p.ok;
    int n_open = 0;
    auto open = [&](const cudaIpcMemHandle_t& h) -> void* {
        void* p = nullptr;
        if (!ok) return nullptr;
        if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
        {
            cudaGetLastError();
            ok = false;
            return nullptr;
        }
        fm.ipc_opened[n_open++] = p;
        return p;
    };
    if (s.part_rank > 0)
    {
        const Pack& p = packs[s.part_rank - 1];
        fm.peer_alloc_lo[0] = static_cast<uint8_t*>(open(p.alloc[0]));
        fm.peer_alloc_lo[1] = static_cast<uint8_t*>(open(p.alloc[1]));
        fm.peer_ov_lo = static_cast<uint8_t*>(open(p.ov_hi));
        fm.peer_ovs_lo[0] = static_cast<uint8_t*>(open(p.ov[0]));
        fm.peer_ovs_lo[1] = static_cast<uint8_t*>(open(p.ov[1]));
        fm.peer_mail_lo = static_cast<uint32_t*>(open(p.mail));
        fm.peer_slab_w_lo = p.slab_w;
    }
    if (s.part_rank + 1 < s.part_world)
    {
        const Pack& p = packs[s.part_rank + 1];
        fm.peer_alloc_hi[0] = static_cast<uint8_t*>(open(p.alloc[0]));
        fm.peer_alloc_hi[1] = static_cast<uint8_t*>(open(p.alloc[1]));
        fm.peer_ov_hi = static_cast<uint8_t*>(open(p.ov_lo));
        fm.peer_ovs_hi[0] = static_cast<uint8_t*>(open(p.ov[0]));
        fm.peer_ovs_hi[1] = static_cast<uint8_t*>(open(p.ov[1]));
        fm.peer_mail_hi = static_cast<uint32_t*>(open(p.mail));
    }
    // everybody must agree on the transport: a second tiny all-gather of the outcome
    uint32_t flag = ok ? 1u : 0u;
    INC_CUDA(cudaMemcpyAsync(d_send, &flag, 4, cudaMemcpyHostToDevice, s.stream));
    INC_TRY(comm_allgather_bytes(s, d_send, d_recv, 4));
    std::vector<uint32_t> flags(s.part_world);
    INC_CUDA(cudaMemcpyAsync(flags.data(), d_recv, 4 * s.part_world, cudaMemcpyDeviceToHost, s.stream));
    INC_CUDA(cudaStreamSynchronize(s.stream));
    for (uint32_t f : flags) ok = ok && f;
    if (!ok)
    {
        forest_peer_free(s);
        return INCERTO_OK; // NCCL send/recv it is
    }
    fm.peer = true;
    fm.pushes = 0;
    fm.rendezvous = 0;
    return INCERTO_OK;
}

void forest_free(Sim& s)
{
    ForestModule& fm = s.mods->forest;
    if (s.stream && fm.d_alloc[0]) cudaStreamSynchronize(s.stream); // the blocks go back to the cache: nothing may still use them
    forest_peer_free(s);
    forest_graph_free(fm);
    for (int i = 0; i < 2; ++i)
    {
        dev_release(fm.d_alloc[i], s.part_world <= 1); // slabs of a row-split grid were exported through CUDA IPC
        fm.d_alloc[i] = fm.d_slab[i] = nullptr;
        dev_release(fm.d_ov[i], s.part_world <= 1);
        fm.d_ov[i] = nullptr;
    }
    dev_release(fm.d_stats_run, s.part_world <= 1);
    dev_release(fm.d_stats, s.part_world <= 1);
    dev_release(fm.d_ov_pos, s.part_world <= 1);
    dev_release(fm.d_ov_lo, s.part_world <= 1);
    dev_release(fm.d_ov_hi, s.part_world <= 1);
    dev_release(fm.d_ov_owner, s.part_world <= 1);
    dev_release(fm.d_near_bits, s.part_world <= 1);
    fm.d_near_bits = nullptr;
    fm.d_stats_run = nullptr;
    fm.d_stats = nullptr;
    fm.d_ov_pos = nullptr;
    fm.d_ov_lo = fm.d_ov_hi = fm.d_ov_owner = nullptr;
    fm.slab_bytes = 0;
}

// upload overlay positions / initial states / owners
static int32_t forest_upload_overlays(Sim& s, ForestModule& fm, const std::vector<uint8_t>& states)
{
    fm.h_ov_owner.assign(fm.n_overlay, 0);
    for (uint32_t k = 0; k < fm.n_overlay; ++k)
    {
        const int o = owner_rank_of_column(s, fm, fm.h_ov_pos[2 * k]);
        uint8_t code = 3;
        if (o == static_cast<int>(s.part_rank))
            code = 0;
        else if (o + 1 == static_cast<int>(s.part_rank))
            code = 1;
        else if (o == static_cast<int>(s.part_rank) + 1)
            code = 2;
        fm.h_ov_owner[k] = code;
    }
    {
        // CTA tiles with an overlay entity within one cell (the kernel stages those overlays in shared memory)
        uint32_t tc, tr;
        forest_tile_shape(&tc, &tr);
        const uint32_t slab_w = static_cast<uint32_t>(fm.x_end - fm.x_begin);
        const uint32_t tiles_along_y = fm.pitch / tr;
        const size_t tiles = static_cast<size_t>(tiles_along_y) * ((forest_alloc_cols(static_cast<int32_t>(slab_w)) - 2) / tc);
        std::vector<uint32_t> bits(tiles / 32 + 1, 0u);
        for (uint32_t k = 0; k < fm.n_overlay; ++k)
            for (int dx = -1; dx <= 1; ++dx)
                for (int dy = -1; dy <= 1; ++dy)
                {
                    const int x = fm.h_ov_pos[2 * k] + dx, y = fm.h_ov_pos[2 * k + 1] + dy;
                    if (x < fm.x_begin || x >= fm.x_end || y < 0 || y >= fm.H) continue;
                    const size_t tile = static_cast<size_t>((x - fm.x_begin) / tc) * tiles_along_y + static_cast<uint32_t>(y) / tr;
                    bits[tile >> 5] |= 1u << (tile & 31);
                }
        INC_CUDA(cudaMemcpyAsync(fm.d_near_bits, bits.data(), bits.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s.stream));
        INC_CUDA(cudaStreamSynchronize(s.stream));
    }
    if (fm.n_overlay)
    {
        INC_CUDA(cudaMemcpyAsync(fm.d_ov_pos, fm.h_ov_pos.data(), fm.n_overlay * 2 * sizeof(int16_t),
                                 cudaMemcpyHostToDevice, s.stream));
        INC_CUDA(cudaMemcpyAsync(fm.d_ov_owner, fm.h_ov_owner.data(), fm.n_overlay, cudaMemcpyHostToDevice, s.stream));
        for (uint8_t* p : {fm.d_ov[0], fm.d_ov[1], fm.d_ov_lo, fm.d_ov_hi, fm.d_ov_lo + kOvSlot, fm.d_ov_hi + kOvSlot})
            INC_CUDA(cudaMemcpyAsync(p, states.data(), fm.n_overlay, cudaMemcpyHostToDevice, s.stream));
        INC_CUDA(cudaStreamSynchronize(s.stream));
    }
    return INCERTO_OK;
}

// FireStats of the freshly spawned slab (+ the overlay entities this rank owns): the step kernel
// keeps them current by counting transitions only.
static int32_t forest_init_stats(Sim& s, ForestModule& fm, const std::vector<uint8_t>& ov_states)
{
    const uint32_t slab_w = static_cast<uint32_t>(fm.x_end - fm.x_begin);
    INC_TRY(ensure_scratch(s, 4 * 8 * kNumSMs * 8 + 64));
    uint64_t* d_out = reinterpret_cast<uint64_t*>(s.d_scratch + 4 * 8 * kNumSMs * 8);
    {
        KTimer kt(s, "forest_stats", 1.0 * slab_w * fm.H);
        launch_forest_stats(fm.d_slab[fm.cur], fm.H, static_cast<int32_t>(slab_w), d_out, s.d_scratch, s.d_ticket, s.stream);
    }
    uint64_t st[5];
    INC_CUDA(cudaMemcpyAsync(st, d_out, sizeof(st), cudaMemcpyDeviceToHost, s.stream));
    INC_CUDA(cudaStreamSynchronize(s.stream));
    for (uint32_t k = 0; k < fm.n_overlay; ++k)
    {
        if (fm.h_ov_owner[k] != 0) continue;
        const uint8_t v = ov_states[k];
        if (v & 0x40)
            ++st[0];
        else if (v & 0x80)
            ++st[3];
        else if (v & 0x3f)
            ++st[1];
        else
            ++st[2];
        ++st[4];
    }
    // totals, last published step (none), two empty delta slots (layout: kernels_forest.cu, forest_publish)
    const uint64_t run[12] = {st[0], st[1], st[3], 0, 0, 0, 0, 0, 0, 0, 0, 0};
    INC_CUDA(cudaMemcpyAsync(fm.d_stats_run, run, sizeof(run), cudaMemcpyHostToDevice, s.stream));
    fm.unpublished_step = 0;
    INC_CUDA(cudaMemcpyAsync(fm.d_stats, st, sizeof(st), cudaMemcpyHostToDevice, s.stream));
    INC_CUDA(cudaStreamSynchronize(s.stream));
    fm.stats_valid = true;
    return INCERTO_OK;
}

// INCERTO_SPAWN_FOREST_GRID: examples/forest_fire.rs:234-279 restated with the
// Philox contract: cell uid = x*H + y is Healthy iff word (uid&3) of
// philox(FOREST_SPAWN, (uid>>2,0,0,0)) < floor(density*2^32); igniter k sits at
// index mulhi32(word 0 of philox(FOREST_IGNITE, (k,0,0,0)), W*H) of the x-major
// position list (:261-263) and gets uid W*H + k.
int32_t forest_spawn_device(Sim& s, const HostSpawn& sp)
{
    ForestModule& fm = s.mods->forest;
    incerto_spawn_forest_params p;
    std::memcpy(&p, sp.params.data(), sizeof(p));
    if (p.width <= 0 || p.height <= 0 || p.width > 32767 || p.height > 32767 || p.burn_duration < 1 ||
        p.burn_duration > 63 || p.initial_fire_count > 256 ||
        static_cast<uint64_t>(p.width) * p.height + p.initial_fire_count >= (1ull << 32))
    {
        set_error("forest spawner: bad grid size / burn duration / fire count");
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    fm.comp_pos = p.position;
    fm.comp_cell = p.cell;
    fm.W = p.width;
    fm.H = p.height;
    fm.n_overlay = p.initial_fire_count;
    forest_partition(s, fm);
    INC_TRY(forest_alloc(s, fm));
    for (int i = 0; i < 2; ++i) INC_CUDA(cudaMemsetAsync(fm.d_alloc[i], 0, fm.slab_bytes + 2 * kLead, s.stream));
    const PhiloxKey key = make_key(s.seed, s.replica, TAG_FOREST_SPAWN);
    // own columns plus the in-grid halo columns (a pure function of uid, so no exchange is needed)
    const int32_t lo = std::max(fm.x_begin - 1, 0), hi = std::min(fm.x_end + 1, fm.W);
    launch_forest_spawn(fm.d_slab[0] + static_cast<size_t>(lo - (fm.x_begin - 1)) * fm.pitch - fm.pitch, fm.W, fm.H,
                        lo, hi, key, bernoulli_threshold(p.initial_forest_density), s.stream);
    INC_CUDA(cudaGetLastError());
    fm.cur = 0;
    fm.h_ov_pos.resize(2 * static_cast<size_t>(fm.n_overlay));
    const PhiloxKey kign = make_key(s.seed, s.replica, TAG_FOREST_IGNITE);
    const uint32_t cells = static_cast<uint32_t>(fm.W) * static_cast<uint32_t>(fm.H);
    for (uint32_t k = 0; k < fm.n_overlay; ++k)
    {
        const uint4 r = philox4x32_10(k, 0u, 0u, 0u, kign);
        const uint32_t idx = uniform_below(r.x, cells);
        fm.h_ov_pos[2 * k] = static_cast<int16_t>(idx / fm.H);
        fm.h_ov_pos[2 * k + 1] = static_cast<int16_t>(idx % fm.H);
    }
    std::vector<uint8_t> st(std::max<uint32_t>(fm.n_overlay, 1), static_cast<uint8_t>(p.burn_duration));
    INC_TRY(forest_upload_overlays(s, fm, st));
    INC_TRY(forest_init_stats(s, fm, st));
    Table& t = s.tables[fm.table];
    t.dense_w = fm.W;
    t.dense_h = fm.H;
    t.x_begin = fm.x_begin;
    t.x_end = fm.x_end;
    t.all_alive = true;
    return INCERTO_OK;
}

// OR of fn(begin, end) over [0, n) split into contiguous ranges, one per host thread (small inputs stay on the caller)
template <class Fn>
static uint32_t host_parallel_or(uint64_t n, Fn fn)
{
    // one process per GPU is the rule on a multi-GPU box: share the host cores with the other local ranks
    unsigned hw = std::thread::hardware_concurrency();
    if (const char* lw = std::getenv("LOCAL_WORLD_SIZE"))
        if (std::atoi(lw) > 1) hw /= static_cast<unsigned>(std::atoi(lw));
    const uint64_t nt = std::min<uint64_t>(std::min<unsigned>(hw ? hw : 1u, 8u), n / (1u << 20));
    if (nt <= 1) return fn(0, n);
    std::vector<uint32_t> res(nt, 0u);
    std::vector<std::thread> th;
    for (uint64_t i = 0; i < nt; ++i)
        th.emplace_back([&, i] { res[i] = fn(n * i / nt, n * (i + 1) / nt); });
    uint32_t r = 0;
    for (uint64_t i = 0; i < nt; ++i)
    {
        th[i].join();
        r |= res[i];
    }
    return r;
}

// Host-spawned forest: find the dense x-major prefix among the spawned rows and
// move the cell column into the pitched slab.
static int32_t forest_adopt_host_table(Sim& s, ForestModule& fm)
{
    stage_mark("(before forest adopt)");
    Table& t = s.tables[fm.table];
    const ComponentInfo& pc = s.comps[fm.comp_pos];
    const ComponentInfo& cc = s.comps[fm.comp_cell];
    if (pc.lanes != 2 || (pc.dtype != INCERTO_I16 && pc.dtype != INCERTO_U16 && pc.dtype != INCERTO_I32) ||
        cc.dtype != INCERTO_U8 || cc.lanes != 1)
    {
        set_error("forest: position must be a 2-lane I16/I32 component and cell a U8 component");
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    // W, H come from the spatial grid bounds (forest_fire.rs:151-157)
    const GridInst* grid = nullptr;
    for (auto& g : s.grids)
        if (g.position == fm.comp_pos && g.dim == 2) grid = &g;
    if (!grid || !grid->bounded || grid->bmin[0] != 0 || grid->bmin[1] != 0)
    {
        set_error("forest systems need add_spatial_grid(position, cell) with bounds (0,0)..(W-1,H-1)");
        return INCERTO_ERR_INVALID_ARGUMENT;
    }
    fm.W = grid->bmax[0] + 1;
    fm.H = grid->bmax[1] + 1;
    const uint64_t cells = static_cast<uint64_t>(fm.W) * fm.H;
    // Walk the host spawns of this archetype in spawn order without copying them: rows [0, cells) must be the
    // dense x-major grid (forest_fire.rs:239-257), the rest are overlay entities.
    struct Piece
    {
        const uint8_t* pos;
        const uint8_t* st;
        uint64_t n;
    };
    std::vector<Piece> pieces;
    uint64_t total = 0;
    for (auto& sp : s.spawns)
    {
        if (sp.device_kind) continue;
        const HostSpawn::Col* pcol = nullptr;
        const HostSpawn::Col* ccol = nullptr;
        size_t ndata = 0;
        for (auto& c : sp.cols)
        {
            if (c.comp == fm.comp_pos) pcol = &c;
            if (c.comp == fm.comp_cell) ccol = &c;
            if (s.comps[c.comp].dtype != INCERTO_MARKER) ++ndata;
        }
        if (!pcol || !ccol || ndata != 2) continue;
        pieces.push_back(Piece{pcol->bytes(), ccol->bytes(), sp.n});
        total += sp.n;
    }
    // A row-partitioned handle may be given just the rows it needs - the columns [lo, hi) of its slab plus one halo
    // column on each interior side, still x-major, followed by ALL overlay entities - instead of the whole world:
    // `row0` is the global row (uid) of the first dense row this handle was handed.
    uint64_t row0 = 0, dense_rows = cells;
    if (s.part_world > 1 && total < cells)
    {
        int32_t xb = 0, xe = 0;
        incerto_partition_range(fm.W, s.part_rank, s.part_world, &xb, &xe);
        const int32_t lo = std::max(xb - 1, 0), hi = std::min(xe + 1, fm.W);
        row0 = static_cast<uint64_t>(lo) * fm.H;
        dense_rows = static_cast<uint64_t>(hi - lo) * fm.H;
    }
    if (total != t.n_spawned || total < dense_rows)
    {
        set_error("forest: the spawned cells do not cover the %dx%d grid%s", fm.W, fm.H,
                  s.part_world > 1 ? " (nor this rank's slab + halo columns)" : "");
        return INCERTO_ERR_UNSUPPORTED;
    }
    const bool wide = pc.dtype == INCERTO_I32;
    auto pos_at = [&](const Piece& pz, uint64_t r, int32_t& x, int32_t& y) {
        if (wide)
        {
            const int32_t* q = reinterpret_cast<const int32_t*>(pz.pos);
            x = q[2 * r];
            y = q[2 * r + 1];
        }
        else
        {
            const int16_t* q = reinterpret_cast<const int16_t*>(pz.pos);
            x = q[2 * r];
            y = q[2 * r + 1];
        }
    };
    // The dense prefix is checked column by column with branch-free loops the host compiler vectorises (a 4096^2
    // grid is 16.8 M rows: this check is on the end-to-end path of every host-spawned job): the first row of
    // column x must be (x, 0) and every following row must add exactly one to y.
    std::vector<int32_t> ov_x, ov_y;
    const uint8_t* dense_states = nullptr; // the states of the dense prefix, when one spawn call holds them all
    std::vector<uint8_t> st;               // otherwise (several pieces): gathered here
    if (!pieces.empty() && pieces[0].n >= dense_rows)
        dense_states = pieces[0].st;
    else
        st.resize(dense_rows);
    std::vector<uint8_t> ov_states;
    // pass A (cheap): the rows after the dense prefix are overlay entities
    {
        uint64_t g = 0; // global row
        for (auto& pz : pieces)
        {
            const uint64_t n_dense = g < dense_rows ? std::min<uint64_t>(pz.n, dense_rows - g) : 0;
            for (uint64_t r = n_dense; r < pz.n; ++r)
            {
                int32_t x, y;
                pos_at(pz, r, x, y);
                if (x < 0 || x >= fm.W || y < 0 || y >= fm.H)
                {
                    set_error("entity at position (%d, %d) outside spatial grid bounds", x, y); // spatial_grid.rs:276-279
                    return INCERTO_ERR_OUT_OF_BOUNDS;
                }
                ov_x.push_back(x);
                ov_y.push_back(y);
                ov_states.push_back(pz.st[r]);
            }
            g += pz.n;
        }
    }
    // pass B (streams every row: 5 bytes per cell): order of the dense prefix and validity of the state codes.
    // Returns 0, 1 (order) or 2 (state code); runs beside the allocation + upload below when the states can be
    // uploaded straight from the caller's buffer.
    auto verify = [&]() -> int {
        uint64_t g = 0; // global row
        uint32_t bad_state = 0;
        for (auto& pz : pieces)
        {
            const uint64_t n_dense = g < dense_rows ? std::min<uint64_t>(pz.n, dense_rows - g) : 0;
            if (n_dense)
            {
                if (!dense_states) std::memcpy(st.data() + g, pz.st, n_dense);
                uint32_t bad = 0;
                if (wide)
                {
                    const int32_t* q = reinterpret_cast<const int32_t*>(pz.pos);
                    for (uint64_t r = 0; r < n_dense; ++r)
                    {
                        const uint64_t gg = row0 + g + r;
                        bad |= static_cast<uint32_t>(q[2 * r] != static_cast<int32_t>(gg / fm.H)) |
                               static_cast<uint32_t>(q[2 * r + 1] != static_cast<int32_t>(gg % fm.H));
                    }
                }
                else
                {
                    // packed (x | y << 16): within a column consecutive rows differ by 0x10000.  The rows are split
                    // over a few host threads.
                    const uint32_t* q = reinterpret_cast<const uint32_t*>(pz.pos);
                    auto check_rows = [&](uint64_t r, uint64_t r_end) {
                        uint32_t b = 0;
                        while (r < r_end)
                        {
                            const uint64_t gg = row0 + g + r;
                            const uint32_t x = static_cast<uint32_t>(gg / fm.H), y = static_cast<uint32_t>(gg % fm.H);
                            const uint64_t run = std::min<uint64_t>(r_end - r, static_cast<uint64_t>(fm.H) - y);
                            b |= static_cast<uint32_t>(q[r] != ((x & 0xffffu) | (y << 16)));
                            uint32_t acc = 0;
                            for (uint64_t k = 1; k < run; ++k) acc |= (q[r + k] - q[r + k - 1]) ^ 0x10000u;
                            b |= acc;
                            r += run;
                        }
                        return b;
                    };
                    bad |= host_parallel_or(n_dense, check_rows);
                }
                if (bad) return 1;
            }
            // state codes: BURNED 0, BURNING 1..63, HEALTHY 0x40, EMPTY 0x80
            {
                const uint8_t* stp = pz.st;
                bad_state |= host_parallel_or(pz.n, [stp](uint64_t r, uint64_t r_end) {
                    uint32_t b = 0;
                    for (; r < r_end; ++r)
                    {
                        const uint8_t v = stp[r];
                        b |= static_cast<uint32_t>(!(v <= 63 || v == INCERTO_CELL_HEALTHY || v == INCERTO_CELL_EMPTY));
                    }
                    return b;
                });
            }
            g += pz.n;
        }
        return bad_state ? 2 : 0;
    };
    auto verdict = [&](int v) -> int32_t {
        if (v == 1)
        {
            set_error("forest: cells must be spawned in x-major order (for x {for y}) as forest_fire.rs:239-257 does");
            return INCERTO_ERR_UNSUPPORTED;
        }
        if (v == 2)
        {
            set_error("forest: invalid cell state code");
            return INCERTO_ERR_INVALID_ARGUMENT;
        }
        return INCERTO_OK;
    };
    int verified = -1;
    std::thread checker;
    if (dense_states)
        checker = std::thread([&] { verified = verify(); });
    else
        INC_TRY(verdict(verify()));
    struct Joiner
    {
        std::thread& t;
        ~Joiner()
        {
            if (t.joinable()) t.join();
        }
    } joiner{checker};
    stage_mark("forest adopt: host checks");
    const uint8_t* dense = dense_states ? dense_states : st.data();
    fm.n_overlay = static_cast<uint32_t>(std::min<uint64_t>(total - dense_rows, 0xffffffffu));
    if (fm.n_overlay > kOvSlot)
    {
        set_error("forest: at most 4096 overlay entities on top of the dense grid");
        return INCERTO_ERR_UNSUPPORTED;
    }
    forest_partition(s, fm);
    INC_TRY(forest_alloc(s, fm));
    for (int i = 0; i < 2; ++i) INC_CUDA(cudaMemsetAsync(fm.d_alloc[i], 0, fm.slab_bytes + 2 * kLead, s.stream));
    const int32_t lo = std::max(fm.x_begin - 1, 0), hi = std::min(fm.x_end + 1, fm.W);
    stage_mark("forest adopt: alloc, memset");
    INC_CUDA(cudaMemcpy2DAsync(fm.d_slab[0] + static_cast<size_t>(lo - (fm.x_begin - 1)) * fm.pitch, fm.pitch,
                               dense + (static_cast<size_t>(lo) * fm.H - row0), fm.H, fm.H, hi - lo, cudaMemcpyHostToDevice,
                               s.stream));
    INC_CUDA(cudaStreamSynchronize(s.stream));
    if (checker.joinable())
    {
        checker.join();
        INC_TRY(verdict(verified));
    }
    stage_mark("forest adopt: H2D (+ verify)");
    fm.cur = 0;
    fm.h_ov_pos.resize(2 * static_cast<size_t>(fm.n_overlay));
    std::vector<uint8_t> ovst(std::max<uint32_t>(fm.n_overlay, 1), 0);
    for (uint32_t k = 0; k < fm.n_overlay; ++k)
    {
        fm.h_ov_pos[2 * k] = static_cast<int16_t>(ov_x[k]);
        fm.h_ov_pos[2 * k + 1] = static_cast<int16_t>(ov_y[k]);
        ovst[k] = ov_states[k];
    }
    INC_TRY(forest_upload_overlays(s, fm, ovst));
    // the generic columns are no longer needed
    for (auto& c : t.cols)
    {
        dev_free(c.d);
        c.d = nullptr;
    }
    dev_free(t.d_flags);
    t.d_flags = nullptr;
    t.dense_w = fm.W;
    t.dense_h = fm.H;
    t.x_begin = fm.x_begin;
    t.x_end = fm.x_end;
    stage_mark("forest adopt: overlays, frees");
    INC_TRY(forest_init_stats(s, fm, ovst));
    stage_mark("forest adopt: init stats");
    return INCERTO_OK;
}

int32_t forest_bind(Sim& s)
{
    ForestModule& fm = s.mods->forest;
    bool first = true;
    for (auto& sys : s.systems)
    {
        if (sys.kind < INCERTO_SYS_FOREST_SPREAD || sys.kind > INCERTO_SYS_FOREST_REGROWTH) continue;
        if (sys.params.size() != sizeof(incerto_forest_params)) return INCERTO_ERR_INVALID_ARGUMENT;
        incerto_forest_params p;
        std::memcpy(&p, sys.params.data(), sizeof(p));
        if (p.position < 0 || p.cell < 0 || p.position >= static_cast<int>(s.comps.size()) ||
            p.cell >= static_cast<int>(s.comps.size()))
            return INCERTO_ERR_INVALID_ARGUMENT;
        if (first)
        {
            if (fm.comp_pos >= 0 && (fm.comp_pos != p.position || fm.comp_cell != p.cell))
            {
                set_error("forest systems and the forest spawner name different components");
                return INCERTO_ERR_INVALID_ARGUMENT;
            }
            fm.comp_pos = p.position;
            fm.comp_cell = p.cell;
            first = false;
        }
        else if (p.position != fm.comp_pos || p.cell != fm.comp_cell)
        {
            set_error("forest systems must agree on their components");
            return INCERTO_ERR_INVALID_ARGUMENT;
        }
        if (sys.kind == INCERTO_SYS_FOREST_SPREAD)
        {
            fm.do_spread = true;
            fm.p_spread = p.probability;
            if (p.burn_duration < 1 || p.burn_duration > 63)
            {
                set_error("forest: burn_duration must be in 1..=63");
                return INCERTO_ERR_INVALID_ARGUMENT;
            }
            fm.burn_duration = p.burn_duration;
        }
        else if (sys.kind == INCERTO_SYS_FOREST_BURN_PROGRESSION)
            fm.do_progress = true;
        else
        {
            fm.do_regrow = true;
            fm.p_regrow = p.probability;
        }
    }
    s.comps[fm.comp_pos].exists = true;
    s.comps[fm.comp_cell].exists = true;
    if (fm.table < 0)
    {
        fm.table = table_for_components(s, {fm.comp_pos, fm.comp_cell});
        if (fm.table < 0)
        {
            set_error("forest systems: no entities carry (position, cell)");
            return INCERTO_ERR_UNSUPPORTED;
        }
        INC_TRY(forest_adopt_host_table(s, fm));
    }
    fm.active = true;
    return INCERTO_OK;
}

// reset() with the peer-memory halo transport: before this rank respawns, the boundary columns its neighbours pushed
// after their last step must have landed (they target the buffers about to be rewritten) ...
int32_t forest_peer_quiesce(Sim& s)
{
    ForestModule& fm = s.mods->forest;
    if (!fm.peer || !s.comm || !fm.pushes) return INCERTO_OK;
    if (fm.fused) // the neighbours announce their steps in their own mailboxes
        launch_forest_stamp_wait(fm.peer_mail_lo, fm.peer_mail_hi, fm.pushes, s.d_error, s.stream);
    else
        launch_forest_halo_wait(fm.d_mail, fm.peer_mail_lo ? fm.pushes : 0u, fm.peer_mail_hi ? fm.pushes : 0u, s.d_error, s.stream);
    INC_CUDA(cudaStreamSynchronize(s.stream));
    return INCERTO_OK;
}

int32_t forest_cold_rendezvous(Sim& s)
{
    ForestModule& fm = s.mods->forest;
    if (!fm.active || !fm.peer || !fm.fused || !s.comm) return INCERTO_OK;
    launch_forest_rendezvous(fm.d_mail, fm.peer_mail_lo, fm.peer_mail_hi, ++fm.rendezvous, s.d_error, s.stream);
    INC_CUDA(cudaGetLastError());
    return INCERTO_OK;
}

int32_t forest_reset_after_spawn(Sim& s)
{
    ForestModule& fm = s.mods->forest;
    forest_graph_free(fm); // the replica key is baked into the captured kernel parameters
    fm.d_step_mirror = ~0ull; // reset rewrote the device StepNumber
    Table& t = s.tables[fm.table];
    bool device_spawned = false;
    for (auto& sp : s.spawns)
        if (sp.device_kind && sp.kind == INCERTO_SPAWN_FOREST_GRID) device_spawned = true;
    if (!device_spawned) INC_TRY(forest_adopt_host_table(s, fm));
    (void)t;
    if (fm.peer && s.comm)
    {
        // ... and nobody may step (and push into a neighbour) before every rank has respawned: a collective as barrier
        INC_TRY(ensure_scratch(s, 64));
        INC_CUDA(cudaMemsetAsync(s.d_scratch, 0, 8, s.stream));
        INC_TRY(comm_allreduce(s, s.d_scratch, 1, 0, 0));
        INC_CUDA(cudaStreamSynchronize(s.stream));
    }
    return INCERTO_OK;
}

int32_t forest_halo_exchange(Sim& s)
{
    ForestModule& fm = s.mods->forest;
    if (s.part_world <= 1) return INCERTO_OK;
    if (!s.comm) return INCERTO_OK; // manual halo transport: the host moves the columns (incerto_sim_halo_*)
    // `cur` already points at the buffer the step just produced
    uint8_t* slab = fm.d_slab[fm.cur];
    const uint32_t slab_w = static_cast<uint32_t>(fm.x_end - fm.x_begin);
    if (fm.peer && fm.fused) return INCERTO_OK; // the step kernel did it (forest_launch_one)
    if (fm.peer)
    {
        // peer-memory transport: my boundary columns go straight into the neighbours' halo columns of the same
        // buffer (all ranks flip `cur` in lock step), then the step stamp lands in their mailboxes
        uint8_t* dst_lo = fm.peer_alloc_lo[fm.cur] ? fm.peer_alloc_lo[fm.cur] + kLead + static_cast<size_t>(fm.peer_slab_w_lo + 1) * fm.pitch : nullptr;
        uint8_t* dst_hi = fm.peer_alloc_hi[fm.cur] ? fm.peer_alloc_hi[fm.cur] + kLead : nullptr;
        fm.pushes += 1;
        KTimer kt(s, "forest_halo_push_peer", 2.0 * fm.pitch);
        launch_forest_halo_push(slab + static_cast<size_t>(1) * fm.pitch, dst_lo, slab + static_cast<size_t>(slab_w) * fm.pitch, dst_hi,
                                fm.pitch, fm.d_ov[fm.cur], fm.peer_ov_lo ? fm.peer_ov_lo + (fm.pushes & 1u) * kOvSlot : nullptr,
                                fm.peer_ov_hi ? fm.peer_ov_hi + (fm.pushes & 1u) * kOvSlot : nullptr, fm.n_overlay,
                                fm.peer_mail_lo, fm.peer_mail_hi, fm.pushes, s.stream);
        INC_CUDA(cudaGetLastError());
        return INCERTO_OK;
    }
    KTimer kt(s, "forest_halo_nccl", 2.0 * fm.pitch);
    const void* send_lo[2] = {slab + static_cast<size_t>(1) * fm.pitch, fm.d_ov[fm.cur]};
    void* recv_lo[2] = {slab, fm.d_ov_lo};
    const void* send_hi[2] = {slab + static_cast<size_t>(slab_w) * fm.pitch, fm.d_ov[fm.cur]};
    void* recv_hi[2] = {slab + static_cast<size_t>(slab_w + 1) * fm.pitch, fm.d_ov_hi};
    const size_t bytes[2] = {fm.pitch, std::max<uint32_t>(fm.n_overlay, 1)};
    return comm_halo(s, send_lo, recv_lo, send_hi, recv_hi, bytes, fm.n_overlay ? 2 : 1);
}

// One step.  `step` is the StepNumber of the launch when `graph_offset` < 0; inside a captured graph the kernel reads
// the device counter and adds `graph_offset`.
static int32_t forest_launch_one(Sim& s, uint64_t step, int graph_offset, ForestArgs* args_out = nullptr)
{
    ForestModule& fm = s.mods->forest;
    const uint32_t slab_w = static_cast<uint32_t>(fm.x_end - fm.x_begin);
    ForestArgs a;
    std::memset(&a, 0, sizeof(a));
    a.cur = fm.d_slab[fm.cur];
    a.next = fm.d_slab[1 - fm.cur];
    a.W = fm.W;
    a.H = fm.H;
    a.x_begin = fm.x_begin;
    a.x_end = fm.x_end;
    a.n_overlay = fm.n_overlay;
    a.ov_pos = fm.d_ov_pos;
    a.near_bits = fm.d_near_bits;
    a.ov_cur = fm.d_ov[fm.cur];
    a.ov_next = fm.d_ov[1 - fm.cur];
    a.ov_uid_base = static_cast<uint32_t>(fm.W) * static_cast<uint32_t>(fm.H);
    a.ov_owned = static_cast<uint32_t>(std::count(fm.h_ov_owner.begin(), fm.h_ov_owner.end(), uint8_t(0)));
    a.do_spread = fm.do_spread;
    a.do_progress = fm.do_progress;
    a.do_regrow = fm.do_regrow;
    a.thr_spread = bernoulli_threshold(fm.p_spread);
    a.thr_regrow = bernoulli_threshold(fm.p_regrow);
    a.regrow_two_level = (fm.p_regrow * 256.0 <= 1.0) ? 1u : 0u;
    a.thr_regrow_l2 = bernoulli_threshold(fm.p_regrow * 256.0);
    a.burn_duration = fm.burn_duration;
    a.key_spread = make_key(s.seed, s.replica, TAG_FOREST_SPREAD);
    a.key_regrow = make_key(s.seed, s.replica, TAG_FOREST_REGROW);
    a.l2_resident = (graph_offset >= 0 && 2 * fm.slab_bytes <= (96u << 20)) ? 1u : 0u; // 126 MB of L2, some of it for the rest
    a.d_step = graph_offset >= 0 ? s.mods->d_step : nullptr;
    a.step_base = graph_offset >= 0 ? static_cast<uint32_t>(graph_offset) : static_cast<uint32_t>(step);
    a.stats_run = fm.d_stats_run;
    a.error = s.d_error;
    a.ticket = s.d_ticket;
    a.stats_out = fm.d_stats;
    if (fm.series >= 0)
    {
        SeriesInst& se = s.series[fm.series];
        a.series_values = se.d_values;
        a.series_counts = se.d_counts;
        a.series_interval = se.interval;
        a.series_capacity = se.capacity;
    }
    // the neighbours' overlay states of the previous step: the peer-memory transports alternate two inbox slots
    const size_t ov_slot = (fm.peer && s.comm) ? (fm.pushes & 1u) * kOvSlot : 0;
    ForestPeer pr;
    const bool fused = fm.peer && fm.fused && s.comm;
    const uint8_t* ov_lo = fm.d_ov_lo + ov_slot;
    const uint8_t* ov_hi = fm.d_ov_hi + ov_slot;
    if (fused)
    {
        // pull protocol (kernels.hpp: ForestPeer): this step reads the neighbours' boundary columns and overlay states
        // straight from the buffers their previous step wrote
        pr.my_mail = fm.d_mail;
        pr.need_lo = fm.peer_mail_lo ? fm.pushes : 0u;
        pr.need_hi = fm.peer_mail_hi ? fm.pushes : 0u;
        pr.stamp = fm.pushes; // this launch announces the step before it; forest_finalize the last one of a run
        fm.pushes += 1;
        pr.src_lo = fm.peer_alloc_lo[fm.cur] ? fm.peer_alloc_lo[fm.cur] + kLead + static_cast<size_t>(fm.peer_slab_w_lo) * fm.pitch : nullptr;
        pr.src_hi = fm.peer_alloc_hi[fm.cur] ? fm.peer_alloc_hi[fm.cur] + kLead + fm.pitch : nullptr;
        pr.mail_lo = fm.peer_mail_lo;
        pr.mail_hi = fm.peer_mail_hi;
        if (const char* dbg = std::getenv("INCERTO_B200_DEBUG_HALO")) // timing experiments only (results are wrong)
        {
            if (dbg[0] == 's') pr.my_mail = nullptr, pr.need_lo = pr.need_hi = 0; // no stamps, no pulls
            if (dbg[0] == 'z') pr.debug = 1; // natural launch order
            if (dbg[0] == 'q') pr.debug = 2; // boundary-first order and stamps, but no pulls
            if (dbg[0] == 'l') // the same code path on LOCAL memory: my own stamp and columns instead of the neighbours'
            {
                pr.mail_lo = pr.mail_hi = fm.d_mail;
                if (pr.src_lo) pr.src_lo = fm.d_slab[fm.cur] + fm.pitch;
                if (pr.src_hi) pr.src_hi = fm.d_slab[fm.cur] + fm.pitch;
            }
            if (dbg[0] == 'm') pr.mail_lo = pr.mail_hi = fm.d_mail; // local stamps, remote pulls
            if (dbg[0] == 'c') // remote stamps, local pulls
            {
                if (pr.src_lo) pr.src_lo = fm.d_slab[fm.cur] + fm.pitch;
                if (pr.src_hi) pr.src_hi = fm.d_slab[fm.cur] + fm.pitch;
            }
        }
        if (pr.need_lo && fm.peer_ovs_lo[fm.cur]) ov_lo = fm.peer_ovs_lo[fm.cur];
        if (pr.need_hi && fm.peer_ovs_hi[fm.cur]) ov_hi = fm.peer_ovs_hi[fm.cur];
    }
    {
        // 2 B per cell-update (1 B read + 1 B written)
        KTimer kt(s, fused ? "forest_step_fused_halo" : "forest_step", 2.0 * slab_w * fm.H);
        // captured runs of steps (one GPU): every step is a programmatic dependent of the one before it
        const bool pdl = graph_offset >= 0 && !fused && fm.graph_pdl;
        INC_CUDA(launch_forest_step_ex(a, ov_lo, ov_hi, fm.d_ov_owner, s.stream, fused ? &pr : nullptr, pdl));
    }
    fm.cur = 1 - fm.cur;
    if (args_out) *args_out = a;
    if (graph_offset < 0) fm.unpublished_step = step; // on record once the next step starts or forest_finalize ran
    return INCERTO_OK;
}

// The FireStats record / series slot of a step are written by the first thread of the NEXT step's grid (off the
// step's critical path; kernels_forest.cu, forest_publish).  After the last step of a run somebody has to do it:
// called at the end of every run()/run_async() (inside the timed region) and at the end of run_cold().
int32_t forest_finalize(Sim& s)
{
    ForestModule& fm = s.mods->forest;
    if (!fm.active || !fm.unpublished_step || !fm.d_alloc[0]) return INCERTO_OK;
    ForestArgs a;
    std::memset(&a, 0, sizeof(a));
    a.H = fm.H;
    a.x_begin = fm.x_begin;
    a.x_end = fm.x_end;
    a.ov_owned = static_cast<uint32_t>(std::count(fm.h_ov_owner.begin(), fm.h_ov_owner.end(), uint8_t(0)));
    a.stats_run = fm.d_stats_run;
    a.stats_out = fm.d_stats;
    a.step_base = static_cast<uint32_t>(fm.unpublished_step);
    if (fm.series >= 0)
    {
        SeriesInst& se = s.series[fm.series];
        a.series_values = se.d_values;
        a.series_counts = se.d_counts;
        a.series_interval = se.interval;
        a.series_capacity = se.capacity;
    }
    {
        // row-split grid with the fused halo transport: the neighbours learn that the run's last step is complete
        ForestPeer pr;
        if (fm.peer && fm.fused && s.comm && fm.pushes)
        {
            pr.my_mail = fm.d_mail;
            pr.stamp = fm.pushes;
        }
        KTimer kt(s, "forest_publish", 0.0);
        INC_CUDA(launch_forest_publish(a, 0u, s.stream, &pr));
    }
    fm.unpublished_step = 0;
    return INCERTO_OK;
}

void forest_graph_free(ForestModule& fm)
{
    if (fm.graph_exec) cudaGraphExecDestroy(fm.graph_exec);
    fm.graph_exec = nullptr;
    fm.graph_steps = 0;
}

// Launch-bound inner loop -> CUDA graph: the step kernel reads StepNumber from device memory and bumps
// it itself, so a captured run of steps replays without new arguments.
static constexpr uint32_t kGraphSteps = 16;

static int32_t forest_graph_ensure(Sim& s)
{
    ForestModule& fm = s.mods->forest;
    const void* sv = fm.series >= 0 ? s.series[fm.series].d_values : nullptr;
    const uint64_t sc = fm.series >= 0 ? s.series[fm.series].capacity : 0;
    if (fm.graph_exec && fm.graph_cur == fm.cur && fm.graph_series_values == sv && fm.graph_series_capacity == sc)
        return INCERTO_OK;
    forest_graph_free(fm);
    const int cur0 = fm.cur;
    // first with programmatic dependencies between the steps (INCERTO_B200_NO_PDL=1 or a runtime that cannot capture
    // them: plain kernel-after-kernel edges)
    static const bool no_pdl = [] { const char* e = std::getenv("INCERTO_B200_NO_PDL"); return e && e[0] == '1'; }();
    for (int attempt = no_pdl ? 1 : 0; attempt < 2 && !fm.graph_exec; ++attempt)
    {
        fm.graph_pdl = attempt == 0;
        cudaGraph_t graph = nullptr;
        const uint64_t launches0 = s.launches;
        INC_CUDA(cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal));
        int32_t st = INCERTO_OK;
        ForestArgs last;
        for (uint32_t i = 0; i < kGraphSteps && st == INCERTO_OK; ++i) st = forest_launch_one(s, 0, static_cast<int>(i), &last);
        if (st == INCERTO_OK)
        {
            // the graph's last node puts its last step on record and advances the device StepNumber by the steps held
            last.step_base = kGraphSteps - 1;
            if (launch_forest_publish(last, kGraphSteps, s.stream) != cudaSuccess) st = INCERTO_ERR_CUDA;
        }
        cudaError_t e = cudaStreamEndCapture(s.stream, &graph);
        fm.cur = cur0; // nothing ran yet
        if (st == INCERTO_OK && e == cudaSuccess) e = cudaGraphInstantiate(&fm.graph_exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (st != INCERTO_OK || e != cudaSuccess)
        {
            fm.graph_exec = nullptr;
            (void)cudaGetLastError();
            s.launches = launches0;
            if (attempt == 1 || !fm.graph_pdl)
            {
                if (st != INCERTO_OK) return st;
                INC_CUDA(e);
            }
            // the launches counted during the failed capture did not run
            kstat(s, "forest_step").launches -= kGraphSteps;
            kstat(s, "forest_step").algorithmic_bytes -= 2.0 * (fm.x_end - fm.x_begin) * fm.H * kGraphSteps;
        }
    }
    fm.graph_steps = kGraphSteps;
    fm.graph_cur = cur0;
    fm.graph_series_values = sv;
    fm.graph_series_capacity = sc;
    // the launches counted during capture did not run
    s.launches -= kGraphSteps;
    kstat(s, "forest_step").launches -= kGraphSteps;
    kstat(s, "forest_step").algorithmic_bytes -= 2.0 * (fm.x_end - fm.x_begin) * fm.H * kGraphSteps;
    return INCERTO_OK;
}

static bool forest_use_graph(const Sim& s, uint64_t nsteps) { return s.part_world == 1 && !s.profiling && nsteps >= 2 * kGraphSteps; }

static int32_t forest_sync_device_step(Sim& s, uint64_t first_step)
{
    // the captured steps read the device StepNumber: bring it to this run's first step (a graph leaves it right for
    // the next graph; single launches pass their step by value and do not touch it)
    ForestModule& fm = s.mods->forest;
    if (fm.d_step_mirror == first_step) return INCERTO_OK;
    launch_set_u32(s.mods->d_step, static_cast<uint32_t>(first_step), s.stream);
    INC_CUDA(cudaGetLastError());
    fm.d_step_mirror = first_step;
    return INCERTO_OK;
}

int32_t forest_prepare(Sim& s, uint64_t nsteps)
{
    if (forest_use_graph(s, nsteps))
    {
        INC_TRY(forest_graph_ensure(s));
        INC_TRY(forest_sync_device_step(s, s.step)); // before the run's first event (first use loads the kernel)
    }
    return INCERTO_OK;
}

int32_t forest_step(Sim& s, uint64_t first_step, uint32_t nsteps)
{
    ForestModule& fm = s.mods->forest;
    uint32_t done = 0;
    if (forest_use_graph(s, nsteps))
    {
        INC_TRY(forest_graph_ensure(s));
        const double bytes = 2.0 * (fm.x_end - fm.x_begin) * fm.H;
        INC_TRY(forest_sync_device_step(s, first_step));
        while (nsteps - done >= kGraphSteps)
        {
            INC_CUDA(cudaGraphLaunch(fm.graph_exec, s.stream));
            done += kGraphSteps;
            fm.d_step_mirror += kGraphSteps;
            fm.unpublished_step = 0; // the graph's last node published
            s.launches += kGraphSteps;
            KernelStat& ks = kstat(s, "forest_step");
            ks.launches += kGraphSteps;
            ks.algorithmic_bytes += bytes * kGraphSteps;
        }
    }
    for (; done < nsteps; ++done)
    {
        if (fm.peer && !fm.fused && s.comm && fm.pushes)
        {
            // the halo columns this step reads were stored by the neighbours after their previous step
            launch_forest_halo_wait(fm.d_mail, fm.peer_mail_lo ? fm.pushes : 0u, fm.peer_mail_hi ? fm.pushes : 0u, s.d_error, s.stream);
        }
        INC_TRY(forest_launch_one(s, first_step + done, -1));
        INC_TRY(forest_halo_exchange(s));
    }
    fm.stats_valid = true;
    return INCERTO_OK;
}

int32_t forest_read_cells(Sim& s, uint8_t* out_rows)
{
    ForestModule& fm = s.mods->forest;
    const uint32_t slab_w = static_cast<uint32_t>(fm.x_end - fm.x_begin);
    INC_CUDA(cudaMemcpy2DAsync(out_rows, fm.H, fm.d_slab[fm.cur] + fm.pitch, fm.pitch, fm.H, slab_w,
                               cudaMemcpyDeviceToHost, s.stream));
    if (fm.n_overlay)
        INC_CUDA(cudaMemcpyAsync(out_rows + static_cast<size_t>(slab_w) * fm.H, fm.d_ov[fm.cur], fm.n_overlay,
                                 cudaMemcpyDeviceToHost, s.stream));
    INC_CUDA(cudaStreamSynchronize(s.stream));
    return INCERTO_OK;
}

int32_t forest_stats_now(Sim& s, uint64_t out5[5])
{
    ForestModule& fm = s.mods->forest;
    INC_CUDA(cudaMemcpyAsync(out5, fm.d_stats, 5 * 8, cudaMemcpyDeviceToHost, s.stream));
    INC_CUDA(cudaStreamSynchronize(s.stream));
    if (s.part_world > 1 && s.comm)
    {
        INC_TRY(ensure_scratch(s, 64));
        INC_CUDA(cudaMemcpyAsync(s.d_scratch, out5, 40, cudaMemcpyHostToDevice, s.stream));
        INC_TRY(comm_allreduce(s, s.d_scratch, 5, 0, 0));
        INC_CUDA(cudaMemcpyAsync(out5, s.d_scratch, 40, cudaMemcpyDeviceToHost, s.stream));
        INC_CUDA(cudaStreamSynchronize(s.stream));
    }
    return INCERTO_OK;
}

} // namespace incerto

using namespace incerto;

extern "C"
{

int32_t incerto_partition_range(int32_t width, uint32_t rank, uint32_t world, int32_t* begin, int32_t* end)
{
    if (width < 0 || world == 0 || rank >= world || !begin || !end) return INCERTO_ERR_INVALID_ARGUMENT;
    *begin = static_cast<int32_t>(static_cast<int64_t>(width) * rank / world);
    *end = static_cast<int32_t>(static_cast<int64_t>(width) * (rank + 1) / world);
    return INCERTO_OK;
}

int32_t incerto_sim_halo_export(incerto_sim* sp, uint32_t side, void* out, size_t capacity, size_t* n_out)
{
    Sim* S = reinterpret_cast<Sim*>(sp);
    if (!S || side > 1 || !out) return INCERTO_ERR_INVALID_ARGUMENT;
    ForestModule& fm = S->mods->forest;
    INC_CUDA(cudaSetDevice(S->device));
    if (!fm.active && S->mods->air.part) return air_mail_export(*S, side, out, capacity, n_out);
    if (!fm.active) return INCERTO_ERR_UNSUPPORTED;
    const size_t n = static_cast<size_t>(fm.H) + fm.n_overlay;
    if (n_out) *n_out = n;
    if (capacity < n) return INCERTO_ERR_INVALID_ARGUMENT;
    const uint32_t slab_w = static_cast<uint32_t>(fm.x_end - fm.x_begin);
    const uint8_t* col = fm.d_slab[fm.cur] + static_cast<size_t>(side == 0 ? 1 : slab_w) * fm.pitch;
    INC_CUDA(cudaMemcpyAsync(out, col, fm.H, cudaMemcpyDeviceToHost, S->stream));
    if (fm.n_overlay)
        INC_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(out) + fm.H, fm.d_ov[fm.cur], fm.n_overlay, cudaMemcpyDeviceToHost, S->stream));
    INC_CUDA(cudaStreamSynchronize(S->stream));
    return INCERTO_OK;
}

int32_t incerto_sim_halo_import(incerto_sim* sp, uint32_t side, const void* in, size_t n)
{
    Sim* S = reinterpret_cast<Sim*>(sp);
    if (!S || side > 1 || !in) return INCERTO_ERR_INVALID_ARGUMENT;
    ForestModule& fm = S->mods->forest;
    INC_CUDA(cudaSetDevice(S->device));
    if (!fm.active && S->mods->air.part) return air_mail_import(*S, side, in, n);
    if (!fm.active) return INCERTO_ERR_UNSUPPORTED;
    if (n != static_cast<size_t>(fm.H) + fm.n_overlay) return INCERTO_ERR_INVALID_ARGUMENT;
    const uint32_t slab_w = static_cast<uint32_t>(fm.x_end - fm.x_begin);
    uint8_t* col = fm.d_slab[fm.cur] + static_cast<size_t>(side == 0 ? 0 : slab_w + 1) * fm.pitch;
    INC_CUDA(cudaMemcpyAsync(col, in, fm.H, cudaMemcpyHostToDevice, S->stream));
    if (fm.n_overlay)
        INC_CUDA(cudaMemcpyAsync(side == 0 ? fm.d_ov_lo : fm.d_ov_hi, static_cast<const uint8_t*>(in) + fm.H, fm.n_overlay,
                                 cudaMemcpyHostToDevice, S->stream));
    INC_CUDA(cudaStreamSynchronize(S->stream));
    return INCERTO_OK;
}

} // extern "C"
