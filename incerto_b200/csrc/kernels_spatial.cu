// pandemic_spatial: examples/pandemic_spatial.rs
//   spawn_population :279-313                       people_move_with_social_distancing :316-411
//   disease_incubation_progression :414-433         spatial_disease_transmission :436-518
//   disease_recovery_and_death :521-541             update_contact_history :544-562
//   process_contact_tracing :565-608                update_quarantine_status :611-625
//   PandemicStats :107-138
// Pinned order: incubation -> transmission -> recovery/death -> contact history -> contact
// tracing -> quarantine status -> move; despawn / try_insert(Quarantined) / remove(Quarantined) are applied at
// the end of the step in that issue order (a despawn wins over a quarantine insert).  The grid every system
// sees is the PreUpdate snapshot = the positions at the start of the step (the move comes last).
//
// Two passes per step over the person table (SoA: pos u32, disease u8, flags u8, quarantine u8, aux u8,
// ContactHistory 20 x u32 per person, contiguous):
//   prepare : PreUpdate grid maintenance (per-cell occupancy counts patched with last step's moves and deaths,
//             the device form of spatial_grid.rs:434-455), incubation, recovery/death draws, contact-history
//             insert, and the two scatter sides of the pair interactions: infectious non-quarantined people
//             are linked into per-cell source lists, non-quarantined people stamp every remembered cell
//             into the mark table (id of the only marker, or "several");
//   interact: the gather sides - a healthy non-quarantined person walks the source lists of the 13 cells within
//             Manhattan distance 2 (pair-keyed draws), a non-quarantined person standing on a cell marked by
//             somebody else is quarantined -, then the move against the occupancy snapshot and the command flush.
// The reference's formulation of tracing (every person A visits every cell it remembers and quarantines the
// others found there) and this one select the same set: B is quarantined iff some other non-quarantined
// person remembers B's cell.
//
// Draw contract:
//   spawn    r = philox(key[SSPW], (uid,0,0,0)): x = mulhi(r.x,G), y = mulhi(r.y,G), social = r.z < thr, exposed = r.w < thr
//   fate     infectious person: r = philox(key[SFAT], (uid,0,step,0)): die = r.x < thr(die) else recover = r.y < thr(rec)
//   transmit pair (healthy T, infectious S): philox(key[STRN], (uid_T, uid_S, step, 0)).x < thr(distance)
//   move     r = philox(key[SMOV], (uid,0,step,0)): attempt = r.x < thr(0.5); choose = mulhi(r.y, #best_moves)
#include "../../include/incerto_b200.h"
#include <algorithm>

#include "common.cuh"
#include "kernels_spatial.hpp"

namespace incerto
{
namespace
{

constexpr int kThreads = 256;
constexpr uint32_t kRowMask = (1u << kSpatialRowBits) - 1;
// pending ops carried in aux bits 5-7 from one step's flush to the next PreUpdate
enum : uint32_t
{
    OP_NONE = 0,
    OP_UP = 1,    // moved to (x, y - 1)
    OP_LEFT = 2,  // moved to (x - 1, y)
    OP_RIGHT = 3, // moved to (x + 1, y)
    OP_DOWN = 4,  // moved to (x, y + 1)
    OP_DIED = 5,  // despawned at the last flush: still counted in its cell
    OP_NEW = 6,   // spawned: not counted yet
    OP_DYING = 7  // prepare -> interact of the same step: the death draw succeeded
};

__global__ void __launch_bounds__(kThreads) spatial_spawn_kernel(SpatialArgs a)
{
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n; i += gstride)
    {
        const uint4 r = philox4x32_10(static_cast<uint32_t>(i), 0u, 0u, 0u, a.rk_spawn);
        a.pos[i] = uniform_below(r.x, a.G) | (uniform_below(r.y, a.G) << 16);
        a.flags[i] = static_cast<uint8_t>(1u | (bernoulli(r.z, a.thr_social) ? a.bit_social : 0u));
        a.disease[i] = bernoulli(r.w, a.thr_start) ? static_cast<uint8_t>(a.incubation) : static_cast<uint8_t>(INCERTO_DISEASE_HEALTHY);
        a.quar[i] = 0;
    }
}

// mark word of a cell: number of non-quarantined people remembering it (high half) | sum of their ids (low half).
// One fire-and-forget 64-bit reduction per mark; a person's history is a set, so nobody marks a cell twice.
// The reader only needs "somebody other than me": count >= 2, or count == 1 and the sum is not my id (with a
// single marker the sum cannot have carried into the count).
__device__ __forceinline__ void mark_cell(const SpatialArgs& a, uint32_t key, uint32_t id)
{
    const uint32_t cell = (key & 0xffffu) * static_cast<uint32_t>(a.G) + (key >> 16);
    atomicAdd(&a.mark[cell], (1ull << 32) | id);
}
// the exact inverse: the mark table is kept incrementally (a person's marks are added when it becomes free,
// extended when it remembers a new cell, and taken back when it is quarantined, forgets everything or dies)
__device__ __forceinline__ void unmark_cell(const SpatialArgs& a, uint32_t key, uint32_t id)
{
    const uint32_t cell = (key & 0xffffu) * static_cast<uint32_t>(a.G) + (key >> 16);
    atomicAdd(&a.mark[cell], 0ull - ((1ull << 32) | id));
}

// ---- eight persons per lane: byte columns travel as one 64-bit word, positions as two 128-bit words
constexpr unsigned long long kOnes = 0x0101010101010101ull;

// 0x01 in every byte of x that is zero (exact per byte)
__device__ __forceinline__ unsigned long long zero_bytes(unsigned long long x)
{
    const unsigned long long m = 0x7f7f7f7f7f7f7f7full;
    const unsigned long long t = ~(((x & m) + m) | x | m); // 0x80 where the byte is zero
    return t >> 7;
}
// bit k of the result = bit 0 of byte k
__device__ __forceinline__ uint32_t collect_lsb(unsigned long long x)
{
    return static_cast<uint32_t>(((x & kOnes) * 0x0102040810204080ull) >> 56);
}
__device__ __forceinline__ uint32_t byte_at(unsigned long long v, uint32_t k) { return static_cast<uint32_t>(v >> (8 * k)) & 0xffu; }
struct Tile8
{
    unsigned long long fl, ax, q, d;
    uint32_t pos[8];
};

__device__ __forceinline__ void load_tile(const SpatialArgs& a, uint64_t i0, Tile8& t)
{
    if (i0 + 8 <= a.n)
    {
        t.fl = *reinterpret_cast<const unsigned long long*>(a.flags + i0);
        t.ax = *reinterpret_cast<const unsigned long long*>(a.aux + i0);
        t.q = *reinterpret_cast<const unsigned long long*>(a.quar + i0);
        t.d = *reinterpret_cast<const unsigned long long*>(a.disease + i0);
        const uint4 p0 = *reinterpret_cast<const uint4*>(a.pos + i0), p1 = *reinterpret_cast<const uint4*>(a.pos + i0 + 4);
        t.pos[0] = p0.x;
        t.pos[1] = p0.y;
        t.pos[2] = p0.z;
        t.pos[3] = p0.w;
        t.pos[4] = p1.x;
        t.pos[5] = p1.y;
        t.pos[6] = p1.z;
        t.pos[7] = p1.w;
        return;
    }
    t.fl = t.ax = t.q = t.d = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
    {
        t.pos[k] = 0;
        if (i0 + k < a.n)
        {
            t.fl |= static_cast<unsigned long long>(a.flags[i0 + k]) << (8 * k);
            t.ax |= static_cast<unsigned long long>(a.aux[i0 + k]) << (8 * k);
            t.q |= static_cast<unsigned long long>(a.quar[i0 + k]) << (8 * k);
            t.d |= static_cast<unsigned long long>(a.disease[i0 + k]) << (8 * k);
            t.pos[k] = a.pos[i0 + k];
        }
    }
}
__device__ __forceinline__ void store_bytes(uint8_t* col, uint64_t i0, uint64_t n, unsigned long long v)
{
    if (i0 + 8 <= n)
        *reinterpret_cast<unsigned long long*>(col + i0) = v;
    else
        for (int k = 0; k < 8 && i0 + k < n; ++k) col[i0 + k] = static_cast<uint8_t>(v >> (8 * k));
}

// warp-wide exclusive prefix sum of a small per-lane count; the total is returned through `total`
__device__ __forceinline__ uint32_t warp_exclusive_scan(uint32_t v, uint32_t& total)
{
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t y = __shfl_up_sync(kFullMask, x, o);
        if (lane >= static_cast<uint32_t>(o)) x += y;
    }
    total = __shfl_sync(kFullMask, x, 31);
    return x - v;
}

// The persons of a 256-row warp tile that need any work this step, compacted: row, its four state bytes
// (flags | aux << 8 | quarantine << 16 | disease << 24) and its position.
// Entries wait in the queue until a full warp's worth (32) has accumulated, across tiles: a tile holds ~13 persons with
// work in the steady state, and dealing out each tile's entries on its own ran the whole slow path with 13 of 32 lanes
// (ncu, round 2: 12.9 active threads per instruction).  The queue is a stack: the top 32 entries are taken whenever there
// are that many; what is left when the warp runs out of tiles is processed at the end.
constexpr uint32_t kQueueCap = 31 + 256;
struct WorkQueue
{
    uint32_t row[kThreads / 32][kQueueCap + 1];
    uint32_t bytes[kThreads / 32][kQueueCap + 1];
    uint32_t pos[kThreads / 32][kQueueCap + 1];
    uint8_t marked[kThreads / 32][kQueueCap + 1]; // prepare only: the person's marks are in the mark table
};

// appends this tile's entries above the `have` entries already waiting; returns the new count
__device__ __forceinline__ uint32_t enqueue(WorkQueue& wq, uint32_t warp, const Tile8& t, uint64_t i0, uint32_t todo,
                                            uint32_t have, uint32_t marked_bits = 0)
{
    uint32_t total;
    uint32_t off = have + warp_exclusive_scan(__popc(todo), total);
    if (total)
    {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if ((todo >> k) & 1u)
            {
                wq.row[warp][off] = static_cast<uint32_t>(i0 + k);
                wq.bytes[warp][off] = byte_at(t.fl, k) | (byte_at(t.ax, k) << 8) | (byte_at(t.q, k) << 16) | (byte_at(t.d, k) << 24);
                wq.pos[warp][off] = t.pos[k];
                wq.marked[warp][off] = static_cast<uint8_t>((marked_bits >> k) & 1u);
                ++off;
            }
    }
    __syncwarp();
    return have + total;
}

// Streaming pass with a rare slow path.  A lane owns eight consecutive persons (128 bytes in flight per lane);
// byte-parallel tests pick the persons that need any work this step - in the steady state most people are
// quarantined, healthy and standing still, and need none.  Those are compacted into a per-warp queue in shared
// memory and dealt out one per lane, so that the per-person code (draws, history, marks) runs with full warps;
// the longer work of a person (reading its contact history, probing 13 cells) is a fully unrolled set of
// independent predicated loads, so that its latencies overlap.  Results go back by byte stores.
__global__ void __launch_bounds__(kThreads) spatial_prepare_kernel(SpatialArgs a)
{
    pdl_launch_dependents();
    pdl_wait(); // everything below reads what the previous kernel of the step wrote
    __shared__ WorkQueue wq;
    const uint32_t G = static_cast<uint32_t>(a.G);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t warp0 = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.stats_acc)
        for (int k = 0; k < 4; ++k) a.stats_acc[k] = 0;
    // the slow path of one person (entry j of this warp's queue)
    auto work = [&](uint32_t j) {
        const uint64_t i = wq.row[warp][j];
        const uint32_t bytes = wq.bytes[warp][j], p = wq.pos[warp][j];
        const uint32_t fl = bytes & 0xffu, ax = (bytes >> 8) & 0xffu, q = (bytes >> 16) & 0xffu;
        uint32_t d = bytes >> 24;
        const uint32_t d0 = d;
        const uint32_t op = ax >> 5;
        uint32_t len = ax & 31u;
        const uint32_t x = p & 0xffffu, y = p >> 16;
        const uint32_t cell = x * G + y;
        bool marked = wq.marked[warp][j] != 0;
        const bool alive = (fl & 1u) != 0;
        const bool free_ = q == 0;
        const uint32_t uid = a.uid ? a.uid[i] : static_cast<uint32_t>(i);
        const bool moved = alive && op >= OP_UP && op <= OP_DOWN;
        // :544-562 update_contact_history: a person who did not move already remembers this cell (it was
        // inserted last step) unless that insert emptied the set
        const bool need_hist = alive && (a.systems & 8u) && (len == 0 || moved);
        // :565-608 scatter side of the tracing, kept incrementally: the mark table holds the remembered cells of
        // exactly the non-quarantined live people
        const bool take_back = marked && !(alive && free_);
        const bool put_all = alive && (a.systems & 16u) && free_ && !marked;
        uint32_t hv[kSpatialHistorySlots];
        if (need_hist || take_back || put_all)
        {
            // all remembered cells at once: a person's 20 slots are contiguous (80 bytes), fetched as up to five
            // independent 128-bit loads (round 1 stored the slots slot-major: 20 predicated 4-byte loads, a 32-byte
            // sector each, and a fifth of this kernel's instructions)
            const uint4* hp = reinterpret_cast<const uint4*>(a.hist + static_cast<uint64_t>(i) * kSpatialHistorySlots);
#pragma unroll
            for (uint32_t v4 = 0; v4 < kSpatialHistorySlots / 4; ++v4)
            {
                uint4 q = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
                if (4 * v4 < len) q = hp[v4];
                hv[4 * v4 + 0] = 4 * v4 + 0 < len ? q.x : 0xffffffffu;
                hv[4 * v4 + 1] = 4 * v4 + 1 < len ? q.y : 0xffffffffu;
                hv[4 * v4 + 2] = 4 * v4 + 2 < len ? q.z : 0xffffffffu;
                hv[4 * v4 + 3] = 4 * v4 + 3 < len ? q.w : 0xffffffffu;
            }
        }
        if (take_back)
        {
            // quarantined at the last flush, or despawned: its marks (the set as it was while it was free) go
#pragma unroll
            for (uint32_t s2 = 0; s2 < kSpatialHistorySlots; ++s2)
                if (s2 < len) unmark_cell(a, hv[s2], uid + 1u);
            marked = false;
        }
        if (!alive)
        {
            // spatial_grid_cleanup_system (spatial_grid.rs:446-455): despawned at the last flush
            atomicSub(&a.count[cell], 1u);
            a.aux[i] = 0;
            return;
        }
        // spatial_grid_update_system (spatial_grid.rs:434-443) for Changed<GridPosition>
        if (op == OP_NEW)
        {
            if (x >= G || y >= G)
            {
                if (atomicCAS(a.error, 0u, static_cast<uint32_t>(INCERTO_ERR_OUT_OF_BOUNDS)) == 0u) a.error[1] = static_cast<uint32_t>(i);
                return; // left untouched
            }
            atomicAdd(&a.count[cell], 1u);
        }
        if (moved)
        {
            const uint32_t old = op == OP_UP ? cell + 1 : (op == OP_DOWN ? cell - 1 : (op == OP_LEFT ? cell + G : cell - G));
            atomicSub(&a.count[old], 1u);
            atomicAdd(&a.count[cell], 1u);
        }
        bool dying = false;
        // :414-433
        if ((a.systems & 1u) && d >= 1u && d < 0x80u) d = d > 1u ? d - 1u : INCERTO_DISEASE_INFECTIOUS;
        // :446-459 the infectious, non-quarantined people are the sources of this step
        if ((a.systems & 2u) && d == INCERTO_DISEASE_INFECTIOUS && free_)
        {
            atomicOr(&a.src_bits[cell >> 5], 1u << (cell & 31));
            const uint32_t old = atomicExch(&a.src_head[cell], (a.tag << kSpatialRowBits) | static_cast<uint32_t>(i));
            a.src_next[i] = ((old >> kSpatialRowBits) == a.tag) ? (old & kRowMask) : kRowMask;
        }
        // :521-541
        if ((a.systems & 4u) && d == INCERTO_DISEASE_INFECTIOUS)
        {
            const uint4 r = philox4x32_10(uid, 0u, a.step, 0u, a.rk_fate);
            if (bernoulli(r.x, a.thr_die))
                dying = true;
            else if (bernoulli(r.y, a.thr_recover))
                d = INCERTO_DISEASE_RECOVERED;
        }
        if (need_hist)
        {
            bool found = false;
#pragma unroll
            for (uint32_t s2 = 0; s2 < kSpatialHistorySlots; ++s2) found = found || hv[s2] == p;
            if (!found)
            {
                if (len >= a.history_cap)
                {
                    // the insert made the set larger than 20: clear() (:555-558)
                    if (marked)
                    {
#pragma unroll
                        for (uint32_t s2 = 0; s2 < kSpatialHistorySlots; ++s2)
                            if (s2 < len) unmark_cell(a, hv[s2], uid + 1u);
                    }
                    len = 0;
                }
                else
                {
                    a.hist[static_cast<uint64_t>(i) * kSpatialHistorySlots + len] = p;
#pragma unroll
                    for (uint32_t s2 = 0; s2 < kSpatialHistorySlots; ++s2)
                        if (s2 == len) hv[s2] = p;
                    len += 1;
                    if (marked) mark_cell(a, p, uid + 1u);
                }
            }
        }
        if (put_all)
        {
#pragma unroll
            for (uint32_t s2 = 0; s2 < kSpatialHistorySlots; ++s2)
                if (s2 < len) mark_cell(a, hv[s2], uid + 1u);
            marked = true;
        }
        if (marked != (wq.marked[warp][j] != 0))
        {
            // flip this person's bit (rare: once per quarantine / release)
            uint32_t* word = reinterpret_cast<uint32_t*>(a.marked) + (i >> 5);
            if (marked)
                atomicOr(word, 1u << (i & 31));
            else
                atomicAnd(word, ~(1u << (i & 31)));
        }
        if (d != d0) a.disease[i] = static_cast<uint8_t>(d);
        const uint32_t nax = len | ((dying ? OP_DYING : OP_NONE) << 5);
        if (nax != ax) a.aux[i] = static_cast<uint8_t>(nax);
    };
    uint32_t waiting = 0; // entries in this warp's queue (warp-uniform)
    for (uint64_t base = warp0 * 256; base < a.n; base += nwarps * 256)
    {
        const uint64_t i0 = base + lane * 8;
        Tile8 t;
        t.fl = t.ax = t.q = t.d = 0;
        if (i0 < a.n) load_tile(a, i0, t);
        const unsigned long long opbits = t.ax & (0xe0ull * kOnes);
        const unsigned long long alive8 = t.fl & kOnes;
        const unsigned long long op_nz = zero_bytes(opbits) ^ kOnes;
        const unsigned long long died = zero_bytes(opbits ^ ((OP_DIED << 5) * kOnes));
        const unsigned long long free8 = zero_bytes(t.q);
        const unsigned long long d_active = (zero_bytes(t.d) | zero_bytes(t.d ^ (0x81ull * kOnes))) ^ kOnes; // exposed or infectious
        const unsigned long long len0 = zero_bytes(t.ax & (0x1full * kOnes));
        // one bit per person: its marks are in the mark table (a tile of 8 persons = one byte)
        const uint32_t mk = ((a.systems & 16u) && i0 < a.n) ? a.marked[i0 >> 3] : 0u;
        const unsigned long long marked8 = zero_bytes((mk * kOnes) & 0x8040201008040201ull) ^ kOnes;
        unsigned long long want = alive8 & (op_nz | d_active);
        if (a.systems & 16u) want |= alive8 & (free8 ^ marked8); // became free: add its marks; quarantined: take them back
        if (a.systems & 8u) want |= alive8 & len0;
        want |= (alive8 ^ kOnes) & died;
        waiting = enqueue(wq, warp, t, i0, i0 < a.n ? collect_lsb(want) : 0u, waiting, mk);
        while (waiting >= 32) // a full warp's worth: take the top 32
        {
            waiting -= 32;
            work(waiting + lane);
            __syncwarp();
        }
    }
    if (lane < waiting) work(lane); // what was left when the tiles ran out
}

// (dx, dy) of the 13 cells within Manhattan distance 2: index 0 -> distance 0, 1..4 -> distance 1, 5..12 -> distance 2
__device__ constexpr int kOff13[13][2] = {{0, 0}, {-1, 0}, {1, 0}, {0, -1}, {0, 1}, {-2, 0}, {2, 0}, {0, -2}, {0, 2}, {-1, -1}, {-1, 1}, {1, -1}, {1, 1}};

__global__ void __launch_bounds__(kThreads) spatial_interact_kernel(SpatialArgs a)
{
    pdl_launch_dependents();
    pdl_wait(); // everything below reads what the previous kernel of the step wrote
    __shared__ WorkQueue wq;
    const int G = a.G;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t warp0 = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
    uint32_t st[4] = {0, 0, 0, 0}; // PandemicStats of the state after this step: healthy, exposed, infectious, recovered
    // the slow path of one person (entry j of this warp's queue)
    auto work = [&](uint32_t j) {
        const uint64_t i = wq.row[warp][j];
        const uint32_t bytes = wq.bytes[warp][j];
        uint32_t p = wq.pos[warp][j];
        uint32_t fl = bytes & 0xffu;
        const uint32_t ax = (bytes >> 8) & 0xffu, q = (bytes >> 16) & 0xffu;
        uint32_t d = bytes >> 24;
        const uint32_t d0 = d;
        const bool dying = (ax >> 5) == OP_DYING;
        const bool free_ = q == 0;
        const int x = static_cast<int>(p & 0xffffu), y = static_cast<int>(p >> 16);
        const uint32_t cell = static_cast<uint32_t>(x) * G + static_cast<uint32_t>(y);
        const uint32_t uid = a.uid ? a.uid[i] : static_cast<uint32_t>(i);

        // :461-508 gather side of the transmission: probe the 13 cells within Manhattan distance 2 (independent,
        // predicated bit loads), then walk the source lists of the few cells that hold a source
        if ((a.systems & 2u) && free_ && d == INCERTO_DISEASE_HEALTHY)
        {
            uint32_t hits = 0;
#pragma unroll
            for (int c13 = 0; c13 < 13; ++c13)
            {
                const int dx = kOff13[c13][0], dy = kOff13[c13][1];
                const int cx = x + dx, cy = y + dy;
                const uint32_t dist = static_cast<uint32_t>((dx < 0 ? -dx : dx) + (dy < 0 ? -dy : dy));
                if (dist <= a.radius && cx >= 0 && cx < G && cy >= 0 && cy < G)
                {
                    const uint32_t c = static_cast<uint32_t>(cx) * G + static_cast<uint32_t>(cy);
                    hits |= ((a.src_bits[c >> 5] >> (c & 31)) & 1u) << c13;
                }
            }
            bool exposed = false;
            while (hits)
            {
                const int c13 = __ffs(hits) - 1;
                hits &= hits - 1;
                int dx = 0, dy = 0;
#pragma unroll
                for (int k = 0; k < 13; ++k)
                    if (c13 == k)
                    {
                        dx = kOff13[k][0];
                        dy = kOff13[k][1];
                    }
                const uint32_t c = static_cast<uint32_t>(x + dx) * G + static_cast<uint32_t>(y + dy);
                const uint64_t thr = c13 <= 4 ? a.thr_infect1 : a.thr_infect2; // distance 0 | 1, else 2 (:495-499)
                const uint32_t h = a.src_head[c];
                uint32_t cur = ((h >> kSpatialRowBits) == a.tag) ? (h & kRowMask) : kRowMask;
                for (uint32_t guard = 0; cur != kRowMask && guard <= kRowMask; ++guard)
                {
                    const uint32_t suid = a.uid ? a.uid[cur] : cur;
                    exposed = exposed || bernoulli(philox4x32_10(uid, suid, a.step, 0u, a.rk_transmit).x, thr);
                    cur = a.src_next[cur];
                }
            }
            if (exposed) d = a.incubation; // :510-517
        }

        // :580-606 gather side of the tracing
        bool quarantine = false;
        if ((a.systems & 16u) && free_)
        {
            const unsigned long long m = a.mark[cell];
            const uint32_t markers = static_cast<uint32_t>(m >> 32);
            quarantine = markers >= 2u || (markers == 1u && static_cast<uint32_t>(m) != uid + 1u);
        }

        // :316-411 (a person despawned at this flush is left where it is: its move is unobservable)
        uint32_t newop = OP_NONE;
        if ((a.systems & 64u) && free_ && !dying)
        {
            const uint4 r = philox4x32_10(uid, 0u, a.step, 0u, a.rk_move);
            if (bernoulli(r.x, a.thr_attempt)) // :332
            {
                const bool social = (fl & a.bit_social) != 0;
                // candidates in the order up, left, right, down (:338-343); bit k of `ok`: candidate k passes
                // the bounds (:348-355) and quarantine-zone (:360-371) filters
                uint32_t ok = 0, cnt[4] = {0, 0, 0, 0};
#pragma unroll
                for (int c = 0; c < 4; ++c)
                {
                    const int nx = c == 1 ? x - 1 : (c == 2 ? x + 1 : x), ny = c == 0 ? y - 1 : (c == 3 ? y + 1 : y);
                    if (nx < 0 || nx >= G || ny < 0 || ny >= G) continue;
                    const int zd = (nx > a.zone_x ? nx - a.zone_x : a.zone_x - nx) + (ny > a.zone_y ? ny - a.zone_y : a.zone_y - ny);
                    if (zd <= a.zone_radius && social) continue;
                    ok |= 1u << c;
                    cnt[c] = a.count[static_cast<uint32_t>(nx) * G + static_cast<uint32_t>(ny)];
                }
                if (social)
                {
                    // keep the least crowded candidates (:377-390)
                    uint32_t mn = 0xffffffffu;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if ((ok >> c) & 1u) mn = min(mn, cnt[c]);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (cnt[c] != mn) ok &= ~(1u << c);
                }
                if (ok)
                {
                    const uint32_t pick = uniform_below(r.y, static_cast<uint32_t>(__popc(ok))); // best_moves.choose (:401)
                    const uint32_t dir = __fns(ok, 0, static_cast<int>(pick) + 1);
                    const uint32_t people = dir == 0 ? cnt[0] : (dir == 1 ? cnt[1] : (dir == 2 ? cnt[2] : cnt[3]));
                    if (people < a.crowding) // :403-408
                    {
                        const int mx = dir == 1 ? x - 1 : (dir == 2 ? x + 1 : x), my = dir == 0 ? y - 1 : (dir == 3 ? y + 1 : y);
                        p = static_cast<uint32_t>(mx) | (static_cast<uint32_t>(my) << 16);
                        newop = dir + 1;
                    }
                }
            }
        }

        // end of Update: despawn (:530), try_insert(Quarantined) (:600-603); a quarantined dying person had its
        // counter decremented above, which nobody can observe
        if (dying)
        {
            fl &= ~1u;
            newop = OP_DIED;
            a.flags[i] = static_cast<uint8_t>(fl);
        }
        else if (quarantine)
            a.quar[i] = static_cast<uint8_t>(a.quarantine_duration);
        if (d != d0) a.disease[i] = static_cast<uint8_t>(d);
        if (newop >= OP_UP && newop <= OP_DOWN) a.pos[i] = p;
        if (newop != (ax >> 5)) a.aux[i] = static_cast<uint8_t>((ax & 31u) | (newop << 5));
        if (fl & 1u)
        {
            st[0] += d == INCERTO_DISEASE_HEALTHY;
            st[1] += d >= 1u && d < 0x80u;
            st[2] += d == INCERTO_DISEASE_INFECTIOUS;
            st[3] += d == INCERTO_DISEASE_RECOVERED;
        }
    };
    uint32_t waiting = 0; // entries in this warp's queue (warp-uniform)
    for (uint64_t base = warp0 * 256; base < a.n; base += nwarps * 256)
    {
        const uint64_t i0 = base + lane * 8;
        Tile8 t;
        t.fl = t.ax = t.q = t.d = 0;
        if (i0 < a.n) load_tile(a, i0, t);
        const unsigned long long alive8 = t.fl & kOnes;
        const unsigned long long free8 = zero_bytes(t.q);
        const unsigned long long dying8 = zero_bytes((t.ax & (0xe0ull * kOnes)) ^ ((OP_DYING << 5) * kOnes));
        const unsigned long long busy8 = alive8 & (free8 | dying8); // everybody else only counts its quarantine down
        // update_quarantine_status (:611-625) for all the quarantined at once: n > 1 -> n - 1, n == 1 -> removed at
        // the flush.  (People quarantined at this flush were not quarantined during the step: disjoint from these.)
        if ((a.systems & 32u) && i0 < a.n)
        {
            const unsigned long long nq = t.q - (alive8 & (free8 ^ kOnes));
            if (nq != t.q) store_bytes(a.quar, i0, a.n, nq);
        }
        {
            // fused PandemicStats (:107-138): the people whose state this kernel does not touch any further
            const unsigned long long al = alive8 & (busy8 ^ kOnes);
            const unsigned long long healthy = zero_bytes(t.d) & al, infectious = zero_bytes(t.d ^ (0x80ull * kOnes)) & al,
                                     recovered = zero_bytes(t.d ^ (0x81ull * kOnes)) & al;
            st[0] += __popcll(healthy);
            st[2] += __popcll(infectious);
            st[3] += __popcll(recovered);
            st[1] += __popcll(al) - __popcll(healthy) - __popcll(infectious) - __popcll(recovered);
        }
        // the packed quarantine store above is ordered before the byte stores below by the barrier inside enqueue()
        waiting = enqueue(wq, warp, t, i0, i0 < a.n ? collect_lsb(busy8) : 0u, waiting);
        while (waiting >= 32) // a full warp's worth: take the top 32
        {
            waiting -= 32;
            work(waiting + lane);
            __syncwarp();
        }
    }
    if (lane < waiting) work(lane); // what was left when the tiles ran out
    if (a.stats_acc)
    {
        __shared__ uint32_t smem[4 * 32];
        block_sum<uint32_t, 4>(st, smem);
        if (threadIdx.x == 0)
            for (int k = 0; k < 4; ++k)
                if (st[k]) atomicAdd(reinterpret_cast<unsigned long long*>(&a.stats_acc[k]), static_cast<unsigned long long>(st[k]));
    }
}

// out6 <- PandemicStats from the counts accumulated by the interact kernel
__global__ void spatial_stats_finalize_kernel(const uint64_t* acc, uint64_t initial_population, uint64_t* out6)
{
    const uint64_t alive = acc[0] + acc[1] + acc[2] + acc[3];
    out6[0] = acc[0];
    out6[1] = acc[1];
    out6[2] = acc[2];
    out6[3] = acc[3];
    out6[4] = initial_population - alive; // dead_count: INITIAL_POPULATION - components.len() (:134)
    out6[5] = alive;
}

// PandemicStats: counts by disease state among the live persons (standalone: 16 persons per thread and load)
__global__ void __launch_bounds__(kThreads) spatial_stats_kernel(const uint8_t* __restrict__ disease, const uint8_t* __restrict__ flags,
                                                                 uint64_t n, uint64_t initial_population, uint64_t* out6,
                                                                 uint32_t* partial, uint32_t* ticket)
{
    uint32_t v[4] = {0, 0, 0, 0}; // healthy, exposed, infectious, recovered
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    const uint64_t nvec = n / 16;
    for (uint64_t j = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j < nvec; j += gstride)
    {
        const uint4 f = ld_stream_v4(flags + 16 * j), d = ld_stream_v4(disease + 16 * j);
        const unsigned long long fw[2] = {f.x | (static_cast<unsigned long long>(f.y) << 32), f.z | (static_cast<unsigned long long>(f.w) << 32)};
        const unsigned long long dw[2] = {d.x | (static_cast<unsigned long long>(d.y) << 32), d.z | (static_cast<unsigned long long>(d.w) << 32)};
#pragma unroll
        for (int h = 0; h < 2; ++h)
        {
            const unsigned long long al = fw[h] & kOnes;
            const uint32_t he = __popcll(zero_bytes(dw[h]) & al), in = __popcll(zero_bytes(dw[h] ^ (0x80ull * kOnes)) & al),
                           re = __popcll(zero_bytes(dw[h] ^ (0x81ull * kOnes)) & al);
            v[0] += he;
            v[2] += in;
            v[3] += re;
            v[1] += __popcll(al) - he - in - re;
        }
    }
    for (uint64_t i = nvec * 16 + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += gstride)
    {
        if (!(flags[i] & 1u)) continue;
        const uint32_t d = disease[i];
        v[0] += d == INCERTO_DISEASE_HEALTHY;
        v[1] += d >= 1u && d < 0x80u;
        v[2] += d == INCERTO_DISEASE_INFECTIOUS;
        v[3] += d == INCERTO_DISEASE_RECOVERED;
    }
    __shared__ uint32_t smem[4 * 32];
    block_sum<uint32_t, 4>(v, smem);
    if (threadIdx.x == 0)
        for (int k = 0; k < 4; ++k) partial[blockIdx.x * 4 + k] = v[k];
    if (last_block_arrives(ticket, gridDim.x))
    {
        if (threadIdx.x == 0)
        {
            uint64_t t[4] = {0, 0, 0, 0};
            for (uint32_t b = 0; b < gridDim.x; ++b)
                for (int k = 0; k < 4; ++k) t[k] += partial[b * 4 + k];
            const uint64_t alive = t[0] + t[1] + t[2] + t[3];
            out6[0] = t[0];
            out6[1] = t[1];
            out6[2] = t[2];
            out6[3] = t[3];
            out6[4] = initial_population - alive; // dead_count: INITIAL_POPULATION - components.len() (:134)
            out6[5] = alive;
        }
    }
}

inline int blocks_for(uint64_t items)
{
    uint64_t b = (items + kThreads - 1) / kThreads;
    const uint64_t cap = static_cast<uint64_t>(kNumSMs) * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<int>(b);
}

} // namespace

void launch_spatial_spawn(const SpatialArgs& a, cudaStream_t st) { spatial_spawn_kernel<<<blocks_for(a.n), kThreads, 0, st>>>(a); }
void launch_spatial_prepare(const SpatialArgs& a, cudaStream_t st)
{
    static const int wave = wave_blocks(spatial_prepare_kernel, kThreads);
    launch_pdl(spatial_prepare_kernel, dim3(std::min(blocks_for((a.n + 7) / 8), wave)), dim3(kThreads), 0, st, a);
}
void launch_spatial_interact(const SpatialArgs& a, cudaStream_t st)
{
    static const int wave = wave_blocks(spatial_interact_kernel, kThreads);
    launch_pdl(spatial_interact_kernel, dim3(std::min(blocks_for((a.n + 7) / 8), wave)), dim3(kThreads), 0, st, a);
}
void launch_spatial_stats_finalize(const uint64_t* acc, uint64_t initial_population, uint64_t* out6, cudaStream_t st)
{
    spatial_stats_finalize_kernel<<<1, 1, 0, st>>>(acc, initial_population, out6);
}
void launch_spatial_stats(const uint8_t* disease, const uint8_t* flags, uint64_t n, uint64_t initial_population,
                          uint64_t* out6, void* scratch, uint32_t* ticket, cudaStream_t st)
{
    // scratch: >= 148 * 8 * 4 u32
    spatial_stats_kernel<<<blocks_for(n), kThreads, 0, st>>>(disease, flags, n, initial_population, out6,
                                                             static_cast<uint32_t*>(scratch), ticket);
}

} // namespace incerto
