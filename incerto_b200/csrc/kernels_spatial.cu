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
// ContactHistory 20 x u32 slot-major):
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
#include "common.cuh"
#include "kernels_spatial.hpp"

namespace incerto
{
namespace
{

constexpr int kThreads = 256;
constexpr uint32_t kRowMask = (1u << kSpatialRowBits) - 1;
constexpr uint32_t kMany = 0xffffffffu;
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

__device__ __forceinline__ void mark_cell(const SpatialArgs& a, uint32_t key, uint32_t id)
{
    const uint32_t cell = (key & 0xffffu) * static_cast<uint32_t>(a.G) + (key >> 16);
    const uint32_t old = atomicCAS(&a.mark[cell], 0u, id);
    if (old != 0u && old != id && old != kMany) a.mark[cell] = kMany;
}

// Rare, long work (walking a contact history, probing 13 cells) is done by the whole warp for one person at
// a time - one load per lane - instead of by one lane with 31 idle: the warp's persons needing it are taken
// from a ballot.
__global__ void __launch_bounds__(kThreads) spatial_prepare_kernel(SpatialArgs a)
{
    const uint32_t G = static_cast<uint32_t>(a.G);
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp0 = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
    for (uint64_t base = warp0 * 32; base < a.n; base += nwarps * 32)
    {
        const uint64_t i = base + lane;
        const bool valid = i < a.n;
        uint32_t fl = 0, ax = 0, p = 0, q = 0, d = 0;
        if (valid)
        {
            fl = a.flags[i];
            ax = a.aux[i];
            p = a.pos[i];
            q = a.quar[i];
            d = a.disease[i];
        }
        const uint32_t op = ax >> 5;
        uint32_t len = ax & 31u;
        const uint32_t x = p & 0xffffu, y = p >> 16;
        const uint32_t cell = x * G + y;
        bool alive = valid && (fl & 1u);
        if (valid && !alive && op == OP_DIED)
        {
            // spatial_grid_cleanup_system (spatial_grid.rs:446-455)
            atomicSub(&a.count[cell], 1u);
            a.aux[i] = 0;
        }
        // spatial_grid_update_system (spatial_grid.rs:434-443) for Changed<GridPosition>
        if (alive && op == OP_NEW)
        {
            if (x >= G || y >= G)
            {
                if (atomicCAS(a.error, 0u, static_cast<uint32_t>(INCERTO_ERR_OUT_OF_BOUNDS)) == 0u) a.error[1] = static_cast<uint32_t>(i);
                alive = false; // left untouched
            }
            else
                atomicAdd(&a.count[cell], 1u);
        }
        const bool moved = alive && op >= OP_UP && op <= OP_DOWN;
        if (moved)
        {
            const uint32_t old = op == OP_UP ? cell + 1 : (op == OP_DOWN ? cell - 1 : (op == OP_LEFT ? cell + G : cell - G));
            atomicSub(&a.count[old], 1u);
            atomicAdd(&a.count[cell], 1u);
        }
        const bool free_ = q == 0;
        const uint32_t d0 = d;
        const uint32_t uid = (alive && a.uid) ? a.uid[i] : static_cast<uint32_t>(i);
        bool dying = false;
        if (alive)
        {
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
        }
        // :544-562 update_contact_history: a person who did not move already remembers this cell (it was inserted
        // last step) unless that insert emptied the set; :565-608 scatter side of the tracing: every cell a
        // non-quarantined person remembers is stamped with its id
        const bool need_hist = alive && (a.systems & 8u) && (len == 0 || moved);
        const bool need_mark = alive && (a.systems & 16u) && free_;
        uint32_t work = __ballot_sync(kFullMask, need_hist || need_mark);
        while (work)
        {
            const int src = __ffs(work) - 1;
            work &= work - 1;
            const uint64_t bi = base + static_cast<uint32_t>(src);
            uint32_t blen = __shfl_sync(kFullMask, len, src);
            const uint32_t bp = __shfl_sync(kFullMask, p, src);
            const uint32_t bid = __shfl_sync(kFullMask, uid, src) + 1u;
            const bool bhist = __shfl_sync(kFullMask, static_cast<int>(need_hist), src) != 0;
            const bool bmark = __shfl_sync(kFullMask, static_cast<int>(need_mark), src) != 0;
            uint32_t h = lane < blen ? a.hist[static_cast<uint64_t>(lane) * a.n_cap + bi] : 0u;
            if (bhist)
            {
                const bool found = __any_sync(kFullMask, lane < blen && h == bp);
                if (!found)
                {
                    if (blen >= a.history_cap)
                        blen = 0; // the insert made the set larger than 20: clear() (:555-558)
                    else
                    {
                        if (lane == blen)
                        {
                            a.hist[static_cast<uint64_t>(lane) * a.n_cap + bi] = bp;
                            h = bp;
                        }
                        blen += 1;
                    }
                }
            }
            if (bmark && lane < blen) mark_cell(a, h, bid);
            if (static_cast<int>(lane) == src) len = blen;
        }
        if (alive)
        {
            if (d != d0) a.disease[i] = static_cast<uint8_t>(d);
            const uint32_t nax = len | ((dying ? OP_DYING : OP_NONE) << 5);
            if (nax != ax) a.aux[i] = static_cast<uint8_t>(nax);
        }
    }
}

// (dx, dy) of the 13 cells within Manhattan distance 2, one per lane
__device__ __forceinline__ void offset13(uint32_t lane, int& dx, int& dy)
{
    // lane: 0 -> (0,0); 1..4 -> distance 1; 5..12 -> distance 2
    const int t[13][2] = {{0, 0}, {-1, 0}, {1, 0}, {0, -1}, {0, 1}, {-2, 0}, {2, 0}, {0, -2}, {0, 2}, {-1, -1}, {-1, 1}, {1, -1}, {1, 1}};
    dx = 0;
    dy = 0;
#pragma unroll
    for (int k = 0; k < 13; ++k)
        if (lane == static_cast<uint32_t>(k))
        {
            dx = t[k][0];
            dy = t[k][1];
        }
}

__global__ void __launch_bounds__(kThreads) spatial_interact_kernel(SpatialArgs a)
{
    const int G = a.G;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp0 = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
    int odx, ody;
    offset13(lane, odx, ody);
    const uint32_t odist = static_cast<uint32_t>((odx < 0 ? -odx : odx) + (ody < 0 ? -ody : ody));
    const bool olane = lane < 13u && odist <= a.radius; // beyond distance 2 the chance is 0.0 (:498)
    const uint64_t othr = odist <= 1u ? a.thr_infect1 : a.thr_infect2;
    for (uint64_t base = warp0 * 32; base < a.n; base += nwarps * 32)
    {
        const uint64_t i = base + lane;
        uint32_t fl = 0, ax = 0, q = 0, d = 0, p = 0;
        if (i < a.n)
        {
            fl = a.flags[i];
            ax = a.aux[i];
            q = a.quar[i];
            d = a.disease[i];
            p = a.pos[i];
        }
        const bool alive = (fl & 1u) != 0;
        const bool dying = (ax >> 5) == OP_DYING;
        const uint32_t q0 = q, d0 = d;
        const bool free_ = q == 0;
        const int x = static_cast<int>(p & 0xffffu), y = static_cast<int>(p >> 16);
        const uint32_t cell = static_cast<uint32_t>(x) * G + static_cast<uint32_t>(y);
        const uint32_t uid = (alive && a.uid) ? a.uid[i] : static_cast<uint32_t>(i);

        // :461-508 gather side of the transmission: the warp probes the 13 cells around one healthy,
        // non-quarantined person at a time
        uint32_t work = __ballot_sync(kFullMask, alive && (a.systems & 2u) && free_ && d == INCERTO_DISEASE_HEALTHY);
        while (work)
        {
            const int src = __ffs(work) - 1;
            work &= work - 1;
            const int bx = __shfl_sync(kFullMask, x, src), by = __shfl_sync(kFullMask, y, src);
            const uint32_t buid = __shfl_sync(kFullMask, uid, src);
            const int cx = bx + odx, cy = by + ody;
            bool hit = false;
            uint32_t c = 0;
            if (olane && cx >= 0 && cx < G && cy >= 0 && cy < G)
            {
                c = static_cast<uint32_t>(cx) * G + static_cast<uint32_t>(cy);
                hit = (a.src_bits[c >> 5] >> (c & 31)) & 1u;
            }
            if (!__any_sync(kFullMask, hit)) continue;
            bool exposed = false;
            if (hit)
            {
                const uint32_t h = a.src_head[c];
                uint32_t cur = ((h >> kSpatialRowBits) == a.tag) ? (h & kRowMask) : kRowMask;
                for (uint32_t guard = 0; cur != kRowMask && guard <= kRowMask; ++guard)
                {
                    const uint32_t suid = a.uid ? a.uid[cur] : cur;
                    exposed = exposed || bernoulli(philox4x32_10(buid, suid, a.step, 0u, a.rk_transmit).x, othr);
                    cur = a.src_next[cur];
                }
            }
            exposed = __any_sync(kFullMask, exposed);
            if (exposed && static_cast<int>(lane) == src) d = a.incubation; // :510-517
        }
        if (!alive) continue;

        // :580-606 gather side of the tracing
        bool quarantine = false;
        if ((a.systems & 16u) && free_)
        {
            const uint32_t m = a.mark[cell];
            quarantine = m != 0u && m != uid + 1u;
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
                for (int k = 0; k < 4; ++k)
                {
                    const int nx = k == 1 ? x - 1 : (k == 2 ? x + 1 : x), ny = k == 0 ? y - 1 : (k == 3 ? y + 1 : y);
                    if (nx < 0 || nx >= G || ny < 0 || ny >= G) continue;
                    const int zd = (nx > a.zone_x ? nx - a.zone_x : a.zone_x - nx) + (ny > a.zone_y ? ny - a.zone_y : a.zone_y - ny);
                    if (zd <= a.zone_radius && social) continue;
                    ok |= 1u << k;
                    cnt[k] = a.count[static_cast<uint32_t>(nx) * G + static_cast<uint32_t>(ny)];
                }
                if (social)
                {
                    // keep the least crowded candidates (:377-390)
                    uint32_t mn = 0xffffffffu;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if ((ok >> k) & 1u) mn = min(mn, cnt[k]);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (cnt[k] != mn) ok &= ~(1u << k);
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

        // end of Update: despawn (:530), try_insert(Quarantined) (:600-603), update_quarantine_status (:611-625)
        if (dying)
        {
            fl &= ~1u;
            newop = OP_DIED;
            a.flags[i] = static_cast<uint8_t>(fl);
        }
        else if (quarantine)
            q = a.quarantine_duration;
        else if ((a.systems & 32u) && q > 0u)
            q = q > 1u ? q - 1u : 0u;
        if (q != q0) a.quar[i] = static_cast<uint8_t>(q);
        if (d != d0) a.disease[i] = static_cast<uint8_t>(d);
        if (newop >= OP_UP && newop <= OP_DOWN) a.pos[i] = p;
        if (newop != (ax >> 5)) a.aux[i] = static_cast<uint8_t>((ax & 31u) | (newop << 5));
    }
}

// PandemicStats: counts by disease state among the live persons
__global__ void __launch_bounds__(kThreads) spatial_stats_kernel(const uint8_t* __restrict__ disease, const uint8_t* __restrict__ flags,
                                                                 uint64_t n, uint64_t initial_population, uint64_t* out6,
                                                                 uint32_t* partial, uint32_t* ticket)
{
    uint32_t v[4] = {0, 0, 0, 0}; // healthy, exposed, infectious, recovered
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += gstride)
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
void launch_spatial_prepare(const SpatialArgs& a, cudaStream_t st) { spatial_prepare_kernel<<<blocks_for(a.n), kThreads, 0, st>>>(a); }
void launch_spatial_interact(const SpatialArgs& a, cudaStream_t st) { spatial_interact_kernel<<<blocks_for(a.n), kThreads, 0, st>>>(a); }
void launch_spatial_stats(const uint8_t* disease, const uint8_t* flags, uint64_t n, uint64_t initial_population,
                          uint64_t* out6, void* scratch, uint32_t* ticket, cudaStream_t st)
{
    // scratch: >= 148 * 8 * 4 u32
    spatial_stats_kernel<<<blocks_for(n), kThreads, 0, st>>>(disease, flags, n, initial_population, out6,
                                                             static_cast<uint32_t*>(scratch), ticket);
}

} // namespace incerto
