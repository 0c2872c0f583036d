// air_traffic_3d: examples/air_traffic_3d.rs
//   spawn_aircraft  :61-90
//   move_aircraft   :93-121
//   check_conflicts :124-162  (SpatialGrid3D::neighbors_of, src/plugins/spatial_grid.rs:331-350)
// Pinned order: check_conflicts -> move_aircraft.  The grid the reference queries is the
// PreUpdate snapshot (positions at the start of the step), and with this order the positions the
// distance test reads are the same ones, so `length_squared <= 1` over the 26 neighbour cells (centre
// excluded) selects exactly the aircraft in the 6 face-neighbour cells.
//
// Device layout: GridPosition3D = int16 x, y, z + one pad lane (8 B), target_altitude int16 (2 B).
// Algorithmic bytes per aircraft-update: 18 (position 8 r + 8 w, target 2 r).
//
// The airspace is sparse (reference density 15 / 4000 cells), so the reference's HashMap<pos, HashSet>
// becomes, per step: (1) a hashed occupancy bitmap set by every aircraft - an exact "nobody there" answer
// for almost every neighbour cell -, (2) an exact hash table of per-cell counts holding only the aircraft that
// saw a set bit next to them (open addressing, 64-bit slots tag:8 | key:32 | count:24; the tag changes every
// step so that last step's entries read as empty and the table is cleared only when the tag wraps, every
// 255 steps), (3) the conflict count: exact lookups for the flagged neighbour cells only.  Two aircraft in
// face-adjacent cells always flag each other (the bitmap has no false negatives), so both are in the table.
//
// Draw contract:
//   spawn  aircraft uid: a = philox(key[ASPW], (uid,0,0,0)), b = philox(key[ASPW], (uid,0,0,1));
//          x = mulhi(a.x,SIZE), y = mulhi(a.y,SIZE), z = mulhi(a.z,HEIGHT), type = mulhi(a.w,3), target = mulhi(b.x,HEIGHT)
//   move   r = philox(key[AMOV], (uid,0,step,0)); horizontal = r.x < thr(0.7); dx = mulhi(r.y,3)-1; dy = mulhi(r.z,3)-1
#include "../../include/incerto_b200.h"
#include <algorithm>

#include "common.cuh"
#include "kernels_air.hpp"

namespace incerto
{
namespace
{

constexpr int kThreads = 256;

__device__ __forceinline__ int lane16(uint32_t w, int hi) { return static_cast<int16_t>(hi ? (w >> 16) : (w & 0xffffu)); }

__device__ __forceinline__ uint32_t slot_of(const AirArgs& a, uint32_t key)
{
    return (key * 0x9E3779B1u) >> a.slot_shift;
}

__device__ __forceinline__ void index_insert(const AirArgs& a, uint32_t key)
{
    uint32_t h = slot_of(a, key);
    const unsigned long long fresh = (static_cast<unsigned long long>(a.tag) << 56) | (static_cast<unsigned long long>(key) << 24) | 1ull;
    for (uint32_t guard = 0; guard <= a.slot_mask; )
    {
        const unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(&a.slots[h]);
        if (static_cast<uint32_t>(old >> 56) != a.tag)
        {
            if (atomicCAS(&a.slots[h], old, fresh) == old) return;
            continue; // somebody claimed the slot first: look at it again
        }
        if (static_cast<uint32_t>(old >> 24) == key)
        {
            atomicAdd(&a.slots[h], 1ull);
            return;
        }
        h = (h + 1) & a.slot_mask;
        ++guard;
    }
}

__device__ __forceinline__ uint32_t index_count(const AirArgs& a, uint32_t key)
{
    uint32_t h = slot_of(a, key);
    for (uint32_t guard = 0; guard <= a.slot_mask; ++guard)
    {
        const unsigned long long s = a.slots[h];
        if (static_cast<uint32_t>(s >> 56) != a.tag) return 0;
        if (static_cast<uint32_t>(s >> 24) == key) return static_cast<uint32_t>(s) & 0x00ffffffu;
        h = (h + 1) & a.slot_mask;
    }
    return 0;
}

__global__ void __launch_bounds__(kThreads) air_spawn_kernel(AirArgs a, int16_t* target_out, uint8_t* flags_out)
{
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n; i += gstride)
    {
        const uint4 r = philox4x32_10(static_cast<uint32_t>(i), 0u, 0u, 0u, a.rk_spawn);
        const uint4 q = philox4x32_10(static_cast<uint32_t>(i), 0u, 0u, 1u, a.rk_spawn);
        const uint32_t x = uniform_below(r.x, a.size), y = uniform_below(r.y, a.size), z = uniform_below(r.z, a.height);
        a.pos[i] = make_uint2(x | (y << 16), z);
        target_out[i] = static_cast<int16_t>(uniform_below(q.x, a.height));
        flags_out[i] = 1u;
    }
}

__device__ __forceinline__ bool cell_of(const AirArgs& a, const uint2& p, int& gx, int& gy, int& gz, uint32_t& key)
{
    gx = lane16(p.x, 0) - a.bmin[0];
    gy = lane16(p.x, 1) - a.bmin[1];
    gz = lane16(p.y, 0) - a.bmin[2];
    if (gx < 0 || gx >= a.ext[0] || gy < 0 || gy >= a.ext[1] || gz < 0 || gz >= a.ext[2]) return false;
    key = (static_cast<uint32_t>(gx) * a.ext[1] + static_cast<uint32_t>(gy)) * a.ext[2] + static_cast<uint32_t>(gz);
    return true;
}
// bit of a cell in the pre-filter: runs of 256 consecutive cell keys (the z column of ~25 adjacent y) share one
// hashed 32-byte sector, so the own cell and its +-z / +-y neighbours are usually answered by one sector
__device__ __forceinline__ uint32_t bit_of(const AirArgs& a, uint32_t key)
{
    return ((((key >> 8) * 0x85EBCA6Bu) >> a.bits_shift) << 8) | (key & 255u);
}

// the six face-neighbour cells inside the grid bounds (neighbors_of filters by bounds, spatial_grid.rs:337-340),
// in the fixed order -x +x -y +y -z +z
__device__ __forceinline__ uint32_t neighbour_keys(const AirArgs& a, int gx, int gy, int gz, uint32_t key, uint32_t (&nk)[6])
{
    const uint32_t sx = static_cast<uint32_t>(a.ext[1]) * a.ext[2], sy = a.ext[2];
    nk[0] = key - sx;
    nk[1] = key + sx;
    nk[2] = key - sy;
    nk[3] = key + sy;
    nk[4] = key - 1;
    nk[5] = key + 1;
    return (gx > 0 ? 1u : 0u) | (gx < a.ext[0] - 1 ? 2u : 0u) | (gy > 0 ? 4u : 0u) | (gy < a.ext[1] - 1 ? 8u : 0u) |
           (gz > 0 ? 16u : 0u) | (gz < a.ext[2] - 1 ? 32u : 0u);
}

// PreUpdate, pass 1: spatial_grid_update_system (spatial_grid.rs:434-443) - every aircraft sets the bit of its cell.
// Only launched when the previous step did not already leave the bitmap behind (first step, after a reset).
__global__ void __launch_bounds__(kThreads) air_bits_kernel(AirArgs a)
{
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n; i += gstride)
    {
        if (a.flags && !(a.flags[i] & 1u)) continue;
        int gx, gy, gz;
        uint32_t key;
        if (!cell_of(a, a.pos[i], gx, gy, gz, key))
        {
            // "entity at position {position:?} outside spatial grid bounds" (spatial_grid.rs:276-279)
            if (atomicCAS(a.error, 0u, static_cast<uint32_t>(INCERTO_ERR_OUT_OF_BOUNDS)) == 0u)
                a.error[1] = static_cast<uint32_t>(i);
            continue;
        }
        const uint32_t h = bit_of(a, key);
        atomicOr(&a.bits[h >> 5], 1u << (h & 31));
    }
}

// PreUpdate, pass 2: aircraft with a possibly occupied neighbour cell enter the exact per-cell index.
// One thread = 2 consecutive rows (one 128-bit position load, twelve independent bit probes in flight).
__global__ void __launch_bounds__(kThreads) air_candidates_kernel(AirArgs a)
{
    pdl_launch_dependents();
    pdl_wait(); // everything below reads what the previous kernel of the step wrote
    const uint64_t npairs = (a.n + 1) >> 1;
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    const uint64_t first = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (first == 0)
    {
        a.acc[0] = 0;
        a.acc[1] = 0;
        if (a.acc[2])
        {
            // an aircraft left the grid bounds in the last step: the reference panics in this PreUpdate (spatial_grid.rs:276-279)
            if (atomicCAS(a.error, 0u, static_cast<uint32_t>(INCERTO_ERR_OUT_OF_BOUNDS)) == 0u) a.error[1] = static_cast<uint32_t>(a.acc[2] - 1);
            a.acc[2] = 0;
        }
    }
    for (uint64_t q = first; q < npairs; q += gstride)
    {
        const uint64_t row0 = q * 2;
        const bool full = row0 + 2 <= a.n;
        uint2 p[2];
        if (full)
        {
            const uint4 v = *reinterpret_cast<const uint4*>(a.pos + row0);
            p[0] = make_uint2(v.x, v.y);
            p[1] = make_uint2(v.z, v.w);
        }
        else
            p[0] = p[1] = a.pos[row0];
        uint32_t key[2] = {0, 0}, probe[2] = {0, 0};
#pragma unroll
        for (int k = 0; k < 2; ++k)
        {
            int gx, gy, gz;
            if ((k == 1 && !full) || (a.flags && !(a.flags[row0 + k] & 1u)) || !cell_of(a, p[k], gx, gy, gz, key[k])) continue;
            uint32_t nk[6];
            const uint32_t valid = neighbour_keys(a, gx, gy, gz, key[k], nk);
#pragma unroll
            for (int j = 0; j < 6; ++j)
            {
                const uint32_t h = bit_of(a, nk[j]);
                if ((valid >> j) & 1u) probe[k] |= ((a.bits[h >> 5] >> (h & 31)) & 1u) << j;
            }
        }
        if (probe[0]) index_insert(a, key[0]);
        if (probe[1]) index_insert(a, key[1]);
        if (full)
            *reinterpret_cast<uint16_t*>(a.cand + row0) = static_cast<uint16_t>(probe[0] | (probe[1] << 8));
        else
            a.cand[row0] = static_cast<uint8_t>(probe[0]);
    }
}

// check_conflicts + move_aircraft; one thread = 2 consecutive rows (one 128-bit position load/store)
__global__ void __launch_bounds__(kThreads) air_step_kernel(AirArgs a)
{
    pdl_launch_dependents();
    pdl_wait(); // everything below reads what the previous kernel of the step wrote
    const uint64_t npairs = (a.n + 1) >> 1;
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    uint32_t conflicts = 0, indexed = 0;
    for (uint64_t q = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < npairs; q += gstride)
    {
        const uint64_t row0 = q * 2;
        const bool full = row0 + 2 <= a.n;
        uint2 p[2];
        uint32_t t2, c2 = 0;
        if (full)
        {
            const uint4 v = *reinterpret_cast<const uint4*>(a.pos + row0);
            p[0] = make_uint2(v.x, v.y);
            p[1] = make_uint2(v.z, v.w);
            t2 = *reinterpret_cast<const uint32_t*>(a.target + row0);
            if (a.do_check) c2 = *reinterpret_cast<const uint16_t*>(a.cand + row0);
        }
        else
        {
            p[0] = a.pos[row0];
            p[1] = p[0];
            t2 = static_cast<uint16_t>(a.target[row0]);
            if (a.do_check) c2 = a.cand[row0];
        }
#pragma unroll
        for (int k = 0; k < 2; ++k)
        {
            if (k == 1 && !full) break;
            const uint64_t row = row0 + k;
            if (a.flags && !(a.flags[row] & 1u)) continue;
            int x = lane16(p[k].x, 0), y = lane16(p[k].x, 1), z = lane16(p[k].y, 0);
            if (a.do_check)
            {
                int gx, gy, gz;
                uint32_t key;
                if (cell_of(a, p[k], gx, gy, gz, key)) // aircraft outside the bounds were reported by the bits kernel
                {
                    indexed += 1;
                    uint32_t mask = (c2 >> (8 * k)) & 0xffu;
                    if (mask)
                    {
                        uint32_t nk[6];
                        neighbour_keys(a, gx, gy, gz, key, nk);
                        while (mask)
                        {
                            const int j = __ffs(mask) - 1;
                            mask &= mask - 1;
                            conflicts += index_count(a, j == 0 ? nk[0] : (j == 1 ? nk[1] : (j == 2 ? nk[2] : (j == 3 ? nk[3] : (j == 4 ? nk[4] : nk[5])))));
                        }
                    }
                }
            }
            if (a.do_move)
            {
                const int target = lane16(t2, k);
                // :98-105 one altitude level towards the target
                z += (z < target) ? 1 : ((z > target) ? -1 : 0);
                const uint32_t uid = a.uid ? a.uid[row] : static_cast<uint32_t>(row);
                const uint4 r = philox4x32_10(uid, 0u, a.step, 0u, a.rk_move);
                if (bernoulli(r.x, a.thr_move)) // :108
                {
                    const int dx = static_cast<int>(uniform_below(r.y, 3u)) - 1, dy = static_cast<int>(uniform_below(r.z, 3u)) - 1;
                    x = min(max(x + dx, 0), a.size - 1); // :113-114 clamp(0, AIRSPACE_SIZE - 1)
                    y = min(max(y + dy, 0), a.size - 1);
                }
                p[k] = make_uint2((static_cast<uint32_t>(x) & 0xffffu) | (static_cast<uint32_t>(y) << 16),
                                  (static_cast<uint32_t>(z) & 0xffffu) | (p[k].y & 0xffff0000u));
            }
            if (a.part)
            {
                // domain decomposition: the aircraft left the slab -> the neighbour takes it over (and it stays around as
                // this rank's ghost: its bit below, its table entry from the sent list next step); it stands in a
                // boundary column -> the neighbour needs it as a ghost
                const int ngx = lane16(p[k].x, 0) - a.bmin[0];
                AirMailbox* mig = ngx < a.x_lo ? a.out_lo : (ngx >= a.x_hi ? a.out_hi : nullptr);
                if (ngx < a.x_lo || ngx >= a.x_hi)
                {
                    if (mig)
                    {
                        const uint32_t i = atomicAdd(&mig->n_migrants, 1u);
                        if (i < mig->capacity)
                            mig->migrants()[i] = make_uint4(a.uid ? a.uid[row] : static_cast<uint32_t>(row), p[k].x, p[k].y, 0u);
                        else
                            mig->overflow = 1u;
                    }
                    a.flags_rw[row] = 0; // no longer owned here
                }
                else
                {
                    if (ngx == a.x_lo && a.out_lo)
                    {
                        const uint32_t i = atomicAdd(&a.out_lo->n_ghosts, 1u);
                        if (i < a.out_lo->capacity) a.out_lo->ghosts()[i] = p[k]; else a.out_lo->overflow = 1u;
                    }
                    if (ngx == a.x_hi - 1 && a.out_hi)
                    {
                        const uint32_t i = atomicAdd(&a.out_hi->n_ghosts, 1u);
                        if (i < a.out_hi->capacity) a.out_hi->ghosts()[i] = p[k]; else a.out_hi->overflow = 1u;
                    }
                }
            }
            if (a.bits_next)
            {
                // the next PreUpdate's first pass, done here while the new position is in registers
                int gx, gy, gz;
                uint32_t key;
                if (cell_of(a, p[k], gx, gy, gz, key))
                {
                    const uint32_t h = bit_of(a, key);
                    atomicOr(&a.bits_next[h >> 5], 1u << (h & 31));
                }
                else
                    atomicMax(reinterpret_cast<unsigned long long*>(&a.acc[2]), static_cast<unsigned long long>(row + 1));
            }
        }
        if (a.do_move)
        {
            if (full)
                *reinterpret_cast<uint4*>(a.pos + row0) = make_uint4(p[0].x, p[0].y, p[1].x, p[1].y);
            else
                a.pos[row0] = p[0];
        }
    }
    if (a.do_check)
    {
        uint32_t v[2] = {conflicts, indexed};
        __shared__ uint32_t smem[64];
        block_sum<uint32_t, 2>(v, smem);
        if (threadIdx.x == 0)
        {
            if (v[0]) atomicAdd(reinterpret_cast<unsigned long long*>(&a.acc[0]), static_cast<unsigned long long>(v[0]));
            if (v[1]) atomicAdd(reinterpret_cast<unsigned long long*>(&a.acc[1]), static_cast<unsigned long long>(v[1]));
        }
    }
}

// ---------------------------------------------------------------------------------------------- domain decomposition
// After the spawn (every rank spawned the whole airspace, same uids): "alive" becomes "owned by this rank", and the
// aircraft standing just outside the slab are the first step's ghosts (written into the inboxes as if received).
__global__ void __launch_bounds__(kThreads) air_own_init_kernel(AirArgs a, AirMailbox* in_lo, AirMailbox* in_hi)
{
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n; i += gstride)
    {
        const uint2 p = a.pos[i];
        const int gx = lane16(p.x, 0) - a.bmin[0];
        const bool mine = gx >= a.x_lo && gx < a.x_hi;
        a.flags_rw[i] = mine ? 1u : 0u;
        AirMailbox* g = (gx == a.x_lo - 1) ? in_lo : (gx == a.x_hi ? in_hi : nullptr);
        if (g)
        {
            const uint32_t k = atomicAdd(&g->n_ghosts, 1u);
            if (k < g->capacity) g->ghosts()[k] = p; else g->overflow = 1u;
        }
    }
}

// Start of a step: the ghosts - what the neighbours sent after the previous step plus the aircraft this rank handed over
// itself (they stand in the neighbours' boundary columns now) - set their bits and enter the exact index.
__global__ void __launch_bounds__(kThreads) air_ghosts_kernel(AirArgs a, const AirMailbox* sent_lo, const AirMailbox* sent_hi)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    const AirMailbox* boxes[4] = {a.in_lo, a.in_hi, sent_lo, sent_hi};
#pragma unroll
    for (int b = 0; b < 4; ++b)
    {
        const AirMailbox* m = boxes[b];
        if (!m) continue;
        const bool sent = b >= 2;
        const uint32_t n = min(sent ? m->n_migrants : m->n_ghosts, m->capacity);
        if (m->overflow && t == 0) atomicCAS(a.error, 0u, static_cast<uint32_t>(INCERTO_ERR_UNSUPPORTED));
        for (uint32_t i = t; i < n; i += nt)
        {
            uint2 p;
            if (sent)
            {
                const uint4 r = m->migrants()[i];
                p = make_uint2(r.y, r.z);
            }
            else
                p = m->ghosts()[i];
            int gx, gy, gz;
            uint32_t key;
            if (!cell_of(a, p, gx, gy, gz, key)) continue;
            const uint32_t h = bit_of(a, key);
            atomicOr(&a.bits[h >> 5], 1u << (h & 31));
            index_insert(a, key);
        }
    }
}

// After the exchange: the aircraft that crossed into this slab come alive here (row found through the uid), with the
// position they moved to; their bit goes into the bitmap the last step kernel left for the next step.
__global__ void __launch_bounds__(kThreads) air_unpack_kernel(AirArgs a, AirMailbox* in_lo, AirMailbox* in_hi)
{
    AirMailbox* boxes[2] = {in_lo, in_hi};
    for (int b = 0; b < 2; ++b)
    {
        AirMailbox* m = boxes[b];
        if (!m) continue;
        const uint32_t n = min(m->n_migrants, m->capacity);
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
        {
            const uint4 r = m->migrants()[i];
            const uint64_t row = a.row_of_uid ? a.row_of_uid[r.x] : r.x;
            const uint2 p = make_uint2(r.y, r.z);
            a.pos[row] = p;
            a.flags_rw[row] = 1u;
            int gx, gy, gz;
            uint32_t key;
            if (cell_of(a, p, gx, gy, gz, key))
            {
                const uint32_t h = bit_of(a, key);
                atomicOr(&a.bits[h >> 5], 1u << (h & 31));
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) m->n_migrants = 0; // consumed (the ghost list is read by the next step's ghost kernel)
        __syncthreads();
    }
}

__global__ void air_mail_reset_kernel(AirMailbox* a, AirMailbox* b, AirMailbox* c, AirMailbox* d)
{
    AirMailbox* boxes[4] = {a, b, c, d};
    if (threadIdx.x < 4 && boxes[threadIdx.x])
    {
        boxes[threadIdx.x]->n_migrants = 0;
        boxes[threadIdx.x]->n_ghosts = 0;
    }
}

__global__ void invert_permutation_kernel(const uint32_t* __restrict__ perm, uint32_t* __restrict__ inv, uint64_t n)
{
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += gstride) inv[perm[i]] = static_cast<uint32_t>(i);
}

inline int blocks_for(uint64_t items)
{
    uint64_t b = (items + kThreads - 1) / kThreads;
    const uint64_t cap = static_cast<uint64_t>(kNumSMs) * 8 * 2;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<int>(b);
}

} // namespace

void launch_air_spawn(const AirArgs& a, int16_t* target_out, uint8_t* flags_out, cudaStream_t st)
{
    air_spawn_kernel<<<blocks_for(a.n), kThreads, 0, st>>>(a, target_out, flags_out);
}
void launch_air_bits(const AirArgs& a, cudaStream_t st) { air_bits_kernel<<<blocks_for(a.n), kThreads, 0, st>>>(a); }
void launch_air_candidates(const AirArgs& a, cudaStream_t st)
{
    static const int wave = wave_blocks(air_candidates_kernel, kThreads);
    launch_pdl(air_candidates_kernel, dim3(std::min(blocks_for(a.n), wave)), dim3(kThreads), 0, st, a);
}
void launch_air_step(const AirArgs& a, cudaStream_t st)
{
    static const int wave = wave_blocks(air_step_kernel, kThreads);
    launch_pdl(air_step_kernel, dim3(std::min(blocks_for((a.n + 1) / 2), wave)), dim3(kThreads), 0, st, a);
}
void launch_air_own_init(const AirArgs& a, AirMailbox* in_lo, AirMailbox* in_hi, cudaStream_t st)
{
    air_own_init_kernel<<<blocks_for(a.n), kThreads, 0, st>>>(a, in_lo, in_hi);
}
void launch_air_ghosts(const AirArgs& a, const AirMailbox* sent_lo, const AirMailbox* sent_hi, cudaStream_t st)
{
    air_ghosts_kernel<<<8, kThreads, 0, st>>>(a, sent_lo, sent_hi);
}
void launch_air_unpack(const AirArgs& a, AirMailbox* in_lo, AirMailbox* in_hi, cudaStream_t st)
{
    air_unpack_kernel<<<1, kThreads, 0, st>>>(a, in_lo, in_hi);
}
void launch_air_mail_reset(AirMailbox* a, AirMailbox* b, AirMailbox* c, AirMailbox* d, cudaStream_t st)
{
    air_mail_reset_kernel<<<1, 32, 0, st>>>(a, b, c, d);
}
void launch_invert_permutation(const uint32_t* perm, uint32_t* inv, uint64_t n, cudaStream_t st)
{
    invert_permutation_kernel<<<blocks_for(n), kThreads, 0, st>>>(perm, inv, n);
}

} // namespace incerto
