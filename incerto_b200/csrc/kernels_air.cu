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
// Bitmap layout (round 2; ncu had the candidates kernel bound by its six scattered 4-byte probes per aircraft, one
// L1 wavefront each): cells are blocked 4 x 4 x 16 (x, y, z) into 256-bit sectors, 2 x 2 sectors into a 128-byte
// line (8 x 8 x 16 cells), and the LINE is what gets hashed.  An aircraft fetches its own sector with one 256-bit
// load and answers +-z and, unless it sits on a sector face, +-x / +-y from registers; only the crossing
// neighbours (on average one per aircraft) cost another probe, and that one mostly hits the line just fetched.
// The flagged aircraft are appended to a compact list (row | neighbour mask << 32, one warp-aggregated atomic per
// 64 rows); the insertions into the exact table and the exact look-ups run over that list with full warps.
// The exact table is allocated for the worst case (every aircraft flagged) but a step only uses a power-of-two
// prefix of it sized, on the device, by the number of aircraft the candidates kernel flagged (a few per cent at the
// reference density): the slots it touches stay in the L2 instead of being 134 MB of random DRAM accesses.
//
// Draw contract:
//   spawn  aircraft uid: a = philox(key[ASPW], (uid,0,0,0)), b = philox(key[ASPW], (uid,0,0,1));
//          x = mulhi(a.x,SIZE), y = mulhi(a.y,SIZE), z = mulhi(a.z,HEIGHT), type = mulhi(a.w,3), target = mulhi(b.x,HEIGHT)
//   move   r = philox(key[AMOV], (uid,0,step,0)); horizontal = r.x < thr(0.7); dx = mulhi(r.y,3)-1; dy = mulhi(r.z,3)-1
#include "../../include/incerto_b200.h"
#include "common.cuh"
#include "kernels_air.hpp"

namespace incerto
{
namespace
{

constexpr int kThreads = 256;

__device__ __forceinline__ int lane16(uint32_t w, int hi) { return static_cast<int16_t>(hi ? (w >> 16) : (w & 0xffffu)); }

// The part of the exact table this step uses: the smallest power of two >= 4 x the aircraft flagged by the
// candidates kernel (counted in acc[4 + (step & 1)]), at least 1024 slots, at most the whole allocation.
struct TableView
{
    uint32_t mask, shift;
};
__device__ __forceinline__ TableView table_view(const AirArgs& a)
{
    const unsigned long long nf = a.acc[4 + (a.step & 1u)];
    const unsigned long long want = nf * 4ull > 1024ull ? nf * 4ull : 1024ull;
    uint32_t bits = 64u - static_cast<uint32_t>(__clzll(static_cast<long long>(want - 1ull)));
    const uint32_t max_bits = 32u - a.slot_shift;
    bits = bits > max_bits ? max_bits : bits;
    TableView v;
    v.mask = (1u << bits) - 1u;
    v.shift = 32u - bits;
    return v;
}

__device__ __forceinline__ uint32_t slot_of(const TableView& tv, uint32_t key) { return (key * 0x9E3779B1u) >> tv.shift; }

__device__ __forceinline__ void index_insert(const AirArgs& a, const TableView& tv, uint32_t key)
{
    uint32_t h = slot_of(tv, key);
    const unsigned long long fresh = (static_cast<unsigned long long>(a.tag) << 56) | (static_cast<unsigned long long>(key) << 24) | 1ull;
    for (uint32_t guard = 0; guard <= tv.mask; )
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
        h = (h + 1) & tv.mask;
        ++guard;
    }
}

__device__ __forceinline__ uint32_t index_count(const AirArgs& a, const TableView& tv, uint32_t key)
{
    uint32_t h = slot_of(tv, key);
    for (uint32_t guard = 0; guard <= tv.mask; ++guard)
    {
        const unsigned long long s = a.slots[h];
        if (static_cast<uint32_t>(s >> 56) != a.tag) return 0;
        if (static_cast<uint32_t>(s >> 24) == key) return static_cast<uint32_t>(s) & 0x00ffffffu;
        h = (h + 1) & tv.mask;
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
// Bit of a cell in the pre-filter.  Sector = 4 x 4 x 16 cells (bit = (x & 3) * 64 + (y & 3) * 16 + (z & 15)), line =
// 2 x 2 sectors = 8 x 8 x 16 cells, the line index is hashed into the table of 2^(filter_bits - 10) lines.
__device__ __forceinline__ uint32_t bit_of(const AirArgs& a, int gx, int gy, int gz)
{
    const uint32_t line = ((static_cast<uint32_t>(gx) >> 3) * a.lines_y + (static_cast<uint32_t>(gy) >> 3)) * a.lines_z +
                          (static_cast<uint32_t>(gz) >> 4);
    const uint32_t slot = (line * 0x85EBCA6Bu) >> a.bits_shift;
    return (slot << 10) | ((static_cast<uint32_t>(gx) & 4u) << 7) | ((static_cast<uint32_t>(gy) & 4u) << 6) |
           ((static_cast<uint32_t>(gx) & 3u) << 6) | ((static_cast<uint32_t>(gy) & 3u) << 4) | (static_cast<uint32_t>(gz) & 15u);
}
__device__ __forceinline__ void ld_sector(const uint32_t* p, unsigned long long (&q)[4])
{
    // one 256-bit load: the four x-slices (64 bits = 4 y-columns of 16 z-bits) of a sector
    asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(q[0]), "=l"(q[1]), "=l"(q[2]), "=l"(q[3]) : "l"(p));
}

__device__ __forceinline__ uint32_t probe_bit(const AirArgs& a, int gx, int gy, int gz)
{
    const uint32_t h = bit_of(a, gx, gy, gz);
    return (a.bits[h >> 5] >> (h & 31u)) & 1u;
}

// validity of the six face-neighbour cells inside the grid bounds (neighbors_of filters by bounds,
// spatial_grid.rs:337-340), in the fixed order -x +x -y +y -z +z
__device__ __forceinline__ uint32_t neighbour_valid(const AirArgs& a, int gx, int gy, int gz)
{
    return (gx > 0 ? 1u : 0u) | (gx < a.ext[0] - 1 ? 2u : 0u) | (gy > 0 ? 4u : 0u) | (gy < a.ext[1] - 1 ? 8u : 0u) |
           (gz > 0 ? 16u : 0u) | (gz < a.ext[2] - 1 ? 32u : 0u);
}
__device__ __forceinline__ uint32_t neighbour_key(const AirArgs& a, uint32_t key, int j)
{
    const uint32_t sx = static_cast<uint32_t>(a.ext[1]) * a.ext[2], sy = a.ext[2];
    return j == 0 ? key - sx : (j == 1 ? key + sx : (j == 2 ? key - sy : (j == 3 ? key + sy : (j == 4 ? key - 1 : key + 1))));
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
        const uint32_t h = bit_of(a, gx, gy, gz);
        atomicOr(&a.bits[h >> 5], 1u << (h & 31u));
    }
}

// which of the six face-neighbour cells of (gx, gy, gz) show a set bit in the pre-filter
__device__ __forceinline__ uint32_t probe_neighbours(const AirArgs& a, int gx, int gy, int gz)
{
    const uint32_t valid = neighbour_valid(a, gx, gy, gz);
    const uint32_t line = ((static_cast<uint32_t>(gx) >> 3) * a.lines_y + (static_cast<uint32_t>(gy) >> 3)) * a.lines_z +
                          (static_cast<uint32_t>(gz) >> 4);
    // position inside the line: sector (bits 9, 8), x-slice (7, 6), y-column (5, 4), level (3..0)
    const uint32_t low = ((static_cast<uint32_t>(gx) & 4u) << 7) | ((static_cast<uint32_t>(gy) & 4u) << 6) |
                         ((static_cast<uint32_t>(gx) & 3u) << 6) | ((static_cast<uint32_t>(gy) & 3u) << 4) | (static_cast<uint32_t>(gz) & 15u);
    const uint32_t h = (((line * 0x85EBCA6Bu) >> a.bits_shift) << 10) | low;
    unsigned long long q[4];
    ld_sector(a.bits + ((h >> 8) << 3), q);
    const uint32_t xs = static_cast<uint32_t>(gx) & 3u, ys = static_cast<uint32_t>(gy) & 3u, zs = static_cast<uint32_t>(gz) & 15u;
    // the neighbours that leave the sector: at most one per axis; their probes are issued before anything is consumed.
    // A crossing neighbour has the mirrored coordinate inside its sector (x-slice 0 <-> 3: flip bits 7, 6; ...) in the
    // adjacent sector of the line (flip the sector bit) or, on a line face, in the adjacent line.
    const bool cx = xs == 0 ? (valid & 1u) != 0 : (xs == 3 ? (valid & 2u) != 0 : false);
    const bool cy = ys == 0 ? (valid & 4u) != 0 : (ys == 3 ? (valid & 8u) != 0 : false);
    const bool cz = zs == 0 ? (valid & 16u) != 0 : (zs == 15 ? (valid & 32u) != 0 : false);
    auto cross = [&](uint32_t line_n, uint32_t flip) -> uint32_t {
        const uint32_t hh = (((line_n * 0x85EBCA6Bu) >> a.bits_shift) << 10) | (low ^ flip);
        return (a.bits[hh >> 5] >> (hh & 31u)) & 1u;
    };
    uint32_t px = 0, py = 0, pz = 0;
    if (cx)
    {
        const uint32_t g7 = static_cast<uint32_t>(gx) & 7u, step_l = a.lines_y * a.lines_z;
        px = cross(xs == 0 ? (g7 == 0 ? line - step_l : line) : (g7 == 7 ? line + step_l : line), 0x2c0u);
    }
    if (cy)
    {
        const uint32_t g7 = static_cast<uint32_t>(gy) & 7u;
        py = cross(ys == 0 ? (g7 == 0 ? line - a.lines_z : line) : (g7 == 7 ? line + a.lines_z : line), 0x130u);
    }
    if (cz) pz = cross(zs == 0 ? line - 1u : line + 1u, 0x00fu);
    // inside the sector: the own x-slice (64 bits = 4 y-columns of 16 levels) shifted so that the own bit lands on
    // bit 0 (`up`: +z is bit 1, +y bit 16) and on bit 63 (`dn`: -z is bit 62, -y bit 47); the x neighbours sit at the
    // own bit position of the adjacent slices
    const unsigned long long own = xs == 0 ? q[0] : (xs == 1 ? q[1] : (xs == 2 ? q[2] : q[3]));
    const unsigned long long lo = xs == 1 ? q[0] : (xs == 2 ? q[1] : q[2]); // slice x-1 (xs > 0)
    const unsigned long long hi = xs == 0 ? q[1] : (xs == 1 ? q[2] : q[3]); // slice x+1 (xs < 3)
    const uint32_t sh = 16u * ys + zs;
    const unsigned long long up = own >> sh, dn = own << (63u - sh);
    uint32_t probe = 0;
    probe |= (xs > 0 ? static_cast<uint32_t>(lo >> sh) & 1u : px) << 0;
    probe |= (xs < 3 ? static_cast<uint32_t>(hi >> sh) & 1u : px) << 1;
    probe |= (ys > 0 ? static_cast<uint32_t>(dn >> 47) & 1u : py) << 2;
    probe |= (ys < 3 ? static_cast<uint32_t>(up >> 16) & 1u : py) << 3;
    probe |= (zs > 0 ? static_cast<uint32_t>(dn >> 62) & 1u : pz) << 4;
    probe |= (zs < 15 ? static_cast<uint32_t>(up >> 1) & 1u : pz) << 5;
    return probe & valid;
}

// PreUpdate, pass 2: which aircraft have a possibly occupied neighbour cell?  One thread = 2 consecutive rows (one
// 128-bit position load, two sector loads + the crossing probes in flight).  Also counts them (the exact table of
// this step is sized from the count) and zeroes the next step's bitmap on the side.
__global__ void __launch_bounds__(kThreads) air_candidates_kernel(AirArgs a)
{
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
    if (a.bits_next)
    {
        uint4* z = reinterpret_cast<uint4*>(a.bits_next);
        for (uint64_t i = first; i < a.bits_vec4; i += gstride) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long* counter = reinterpret_cast<unsigned long long*>(&a.acc[4 + (a.step & 1u)]);
    // every warp runs the same number of iterations (the shuffles below need all lanes)
    const uint64_t iters = (npairs + gstride - 1) / gstride;
    for (uint64_t it = 0; it < iters; ++it)
    {
        const uint64_t q = first + it * gstride;
        uint32_t probe[2] = {0, 0}, keys[2] = {0, 0};
        const uint64_t row0 = q * 2;
        if (q < npairs)
        {
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
#pragma unroll
            for (int k = 0; k < 2; ++k)
            {
                int gx, gy, gz;
                uint32_t key;
                if ((k == 1 && !full) || (a.flags && !(a.flags[row0 + k] & 1u)) || !cell_of(a, p[k], gx, gy, gz, key)) continue;
                probe[k] = probe_neighbours(a, gx, gy, gz);
                keys[k] = key;
            }
        }
        // append the flagged aircraft (cell key | neighbour mask << 32) to the step's list: one atomic per warp iteration that saw any
        const uint32_t mine = (probe[0] ? 1u : 0u) + (probe[1] ? 1u : 0u);
        if (!__any_sync(kFullMask, mine != 0)) continue;
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t up = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= static_cast<uint32_t>(o)) incl += up;
        }
        const uint32_t total = __shfl_sync(kFullMask, incl, 31);
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, static_cast<unsigned long long>(total));
        base = __shfl_sync(kFullMask, base, 0);
        unsigned long long slot = base + (incl - mine);
        if (probe[0]) a.list[slot++] = static_cast<unsigned long long>(keys[0]) | (static_cast<unsigned long long>(probe[0]) << 32);
        if (probe[1]) a.list[slot] = static_cast<unsigned long long>(keys[1]) | (static_cast<unsigned long long>(probe[1]) << 32);
    }
}

// PreUpdate, pass 3: the flagged aircraft enter the exact per-cell index (one list entry per thread: full warps).
__global__ void __launch_bounds__(kThreads) air_insert_kernel(AirArgs a)
{
    const TableView tv = table_view(a);
    const unsigned long long nf = a.acc[4 + (a.step & 1u)];
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nf; i += gstride)
    {
        index_insert(a, tv, static_cast<uint32_t>(a.list[i]));
    }
}

// check_conflicts + move_aircraft; one thread = 2 consecutive rows (one 128-bit position load/store)
__global__ void __launch_bounds__(kThreads) air_step_kernel(AirArgs a)
{
    const uint64_t npairs = (a.n + 1) >> 1;
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    uint32_t conflicts = 0, indexed = 0;
    if (a.do_check)
    {
        // check_conflicts (:124-162): exact look-ups for the neighbour cells the pre-filter flagged, one list entry per
        // thread (the list holds cell keys: nothing here reads the positions this kernel is about to move)
        const TableView tv = table_view(a);
        const unsigned long long nf = a.acc[4 + (a.step & 1u)];
        for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nf; i += gstride)
        {
            const unsigned long long e = a.list[i];
            const uint32_t key = static_cast<uint32_t>(e);
            uint32_t mask = static_cast<uint32_t>(e >> 32);
            while (mask)
            {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1;
                conflicts += index_count(a, tv, neighbour_key(a, key, j));
            }
        }
        // the other parity's counter is the next step's: leave it at zero
        if (blockIdx.x == 0 && threadIdx.x == 0) a.acc[4 + ((a.step + 1u) & 1u)] = 0;
    }
    for (uint64_t q = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < npairs; q += gstride)
    {
        const uint64_t row0 = q * 2;
        const bool full = row0 + 2 <= a.n;
        uint2 p[2];
        uint32_t t2;
        if (full)
        {
            const uint4 v = *reinterpret_cast<const uint4*>(a.pos + row0);
            p[0] = make_uint2(v.x, v.y);
            p[1] = make_uint2(v.z, v.w);
            t2 = *reinterpret_cast<const uint32_t*>(a.target + row0);
        }
        else
        {
            p[0] = a.pos[row0];
            p[1] = p[0];
            t2 = static_cast<uint16_t>(a.target[row0]);
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
                if (cell_of(a, p[k], gx, gy, gz, key)) indexed += 1; // aircraft outside the bounds were reported by the bits kernel
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
            if (a.bits_next)
            {
                // the next PreUpdate's first pass, done here while the new position is in registers
                int gx, gy, gz;
                uint32_t key;
                if (cell_of(a, p[k], gx, gy, gz, key))
                {
                    const uint32_t h = bit_of(a, gx, gy, gz);
                    atomicOr(&a.bits_next[h >> 5], 1u << (h & 31u));
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
void launch_air_candidates(const AirArgs& a, cudaStream_t st) { air_candidates_kernel<<<blocks_for((a.n + 1) / 2), kThreads, 0, st>>>(a); }
void launch_air_insert(const AirArgs& a, cudaStream_t st) { air_insert_kernel<<<kNumSMs * 8, kThreads, 0, st>>>(a); }
void launch_air_step(const AirArgs& a, cudaStream_t st)
{
    air_step_kernel<<<blocks_for((a.n + 1) / 2), kThreads, 0, st>>>(a);
}

} // namespace incerto
