// pandemic: examples/pandemic.rs (SIS with death)
//   spawn_people                :84-107
//   people_move                 :115-166
//   people_infect_others        :169-192   (reference: O(infected x healthy) nested loop)
//   people_infected_may_die     :195-206
//   people_infected_may_recover :209-223
// Pinned order: move -> infect -> die -> recover, commands applied after all four:
// a despawn wins over a recovery, and a person infected in this step can neither die nor recover in it
// because die/recover iterate the pre-flush With<Infected> set.
//
// Device layout: Position{x,y} packed as one u32 per person (x | y << 16), flags u8 (alive + marker bits).
// Algorithmic bytes per person-update: 10 (position 4+4, flags 1+1); the per-cell infected lists (a hashed head
// table of ~n/4 slots stamped with the step - small enough to stay in the L2 -, next 4 B/infected, 1 bit/cell
// "somebody infected stands here") are scratch traffic reported separately.
//
// Draw contract:
//   spawn   person uid: r = philox(key[PSPW], (uid,0,0,0)); x = mulhi(r.x,G), y = mulhi(r.y,G), infected = r.z < thr(p0)
//   move    nibble (uid&7) of word ((uid>>3)&3) of philox(key[PMOV], (uid>>5,0,step,0)): the three p = 1/2 draws of
//           :121-148 are bits 3, 2, 1 of the nibble ("true" = bit clear): stay, x-axis, negative direction - one call
//           serves 32 people, one word the eight consecutive rows a thread owns
//   fate    two-level when 256*p <= 1 for both probabilities: L1 = philox(key[PFAT], (uid>>3,0,step,0)), byte
//           2*(uid&7) (die) / 2*(uid&7)+1 (recover) must be zero; L2 = philox(key[PFAT], (uid,0,step,1)):
//           die = L2.x < thr(256 p_die), recover = L2.y < thr(256 p_rec).  Otherwise L2 alone with thr(p).
//   infect  pair (healthy h, infected i) in the same cell: word 0 of philox(key[PINF], (uid_h, uid_i, step, 0)) < thr(p)
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels_pandemic.hpp"

namespace incerto
{
namespace
{

constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t byte_of(const uint4& v, uint32_t b)
{
    const uint32_t w = word_of(v, b >> 2);
    return (w >> (8 * (b & 3))) & 0xffu;
}

__global__ void __launch_bounds__(kThreads) pandemic_spawn_kernel(PandemicArgs a)
{
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n; i += gstride)
    {
        const uint4 r = philox4x32_10(static_cast<uint32_t>(i), 0u, 0u, 0u, a.rk_spawn);
        const uint32_t x = uniform_below(r.x, a.grid_size), y = uniform_below(r.y, a.grid_size);
        a.pos[i] = x | (y << 16);
        a.flags[i] = static_cast<uint8_t>(1u | a.bit_person | (bernoulli(r.z, a.thr_start_infected) ? a.bit_infected : 0u));
    }
}

constexpr int kRows = 8; // consecutive rows per thread: two 128-bit position loads, one 64-bit flags load

struct Rows8
{
    uint32_t p[kRows];
    unsigned long long f; // flags bytes
};

__device__ __forceinline__ void load_rows(const PandemicArgs& a, uint64_t row0, bool need_pos, Rows8& r)
{
    if (row0 + kRows <= a.n)
    {
        r.f = *reinterpret_cast<const unsigned long long*>(a.flags + row0);
        if (need_pos)
        {
            const uint4 v0 = *reinterpret_cast<const uint4*>(a.pos + row0), v1 = *reinterpret_cast<const uint4*>(a.pos + row0 + 4);
            r.p[0] = v0.x;
            r.p[1] = v0.y;
            r.p[2] = v0.z;
            r.p[3] = v0.w;
            r.p[4] = v1.x;
            r.p[5] = v1.y;
            r.p[6] = v1.z;
            r.p[7] = v1.w;
        }
        return;
    }
    r.f = 0;
#pragma unroll
    for (int k = 0; k < kRows; ++k)
    {
        r.p[k] = 0;
        if (row0 + k < a.n)
        {
            r.f |= static_cast<unsigned long long>(a.flags[row0 + k]) << (8 * k);
            if (need_pos) r.p[k] = a.pos[row0 + k];
        }
    }
}

__device__ __forceinline__ uint32_t head_slot(const PandemicArgs& a, uint32_t cell)
{
    // exact table (one slot per cell) when the grid is small, else a multiplicative hash into ~n/4 slots: a slot's list
    // may then hold infected people of several cells and the walker compares positions
    return a.head_shift ? (cell * 0x9E3779B1u) >> a.head_shift : cell;
}

// people_move + building the per-cell lists of infected people for the interaction kernel.
// One thread = 8 consecutive rows = (uid == row) one word of a Philox call.  All memory operations of the tile are
// issued before any result is consumed: the returning exchanges that link the infected into their cells overlap.
__global__ void __launch_bounds__(kThreads) pandemic_move_kernel(PandemicArgs a)
{
    pdl_launch_dependents();
    pdl_wait(); // positions and flags are the previous kernel's (interact of the step before)
    const uint64_t ntiles = (a.n + kRows - 1) / kRows;
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    const uint32_t G = a.grid_size;
    const uint32_t tag = ((a.step & 0xffu) << 24) & ~a.link_mask; // no stamp in the heads of a population of >= 2^24 rows
    for (uint64_t q = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < ntiles; q += gstride)
    {
        const uint64_t row0 = q * kRows;
        const bool full = row0 + kRows <= a.n;
        Rows8 r;
        load_rows(a, row0, true, r);
        uint32_t nib8 = 0; // nibble k: the draws of row k
        if (a.do_move)
        {
            if (a.uid == nullptr)
                nib8 = word_of(philox4x32_10(static_cast<uint32_t>(q >> 2), 0u, a.step, 0u, a.rk_move), static_cast<uint32_t>(q) & 3u);
            else
            {
#pragma unroll
                for (int k = 0; k < kRows; ++k)
                    if ((r.f >> (8 * k)) & 1ull)
                    {
                        const uint32_t u = a.uid[row0 + k];
                        const uint32_t w = word_of(philox4x32_10(u >> 5, 0u, a.step, 0u, a.rk_move), (u >> 3) & 3u);
                        nib8 |= ((w >> (4 * (u & 7u))) & 0xfu) << (4 * k);
                    }
            }
        }
        if (a.do_move)
        {
#pragma unroll
            for (int k = 0; k < kRows; ++k)
            {
                // :121 stay put (bit 3 clear); :125 x axis (bit 2 clear) else y axis; :128/:147 negative (bit 1 clear)
                // else positive direction; saturating at 0 and G-1.  Branch-free: the stepped coordinate t = c +- 1 is in
                // range iff (unsigned) t < G (c - 1 wraps to 2^32 - 1 at the lower wall), and the halves of the packed
                // word never carry into each other because the move is refused at the walls.
                const uint32_t fl = static_cast<uint32_t>(r.f >> (8 * k)) & 0xffu;
                const uint32_t nib = nib8 >> (4 * k);
                const uint32_t sh = (nib & 4u) << 2;                               // 0: x, 16: y
                const uint32_t dlt = (nib & 2u) - 1u;                              // +1 / -1
                const uint32_t t = ((r.p[k] >> sh) & 0xffffu) + dlt;
                const uint32_t go = (t < G ? 1u : 0u) & (nib >> 3) & fl;           // in range, moves, alive (bit 0)
                r.p[k] += (dlt << sh) & (0u - (go & 1u));
            }
            if (full)
            {
                *reinterpret_cast<uint4*>(a.pos + row0) = make_uint4(r.p[0], r.p[1], r.p[2], r.p[3]);
                *reinterpret_cast<uint4*>(a.pos + row0 + 4) = make_uint4(r.p[4], r.p[5], r.p[6], r.p[7]);
            }
            else
                for (uint64_t k = 0; row0 + k < a.n; ++k) a.pos[row0 + k] = r.p[k];
        }
        if (!a.do_infect) continue;
        // Per-cell lists of infected people (order irrelevant: the draws are keyed by the pair).  The infected are 5-10 %
        // of the rows: instead of one predicated pass per row position (eight passes, each run by the warp whenever ANY
        // lane has an infected person at that position - almost always -, half of this kernel's instructions), every
        // lane walks the set bits of its own 8-bit "alive and infected" mask.  Three unrolled passes keep up to three
        // returning exchanges per lane in flight before any result is used; the rare fourth and later ones follow one
        // at a time.
        const unsigned long long sick8 = r.f & (r.f >> (__ffs(a.bit_infected) - 1)) & 0x0101010101010101ull;
        uint32_t m8 = static_cast<uint32_t>((sick8 * 0x0102040810204080ull) >> 56); // bit k: row k
        auto link = [&](uint32_t k, uint32_t& old_head) {
            const uint32_t p = k & 4u ? (k & 2u ? (k & 1u ? r.p[7] : r.p[6]) : (k & 1u ? r.p[5] : r.p[4]))
                                      : (k & 2u ? (k & 1u ? r.p[3] : r.p[2]) : (k & 1u ? r.p[1] : r.p[0]));
            const uint32_t cell = (p & 0xffffu) * G + (p >> 16);
            atomicOr(&a.cell_bits[cell >> 5], 1u << (cell & 31));
            old_head = atomicExch(&a.cell_head[head_slot(a, cell)], tag | static_cast<uint32_t>(row0 + k));
        };
        auto chain = [&](uint32_t k, uint32_t old_head) {
            a.next_infected[row0 + k] = ((old_head & ~a.link_mask) == tag) ? (old_head & a.link_mask) : a.link_mask;
        };
        uint32_t kk[3] = {8u, 8u, 8u}, oo[3] = {0u, 0u, 0u};
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (m8)
            {
                kk[j] = __ffs(m8) - 1;
                m8 &= m8 - 1;
                link(kk[j], oo[j]);
            }
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (kk[j] < 8u) chain(kk[j], oo[j]);
        while (m8)
        {
            const uint32_t k = __ffs(m8) - 1;
            m8 &= m8 - 1;
            uint32_t o;
            link(k, o);
            chain(k, o);
        }
    }
}

// warp-wide exclusive prefix sum of a small per-lane count; total returned through `total`
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

// people_infect_others + people_infected_may_die + people_infected_may_recover + the command flush.
// A warp streams 256 rows (8 per lane, packed loads).  The rows with real work - the infected (fate draws) and
// the healthy standing in a cell that holds an infected person (pair draws) - are rare and scattered, so they are
// first compacted into a per-warp queue in shared memory and then dealt out one per lane: the Philox paths run
// with full warps instead of once per row position with two or three lanes active.
__global__ void __launch_bounds__(kThreads) pandemic_interact_kernel(PandemicArgs a)
{
    pdl_launch_dependents();
    pdl_wait(); // everything below reads what pandemic_move wrote
    __shared__ uint16_t s_queue[kThreads / 32][2][32 * kRows]; // item = row - tile base | flags << 8 (half the bytes: the L1 gets them)
    __shared__ uint32_t s_qpos[kThreads / 32][32 * kRows];     // packed position of the rows in queue 1
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t warp0 = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
    const uint32_t G = a.grid_size;
    const uint32_t tag = ((a.step & 0xffu) << 24) & ~a.link_mask; // no stamp in the heads of a population of >= 2^24 rows
    const bool fate = (a.do_die | a.do_recover) != 0;
    // the other bitmap (the next step's) is cleared on the side: no memset between the steps
    if (a.cell_bits_clear)
    {
        uint4* z = reinterpret_cast<uint4*>(a.cell_bits_clear);
        const uint64_t gsize = static_cast<uint64_t>(gridDim.x) * blockDim.x;
        for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.bits_vec4; i += gsize)
            z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    for (uint64_t base = warp0 * (32 * kRows); base < a.n; base += nwarps * (32 * kRows))
    {
        const uint64_t row0 = base + lane * kRows;
        Rows8 r;
        r.f = 0;
        if (row0 < a.n) load_rows(a, row0, a.do_infect != 0, r);
        // healthy people: is anybody infected standing in my cell?  (all eight bit loads first)
        uint32_t hit = 0, sick = 0;
        {
            uint32_t bw[kRows];
#pragma unroll
            for (int k = 0; k < kRows; ++k)
            {
                const uint32_t fl = static_cast<uint32_t>(r.f >> (8 * k)) & 0xffu;
                bw[k] = 0;
                if (fl & 1u)
                {
                    if (fl & a.bit_infected)
                        sick |= 1u << k;
                    else if (a.do_infect)
                    {
                        const uint32_t cell = (r.p[k] & 0xffffu) * G + (r.p[k] >> 16);
                        bw[k] = (a.cell_bits[cell >> 5] >> (cell & 31)) & 1u;
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < kRows; ++k) hit |= bw[k] << k;
        }
        if (!fate) sick = 0;
        uint32_t n_sick, n_hit;
        uint32_t o_sick = warp_exclusive_scan(__popc(sick), n_sick);
        uint32_t o_hit = warp_exclusive_scan(__popc(hit), n_hit);
        if (!(n_sick | n_hit)) continue;
#pragma unroll
        for (int k = 0; k < kRows; ++k)
        {
            const uint16_t item = static_cast<uint16_t>((lane * kRows + k) | ((static_cast<uint32_t>(r.f >> (8 * k)) & 0xffu) << 8));
            if ((sick >> k) & 1u) s_queue[warp][0][o_sick++] = item;
            if ((hit >> k) & 1u)
            {
                s_qpos[warp][o_hit] = r.p[k];
                s_queue[warp][1][o_hit++] = item;
            }
        }
        __syncwarp();
        for (uint32_t j = lane; j < n_sick; j += 32)
        {
            const uint32_t item = s_queue[warp][0][j];
            const uint32_t row = static_cast<uint32_t>(base) + (item & 0xffu), fl = item >> 8;
            const uint32_t uid = a.uid ? a.uid[row] : row;
            bool l1_die = true, l1_rec = true;
            if (a.fate_two_level)
            {
                const uint4 l1 = philox4x32_10(uid >> 3, 0u, a.step, 0u, a.rk_fate);
                l1_die = byte_of(l1, 2 * (uid & 7u)) == 0;
                l1_rec = byte_of(l1, 2 * (uid & 7u) + 1) == 0;
            }
            if ((l1_die && a.do_die) || (l1_rec && a.do_recover))
            {
                const uint4 l2 = philox4x32_10(uid, 0u, a.step, 1u, a.rk_fate);
                const bool die = a.do_die && l1_die && bernoulli(l2.x, a.thr_die);
                const bool rec = a.do_recover && l1_rec && bernoulli(l2.y, a.thr_recover);
                uint32_t nfl = fl;
                if (rec) nfl &= ~a.bit_infected;   // commands.entity(e).remove::<Infected>() (:220)
                if (die) nfl = 0;                   // commands.entity(e).despawn() (:203) wins
                if (nfl != fl) a.flags[row] = static_cast<uint8_t>(nfl);
            }
        }
        for (uint32_t j = lane; j < n_hit; j += 32)
        {
            const uint32_t item = s_queue[warp][1][j];
            const uint32_t row = static_cast<uint32_t>(base) + (item & 0xffu), fl = item >> 8;
            const uint32_t uid = a.uid ? a.uid[row] : row;
            const uint32_t p = s_qpos[warp][j];
            const uint32_t cell = (p & 0xffffu) * G + (p >> 16);
            const uint32_t h = a.cell_head[head_slot(a, cell)];
            uint32_t cur = ((h & ~a.link_mask) == tag) ? (h & a.link_mask) : a.link_mask;
            bool infected = false;
            // (the guard only bounds the walk should the list ever be corrupt: never hang the device)
            for (uint64_t guard = 0; cur != a.link_mask && guard < a.n; ++guard)
            {
                const uint32_t nxt = a.next_infected[cur];
                if (!a.head_shift || a.pos[cur] == p) // a hashed slot also lists the infected of other cells
                {
                    const uint32_t iuid = a.uid ? a.uid[cur] : cur;
                    const uint4 rr = philox4x32_10(uid, iuid, a.step, 0u, a.rk_infect);
                    infected = infected || bernoulli(rr.x, a.thr_infect); // :184-188; every pair draws
                }
                cur = nxt;
            }
            if (infected) a.flags[row] = static_cast<uint8_t>(fl | a.bit_infected); // commands.entity(healthy).insert(Infected) (:187)
        }
        __syncwarp();
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

void launch_pandemic_spawn(const PandemicArgs& a, cudaStream_t st)
{
    pandemic_spawn_kernel<<<blocks_for(a.n), kThreads, 0, st>>>(a);
}
void launch_pandemic_move(const PandemicArgs& a, cudaStream_t st)
{
    static const bool plain = [] { const char* e = std::getenv("INCERTO_B200_NO_PDL"); return e && e[0] == '1'; }();
    if (plain)
        pandemic_move_kernel<<<blocks_for((a.n + kRows - 1) / kRows), kThreads, 0, st>>>(a);
    else
    {
        static const int wave = wave_blocks(pandemic_move_kernel, kThreads, 2.0);
        launch_pdl(pandemic_move_kernel, dim3(std::min(blocks_for((a.n + kRows - 1) / kRows), wave)), dim3(kThreads), 0, st, a);
    }
}
void launch_pandemic_interact(const PandemicArgs& a, cudaStream_t st)
{
    static const bool plain = [] { const char* e = std::getenv("INCERTO_B200_NO_PDL"); return e && e[0] == '1'; }();
    if (plain)
        pandemic_interact_kernel<<<blocks_for((a.n + kRows - 1) / kRows), kThreads, 0, st>>>(a);
    else
    {
        static const int wave = wave_blocks(pandemic_interact_kernel, kThreads, 2.0);
        launch_pdl(pandemic_interact_kernel, dim3(std::min(blocks_for((a.n + kRows - 1) / kRows), wave)), dim3(kThreads), 0, st, a);
    }
}

} // namespace incerto
