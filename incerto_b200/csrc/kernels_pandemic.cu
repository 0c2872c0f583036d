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
// Algorithmic bytes per person-update: 10 (position 4+4, flags 1+1); the per-cell infected lists
// (head 4 B/cell stamped with the step, next 4 B/infected) are scratch traffic reported separately.
//
// Draw contract:
//   spawn   person uid: r = philox(key[PSPW], (uid,0,0,0)); x = mulhi(r.x,G), y = mulhi(r.y,G), infected = r.z < thr(p0)
//   move    word w = word (uid&3) of philox(key[PMOV], (uid>>2,0,step,0)); the three p = 1/2 draws of :121-148
//           are bits 31, 30, 29 of w ("true" = bit clear, i.e. (w << k) < 2^31): stay, x-axis, negative direction
//   fate    two-level when 256*p <= 1 for both probabilities: L1 = philox(key[PFAT], (uid>>3,0,step,0)), byte
//           2*(uid&7) (die) / 2*(uid&7)+1 (recover) must be zero; L2 = philox(key[PFAT], (uid,0,step,1)):
//           die = L2.x < thr(256 p_die), recover = L2.y < thr(256 p_rec).  Otherwise L2 alone with thr(p).
//   infect  pair (healthy h, infected i) in the same cell: word 0 of philox(key[PINF], (uid_h, uid_i, step, 0)) < thr(p)
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

// people_move + building the per-cell lists of infected people for the interaction kernel.
// One thread = 4 consecutive rows (one 128-bit position load/store, one Philox call when uid == row).
__global__ void __launch_bounds__(kThreads) pandemic_move_kernel(PandemicArgs a)
{
    const uint64_t nquads = (a.n + 3) >> 2;
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    const uint32_t G = a.grid_size;
    const uint32_t tag = (a.step & 0xffu) << 24;
    for (uint64_t q = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < nquads; q += gstride)
    {
        const uint64_t row0 = q * 4;
        const bool full = row0 + 4 <= a.n;
        uint32_t p[4] = {0, 0, 0, 0};
        uint32_t f4 = 0;
        if (full)
        {
            const uint4 pv = *reinterpret_cast<const uint4*>(a.pos + row0);
            p[0] = pv.x;
            p[1] = pv.y;
            p[2] = pv.z;
            p[3] = pv.w;
            f4 = *reinterpret_cast<const uint32_t*>(a.flags + row0);
        }
        else
        {
            for (uint64_t k = 0; row0 + k < a.n; ++k)
            {
                p[k] = a.pos[row0 + k];
                f4 |= static_cast<uint32_t>(a.flags[row0 + k]) << (8 * k);
            }
        }
        uint4 r = make_uint4(0, 0, 0, 0);
        if (a.do_move && a.uid == nullptr) r = philox4x32_10(static_cast<uint32_t>(q), 0u, a.step, 0u, a.rk_move);
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const uint32_t fl = (f4 >> (8 * k)) & 0xffu;
            if (!(fl & 1u)) continue; // despawned
            uint32_t x = p[k] & 0xffffu, y = p[k] >> 16;
            if (a.do_move)
            {
                uint32_t w;
                if (a.uid == nullptr)
                    w = k == 0 ? r.x : (k == 1 ? r.y : (k == 2 ? r.z : r.w));
                else
                {
                    const uint32_t u = a.uid[row0 + k];
                    w = word_of(philox4x32_10(u >> 2, 0u, a.step, 0u, a.rk_move), u & 3u);
                }
                // :121 stay put; :125 x axis else y axis; :128/:147 negative else positive direction; saturating
                if (w & 0x80000000u)
                {
                    const bool neg = !(w & 0x20000000u);
                    if (!(w & 0x40000000u))
                        x = neg ? (x > 0 ? x - 1 : x) : (x < G - 1 ? x + 1 : x);
                    else
                        y = neg ? (y > 0 ? y - 1 : y) : (y < G - 1 ? y + 1 : y);
                }
                p[k] = x | (y << 16);
            }
            if (a.do_infect && (fl & a.bit_infected))
            {
                // per-cell list of infected people (order irrelevant: the draws are keyed by the pair)
                const uint32_t cell = x * G + y;
                const uint32_t row = static_cast<uint32_t>(row0 + k);
                atomicOr(&a.cell_bits[cell >> 5], 1u << (cell & 31));
                const uint32_t old = atomicExch(&a.cell_head[cell], tag | row);
                a.next_infected[row] = ((old & 0xff000000u) == tag) ? (old & 0x00ffffffu) : 0x00ffffffu;
            }
        }
        if (a.do_move)
        {
            if (full)
                *reinterpret_cast<uint4*>(a.pos + row0) = make_uint4(p[0], p[1], p[2], p[3]);
            else
                for (uint64_t k = 0; row0 + k < a.n; ++k) a.pos[row0 + k] = p[k];
        }
    }
}

// people_infect_others + people_infected_may_die + people_infected_may_recover + the command flush.
__global__ void __launch_bounds__(kThreads) pandemic_interact_kernel(PandemicArgs a)
{
    const uint64_t nquads = (a.n + 3) >> 2;
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    const uint32_t G = a.grid_size;
    const uint32_t tag = (a.step & 0xffu) << 24;
    for (uint64_t q = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < nquads; q += gstride)
    {
        const uint64_t row0 = q * 4;
        const bool full = row0 + 4 <= a.n;
        uint32_t f4 = 0;
        if (full)
            f4 = *reinterpret_cast<const uint32_t*>(a.flags + row0);
        else
            for (uint64_t k = 0; row0 + k < a.n; ++k) f4 |= static_cast<uint32_t>(a.flags[row0 + k]) << (8 * k);
        if (!(f4 & 0x01010101u)) continue;
        uint32_t p[4] = {0, 0, 0, 0};
        if (a.do_infect)
        {
            if (full)
            {
                const uint4 pv = *reinterpret_cast<const uint4*>(a.pos + row0);
                p[0] = pv.x;
                p[1] = pv.y;
                p[2] = pv.z;
                p[3] = pv.w;
            }
            else
                for (uint64_t k = 0; row0 + k < a.n; ++k) p[k] = a.pos[row0 + k];
        }
        uint32_t nf4 = f4;
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const uint32_t fl = (f4 >> (8 * k)) & 0xffu;
            if (!(fl & 1u)) continue;
            const uint32_t row = static_cast<uint32_t>(row0 + k);
            const uint32_t uid = a.uid ? a.uid[row] : row;
            uint32_t nfl = fl;
            if (fl & a.bit_infected)
            {
                if (a.do_die | a.do_recover)
                {
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
                        if (rec) nfl &= ~a.bit_infected;   // commands.entity(e).remove::<Infected>() (:220)
                        if (die) nfl = 0;                   // commands.entity(e).despawn() (:203) wins
                    }
                }
            }
            else if (a.do_infect)
            {
                const uint32_t x = p[k] & 0xffffu, y = p[k] >> 16;
                const uint32_t cell = x * G + y;
                if ((a.cell_bits[cell >> 5] >> (cell & 31)) & 1u)
                {
                    uint32_t h = a.cell_head[cell];
                    uint32_t cur = ((h & 0xff000000u) == tag) ? (h & 0x00ffffffu) : 0x00ffffffu;
                    bool infected = false;
                    // (the guard only bounds the walk should the list ever be corrupt: never hang the device)
                    for (uint32_t guard = 0; cur != 0x00ffffffu && guard < 0x01000000u; ++guard)
                    {
                        const uint32_t iuid = a.uid ? a.uid[cur] : cur;
                        const uint4 r = philox4x32_10(uid, iuid, a.step, 0u, a.rk_infect);
                        infected = infected || bernoulli(r.x, a.thr_infect); // :184-188; every pair draws
                        cur = a.next_infected[cur];
                    }
                    if (infected) nfl |= a.bit_infected; // commands.entity(healthy).insert(Infected) (:187)
                }
            }
            nf4 = (nf4 & ~(0xffu << (8 * k))) | (nfl << (8 * k));
        }
        if (nf4 != f4)
        {
            if (full)
                *reinterpret_cast<uint32_t*>(a.flags + row0) = nf4;
            else
                for (uint64_t k = 0; row0 + k < a.n; ++k) a.flags[row0 + k] = static_cast<uint8_t>(nf4 >> (8 * k));
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

void launch_pandemic_spawn(const PandemicArgs& a, cudaStream_t st)
{
    pandemic_spawn_kernel<<<blocks_for(a.n), kThreads, 0, st>>>(a);
}
void launch_pandemic_move(const PandemicArgs& a, cudaStream_t st)
{
    pandemic_move_kernel<<<blocks_for((a.n + 3) / 4), kThreads, 0, st>>>(a);
}
void launch_pandemic_interact(const PandemicArgs& a, cudaStream_t st)
{
    pandemic_interact_kernel<<<blocks_for((a.n + 3) / 4), kThreads, 0, st>>>(a);
}

} // namespace incerto
