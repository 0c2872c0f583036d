// coin_toss: examples/coin_toss.rs:67-81 (system) and :33-49 (aggregate).
//
// Draw contract: the toss of entity `uid` at step `t` is word
// (uid & 3) of philox(key[COIN], ctr = (uid >> 2, 0, t, 0)); heads iff
// word < floor(ODDS_HEADS * 2^32).  One Philox call therefore serves the four
// entities a thread updates with one 128-bit load/store per column.
//
// Algorithmic bytes per entity-update: 16 (u32 heads + u32 tosses, read+write).
#include <algorithm>

#include "common.cuh"
#include "kernels.hpp"
#include "../../include/incerto_b200.h"

namespace incerto
{
namespace
{

constexpr int kCoinThreads = 256;

__device__ __forceinline__ uint4 ld_v4_rw(const void* p)
{
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
}

// Fast path: uid == row, every row alive, u32 columns. One thread = one quad.
__global__ void __launch_bounds__(kCoinThreads)
    coin_toss_quad_kernel(uint32_t* __restrict__ heads, uint32_t* __restrict__ tosses, uint64_t n, PhiloxKey key,
                          uint64_t threshold, uint32_t first_step, uint32_t steps)
{
    const uint64_t nquads = n >> 2;
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t q = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < nquads; q += gstride)
    {
        uint4 h = ld_v4_rw(heads + 4 * q);
        uint4 t = ld_v4_rw(tosses + 4 * q);
        for (uint32_t s = 0; s < steps; ++s)
        {
            const uint4 r = philox4x32_10(static_cast<uint32_t>(q), 0u, first_step + s, 0u, key);
            h.x += bernoulli(r.x, threshold) ? 1u : 0u;
            h.y += bernoulli(r.y, threshold) ? 1u : 0u;
            h.z += bernoulli(r.z, threshold) ? 1u : 0u;
            h.w += bernoulli(r.w, threshold) ? 1u : 0u;
        }
        t.x += steps;
        t.y += steps;
        t.z += steps;
        t.w += steps;
        st_stream_v4(heads + 4 * q, h);
        st_stream_v4(tosses + 4 * q, t);
    }
    // tail (n % 4 entities): one thread
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        for (uint64_t i = nquads * 4; i < n; ++i)
        {
            uint32_t h = heads[i];
            for (uint32_t s = 0; s < steps; ++s)
            {
                const uint4 r = philox4x32_10(static_cast<uint32_t>(i >> 2), 0u, first_step + s, 0u, key);
                h += bernoulli(word_of(r, static_cast<uint32_t>(i & 3)), threshold) ? 1u : 0u;
            }
            heads[i] = h;
            tosses[i] += steps;
        }
    }
}

// General path: explicit uids and/or dead rows and/or 64-bit counters.
template <typename T>
__global__ void __launch_bounds__(kCoinThreads)
    coin_toss_rows_kernel(T* __restrict__ heads, T* __restrict__ tosses, const uint32_t* __restrict__ uid,
                          const uint8_t* __restrict__ flags, uint64_t n, PhiloxKey key, uint64_t threshold,
                          uint32_t first_step, uint32_t steps)
{
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += gstride)
    {
        if (flags && !(flags[i] & 1u)) continue;
        const uint32_t u = uid ? uid[i] : static_cast<uint32_t>(i);
        T h = heads[i];
        for (uint32_t s = 0; s < steps; ++s)
        {
            const uint4 r = philox4x32_10(u >> 2, 0u, first_step + s, 0u, key);
            h += bernoulli(word_of(r, u & 3u), threshold) ? 1u : 0u;
        }
        heads[i] = h;
        tosses[i] += steps;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
    coin_odds_kernel(const T* __restrict__ heads, const T* __restrict__ tosses, const uint8_t* __restrict__ flags,
                     uint8_t need_set, uint8_t need_clear, uint64_t n, double* out, double* partials,
                     uint32_t* ticket)
{
    double sum = 0.0;
    double cnt = 0.0;
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += gstride)
    {
        if (flags)
        {
            const uint8_t f = flags[i];
            if ((f & need_set) != need_set || (f & need_clear)) continue;
        }
        sum += static_cast<double>(heads[i]) / static_cast<double>(tosses[i]);
        cnt += 1.0;
    }
    double v[2] = {sum, cnt};
    __shared__ double smem[64];
    block_sum<double, 2>(v, smem);
    if (threadIdx.x == 0)
    {
        partials[2 * blockIdx.x] = v[0];
        partials[2 * blockIdx.x + 1] = v[1];
    }
    if (last_block_arrives(ticket, gridDim.x))
    {
        if (threadIdx.x == 0)
        {
            double s = 0.0, c = 0.0;
            for (uint32_t b = 0; b < gridDim.x; ++b)
            {
                s += partials[2 * b];
                c += partials[2 * b + 1];
            }
            out[0] = s;
            reinterpret_cast<uint64_t*>(out)[1] = static_cast<uint64_t>(c);
        }
    }
}

// The same aggregate for unfiltered u32 columns, 4 entities per lane and load.  Every entity of a run normally has
// the same num_tosses (= the steps taken): when a whole warp agrees on it, the 128 ratios are summed as
// (sum of heads) / tosses - one exact integer sum and one division instead of 128 f64 divisions (the kernel is
// otherwise bound by the fp64 pipe).  The result differs from the entity-by-entity sum only by rounding, which the
// aggregate's contract already leaves open: the reference sums in arbitrary query order (traits.rs:16-19).
__global__ void __launch_bounds__(256)
    coin_odds_quad_kernel(const uint32_t* __restrict__ heads, const uint32_t* __restrict__ tosses, uint64_t n, double* out,
                          double* partials, uint32_t* ticket)
{
    double sum = 0.0;
    double cnt = 0.0;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp0 = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
    const uint64_t nfull = n / 128; // whole warp tiles of 128 entities
    constexpr int kU = 4;           // tiles per warp iteration: eight 128-bit loads in flight per lane before any use
    auto tile = [&](const uint4& h, const uint4& t) {
        const bool same = t.x == t.y && t.x == t.z && t.x == t.w && t.x == __shfl_sync(kFullMask, t.x, 0);
        if (__all_sync(kFullMask, same))
        {
            const unsigned long long hs = warp_sum(static_cast<unsigned long long>(h.x) + h.y + h.z + h.w);
            if (lane == 0) sum += static_cast<double>(hs) / static_cast<double>(t.x);
        }
        else
            sum += static_cast<double>(h.x) / static_cast<double>(t.x) + static_cast<double>(h.y) / static_cast<double>(t.y) +
                   static_cast<double>(h.z) / static_cast<double>(t.z) + static_cast<double>(h.w) / static_cast<double>(t.w);
        cnt += 4.0;
    };
    const uint64_t ngroups = nfull / kU;
    for (uint64_t g = warp0; g < ngroups; g += nwarps)
    {
        uint4 h[kU], t[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u)
        {
            const uint64_t i = (g * kU + u) * 128 + lane * 4;
            h[u] = ld_stream_v4(heads + i);
            t[u] = ld_stream_v4(tosses + i);
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) tile(h[u], t[u]);
    }
    for (uint64_t w = ngroups * kU + warp0; w < nfull; w += nwarps)
    {
        const uint64_t i = w * 128 + lane * 4;
        tile(ld_stream_v4(heads + i), ld_stream_v4(tosses + i));
    }
    for (uint64_t i = nfull * 128 + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    {
        sum += static_cast<double>(heads[i]) / static_cast<double>(tosses[i]);
        cnt += 1.0;
    }
    double v[2] = {sum, cnt};
    __shared__ double smem[64];
    block_sum<double, 2>(v, smem);
    if (threadIdx.x == 0)
    {
        partials[2 * blockIdx.x] = v[0];
        partials[2 * blockIdx.x + 1] = v[1];
    }
    if (last_block_arrives(ticket, gridDim.x))
    {
        if (threadIdx.x == 0)
        {
            double s2 = 0.0, c = 0.0;
            for (uint32_t b = 0; b < gridDim.x; ++b)
            {
                s2 += partials[2 * b];
                c += partials[2 * b + 1];
            }
            out[0] = s2;
            reinterpret_cast<uint64_t*>(out)[1] = static_cast<uint64_t>(c);
        }
    }
}

inline int coin_blocks(uint64_t work_items)
{
    uint64_t b = (work_items + kCoinThreads - 1) / kCoinThreads;
    const uint64_t cap = static_cast<uint64_t>(kNumSMs) * 8 * 4; // 4 resident waves, grid-stride beyond
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<int>(b);
}

} // namespace

void launch_coin_toss(uint32_t* heads, uint32_t* tosses, const uint32_t* uid, const uint8_t* flags, uint64_t n,
                      PhiloxKey key, uint64_t threshold, uint32_t first_step, uint32_t steps, cudaStream_t st)
{
    if (uid == nullptr && flags == nullptr)
        coin_toss_quad_kernel<<<coin_blocks((n + 3) / 4), kCoinThreads, 0, st>>>(heads, tosses, n, key, threshold,
                                                                                 first_step, steps);
    else
        coin_toss_rows_kernel<uint32_t><<<coin_blocks(n), kCoinThreads, 0, st>>>(heads, tosses, uid, flags, n, key,
                                                                                 threshold, first_step, steps);
}

void launch_coin_toss_u64(uint64_t* heads, uint64_t* tosses, const uint32_t* uid, const uint8_t* flags, uint64_t n,
                          PhiloxKey key, uint64_t threshold, uint32_t first_step, uint32_t steps, cudaStream_t st)
{
    coin_toss_rows_kernel<uint64_t><<<coin_blocks(n), kCoinThreads, 0, st>>>(heads, tosses, uid, flags, n, key,
                                                                             threshold, first_step, steps);
}

void launch_coin_odds(const void* heads, const void* tosses, int dtype, const uint8_t* flags, uint8_t need_set,
                      uint8_t need_clear, uint64_t n, double* out, void* scratch, uint32_t* ticket, cudaStream_t st)
{
    uint64_t b = (n + 256 * 8 - 1) / (256 * 8);
    if (b < 1) b = 1;
    if (b > static_cast<uint64_t>(kNumSMs) * 8) b = kNumSMs * 8;
    if (dtype == INCERTO_U32)
    {
        if (flags == nullptr && n >= 128)
        {
            // exactly one resident wave of a grid-stride loop: every warp gets the same share (1184 blocks on the 444
            // slots its 72 registers leave were 2.67 waves, the last one two thirds empty)
            static const int per_sm = [] {
                int o = 0;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, coin_odds_quad_kernel, 256, 0) != cudaSuccess || o < 1) o = 1;
                return o;
            }();
            const uint64_t wave = static_cast<uint64_t>(kNumSMs) * per_sm;
            coin_odds_quad_kernel<<<static_cast<int>(std::min(b, wave)), 256, 0, st>>>(static_cast<const uint32_t*>(heads),
                                                                                     static_cast<const uint32_t*>(tosses), n, out,
                                                                                     static_cast<double*>(scratch), ticket);
        }
        else
            coin_odds_kernel<uint32_t><<<static_cast<int>(b), 256, 0, st>>>(
            static_cast<const uint32_t*>(heads), static_cast<const uint32_t*>(tosses), flags, need_set, need_clear,
            n, out, static_cast<double*>(scratch), ticket);
    }
    else
        coin_odds_kernel<uint64_t><<<static_cast<int>(b), 256, 0, st>>>(
            static_cast<const uint64_t*>(heads), static_cast<const uint64_t*>(tosses), flags, need_set, need_clear,
            n, out, static_cast<double*>(scratch), ticket);
}

} // namespace incerto
