// traders: examples/traders.rs
//   traders_may_buy_shares      :165-197
//   traders_may_sell_shares     :199-236
//   traders_calculate_net_worth :239-263
//   stocks_price_change         :140-163
// Pinned order: buy -> sell -> net worth (one fused pass per trader, stocks visited in
// StockId order) -> price change (second kernel), StockDelisted inserted at the end of the step.
//
// Device layout: cash f64, net worth f64, portfolio = 4 bits per stock (#shares, 0 = not owned; the
// reference's HashMap<StockId, usize>) stored word-major (word w of every trader contiguous) so that a
// warp reads 128 contiguous bytes per word.  The stock table (price f64 + delisted) of the replica is
// staged in shared memory.  Algorithmic bytes per trader-update at 100 stocks: 124 (portfolio 52 r +
// up to 52 w - only changed words are written -, cash 8 + 8, net worth 8 w).  The kernel is bound by
// the integer pipe (Philox), not by HBM: see DESIGN.md §5.
//
// Draw contract: two-level Bernoulli.
//   buy   (trader t, stock s), p <= 1/32: L1 = s wins the block lottery (lottery.cuh, P = 2^-k with the largest k in
//         5..8 such that 2^k p <= 1) of block t, sub-stream 0, key[TBUY] - ONE call per trader and step;
//         L2 = philox(key[TBUY], (t, s, step, 1)): decides = L2.x < floor(2^k p * 2^32),
//         num_shares = min + mulhi(L2.y, max - min + 1)                              (:178, :182)
//         p > 1/32: L1 = byte (s & 15) of philox(key[TBUY], (t, s >> 4, step, 0)) < T1 = min(256, ceil(256 p)),
//         L2 against floor(256 p / T1 * 2^32)
//   sell  same with key[TSEL]; drawn for owned stocks only                          (:218)
//   price stock s: z = det_normal(philox(key[TPRC], (s, 0, step, 0))) (detmath.cuh); change = mean + sd * z (:148-155)
#include <algorithm>

#include "common.cuh"
#include "detmath.cuh"
#include "kernels_traders.hpp"
#include "lottery.cuh"

namespace incerto
{
namespace
{

constexpr int kThreads = 128;

__device__ __forceinline__ uint32_t byte_of(const uint4& v, uint32_t b)
{
    const uint32_t w = word_of(v, b >> 2);
    return (w >> (8 * (b & 3))) & 0xffu;
}

// Level-1 byte tests, SWAR (the emulated __vcmpltu4 of round 1 cost as many instructions as the Philox call that
// produced the word): 0x80 in every byte of x that is < t, for a threshold t in 1..=256 fixed per launch.
//   t <= 128:  byte < t  <=>  high bit clear  and  (low 7 bits + 128 - t) does not reach the high bit
//   t <  256:  byte < t  <=>  high bit clear  or   (low 7 bits + 256 - t) does not reach the high bit
struct ByteLess
{
    uint32_t addc, mode; // mode 0: t <= 128, 1: 128 < t < 256, 2: t >= 256 (always true)
    __device__ __forceinline__ explicit ByteLess(uint32_t t)
    {
        mode = t >= 256u ? 2u : (t > 128u ? 1u : 0u);
        addc = (mode == 1u ? 256u - t : 128u - (t > 128u ? 128u : t)) * 0x01010101u;
    }
    __device__ __forceinline__ uint32_t mask80(uint32_t x) const
    {
        if (mode == 2u) return 0x80808080u;
        const uint32_t y = (x & 0x7f7f7f7fu) + addc;
        return (mode == 0u ? (~x & ~y) : (~x | ~y)) & 0x80808080u;
    }
    // bit k of the result: byte k of x is below the threshold (the multiply gathers bits 0, 8, 16, 24 into 24..27)
    __device__ __forceinline__ uint32_t bits4(uint32_t x) const { return ((mask80(x) >> 7) * 0x01020408u) >> 24; }
    // the sixteen bytes of a Philox call
    __device__ __forceinline__ uint32_t bits16(const uint4& v) const
    {
        return bits4(v.x) | (bits4(v.y) << 4) | (bits4(v.z) << 8) | (bits4(v.w) << 12);
    }
};
// bit k of the result: nibble k of w is non-zero
__device__ __forceinline__ uint32_t nibbles_nonzero_bits(uint32_t w)
{
    uint32_t y = (w | (w >> 1) | (w >> 2) | (w >> 3)) & 0x11111111u;
    y = (y | (y >> 3)) & 0x03030303u;
    y = (y | (y >> 6)) & 0x000f000fu;
    return (y | (y >> 12)) & 0xffu;
}

// lowest set bit of a multi-word mask, cleared on return; -1 when empty
template <int NC>
__device__ __forceinline__ int pop_lowest(unsigned long long (&m)[NC])
{
    int s = -1;
#pragma unroll
    for (int c = 0; c < NC; ++c)
        if (s < 0 && m[c])
        {
            s = 64 * c + __ffsll(static_cast<long long>(m[c])) - 1;
            m[c] &= m[c] - 1;
        }
    return s;
}
template <int NC>
__device__ __forceinline__ void mask_set(unsigned long long (&m)[NC], uint32_t s, bool on)
{
#pragma unroll
    for (int c = 0; c < NC; ++c)
        if (static_cast<uint32_t>(c) == (s >> 6))
        {
            const unsigned long long bit = 1ull << (s & 63u);
            m[c] = on ? (m[c] | bit) : (m[c] & ~bit);
        }
}
template <int NC>
__device__ __forceinline__ uint32_t mask16(const unsigned long long (&m)[NC], int b)
{
    return static_cast<uint32_t>(m[b >> 2] >> (16 * (b & 3))) & 0xffffu;
}

// One thread = one trader.  The trader's portfolio words live in shared memory (column `threadIdx.x`, conflict free:
// the rare-event paths index them by stock; in registers that was a 14-way select per access and 28 registers), next
// to a 1-bit-per-stock ownership mask in registers that buy, sell and net worth share.  The level-1 draws of all
// stocks are taken first (one Philox call per 16 stocks, no divergence) and reduced to candidate bit masks; the rare
// level-2 events are then handled by a short loop in ascending StockId order.
template <int NW>
__global__ void __launch_bounds__(kThreads) traders_step_kernel(TradersArgs a)
{
    pdl_launch_dependents();
    pdl_wait(); // everything below reads what the previous kernel of the step wrote
    constexpr int NB = (NW + 1) / 2;  // 16-stock blocks
    constexpr int NC = (NW + 7) / 8;  // 64-stock mask words
    __shared__ double s_price[kTradersMaxStocks];
    __shared__ uint32_t s_listed[kTradersMaxStocks / 32]; // bit s: stock s exists and is not delisted
    __shared__ uint32_t s_pw[NW][kThreads];
    for (uint32_t i = threadIdx.x; i < kTradersMaxStocks / 32; i += blockDim.x) s_listed[i] = 0;
    __syncthreads();
    for (uint32_t s = threadIdx.x; s < a.num_stocks; s += blockDim.x)
    {
        s_price[s] = a.price[s];
        const uint32_t fl = a.stock_flags[s];
        if ((fl & 1u) && !(fl & a.bit_delisted)) atomicOr(&s_listed[s >> 5], 1u << (s & 31));
    }
    __syncthreads();
    const ByteLess buy_lt(a.buy_t1), sell_lt(a.sell_t1);
    const uint32_t nw_used = (a.num_stocks + 7) >> 3;
    const uint32_t nblk = (a.num_stocks + 15) >> 4;
    const uint32_t tid = threadIdx.x;
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t row = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; row < a.n; row += gstride)
    {
        if (a.trader_flags && !(a.trader_flags[row] & 1u)) continue;
        const uint32_t uid = a.trader_uid ? a.trader_uid[row] : static_cast<uint32_t>(row);
        unsigned long long owned[NC]; // bit s: the trader holds shares of stock s
#pragma unroll
        for (int c = 0; c < NC; ++c) owned[c] = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w)
        {
            const uint32_t v = (static_cast<uint32_t>(w) < nw_used) ? a.portfolio[static_cast<uint64_t>(w) * a.n_cap + row] : 0u;
            s_pw[w][tid] = v;
            owned[w >> 3] |= static_cast<unsigned long long>(nibbles_nonzero_bits(v)) << (8 * (w & 7));
        }
        uint32_t dirty = 0; // bit w: portfolio word w changed
        double cash = a.cash[row];

        if (a.do_buy)
        {
            // candidates: level-1 byte below T1, stock listed, not owned yet (:176-180)
            unsigned long long cand[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) cand[c] = 0;
            if (a.buy_kexp) // warp-uniform: the trader's lottery, winners that are listed and not owned
            {
                lottery256_k(a.rk_buy, a.buy_kexp, uid, 0u, a.step, [&](uint32_t b) { mask_set<NC>(cand, b, true); });
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    cand[c] &= (static_cast<unsigned long long>(s_listed[2 * c]) | (static_cast<unsigned long long>(s_listed[2 * c + 1]) << 32)) & ~owned[c];
            }
            else
            {
                // (the 16-stock blocks of one 64-stock mask word run as a rolled loop: unrolling all of them made the
                // kernel 41 KB of code - past the instruction cache - for no gain, ncu: stall no_instruction 2.0)
#pragma unroll
                for (int c = 0; c < NC; ++c)
                {
#pragma unroll 1
                    for (uint32_t q = 0; q < 4; ++q)
                    {
                        const uint32_t b = 4u * c + q;
                        if (b >= nblk) break;
                        const uint32_t hit = buy_lt.bits16(philox4x32_10(uid, b, a.step, 0u, a.rk_buy));
                        const uint32_t listed = (s_listed[b >> 1] >> (16 * (b & 1))) & 0xffffu;
                        const uint32_t own16 = static_cast<uint32_t>(owned[c] >> (16 * q)) & 0xffffu;
                        cand[c] |= static_cast<unsigned long long>(hit & listed & ~own16) << (16 * q);
                    }
                }
            }
            for (int s = pop_lowest<NC>(cand); s >= 0; s = pop_lowest<NC>(cand))
            {
                const uint4 l2 = philox4x32_10(uid, static_cast<uint32_t>(s), a.step, 1u, a.rk_buy);
                if (!bernoulli(l2.x, a.buy_thr2)) continue;
                const uint32_t num_shares = a.shares_min + uniform_below(l2.y, a.shares_range);
                const double amount = dm_mul(static_cast<double>(num_shares), s_price[s]);
                if (amount > cash) continue; // :184-188
                s_pw[s >> 3][tid] |= num_shares << (4 * (s & 7)); // the nibble was empty: not owned
                dirty |= 1u << (s >> 3);
                mask_set<NC>(owned, static_cast<uint32_t>(s), num_shares != 0);
                cash = dm_sub(cash, amount);
            }
        }

        if (a.do_sell)
        {
            // candidates: owned and level-1 byte below T1 (:216-223); blocks without an owned stock draw nothing
            unsigned long long cand[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) cand[c] = 0;
            if (a.sell_kexp)
            {
                unsigned long long any = 0;
#pragma unroll
                for (int c = 0; c < NC; ++c) any |= owned[c];
                if (any) // a trader without shares draws nothing
                {
                    lottery256_k(a.rk_sell, a.sell_kexp, uid, 0u, a.step, [&](uint32_t b) { mask_set<NC>(cand, b, true); });
#pragma unroll
                    for (int c = 0; c < NC; ++c) cand[c] &= owned[c];
                }
            }
            else
            {
#pragma unroll
                for (int c = 0; c < NC; ++c)
                {
#pragma unroll 1
                    for (uint32_t q = 0; q < 4; ++q)
                    {
                        const uint32_t b = 4u * c + q;
                        const uint32_t own16 = static_cast<uint32_t>(owned[c] >> (16 * q)) & 0xffffu;
                        if (!own16) continue;
                        const uint32_t hit = sell_lt.bits16(philox4x32_10(uid, b, a.step, 0u, a.rk_sell));
                        cand[c] |= static_cast<unsigned long long>(hit & own16) << (16 * q);
                    }
                }
            }
            for (int s = pop_lowest<NC>(cand); s >= 0; s = pop_lowest<NC>(cand))
            {
                if (!bernoulli(philox4x32_10(uid, static_cast<uint32_t>(s), a.step, 1u, a.rk_sell).x, a.sell_thr2)) continue;
                const uint32_t sh = 4 * (s & 7);
                const uint32_t wv = s_pw[s >> 3][tid];
                cash = dm_add(cash, dm_mul(static_cast<double>((wv >> sh) & 0xfu), s_price[s])); // :229-233, ascending StockId
                s_pw[s >> 3][tid] = wv & ~(0xfu << sh);
                dirty |= 1u << (s >> 3);
                mask_set<NC>(owned, static_cast<uint32_t>(s), false);
            }
        }
        double value = 0.0;
        if (a.do_net_worth)
        {
            // :250-261 portfolio value summed in ascending StockId order
            // (the owned stocks are walked 32 at a time: plain 32-bit find-first-set loops instead of the multi-word
            // pop_lowest, which was a fifth of the kernel's instructions in this loop alone)
#pragma unroll
            for (int h = 0; h < 2 * NC; ++h)
            {
                uint32_t rest = static_cast<uint32_t>(owned[h >> 1] >> (32 * (h & 1)));
                while (rest)
                {
                    const uint32_t s = 32u * h + (__ffs(rest) - 1);
                    rest &= rest - 1;
                    const uint32_t n_sh = (s_pw[s >> 3][tid] >> (4 * (s & 7))) & 0xfu;
                    value = dm_add(value, dm_mul(static_cast<double>(n_sh), s_price[s]));
                }
            }
        }

#pragma unroll
        for (int w = 0; w < NW; ++w)
            if ((dirty >> w) & 1u) a.portfolio[static_cast<uint64_t>(w) * a.n_cap + row] = s_pw[w][tid];
        if (a.do_buy || a.do_sell) a.cash[row] = cash;
        if (a.do_net_worth) a.net_worth[row] = dm_add(cash, value); // :261
    }
}

__global__ void stocks_price_change_kernel(TradersArgs a)
{
    pdl_launch_dependents();
    pdl_wait(); // everything below reads what the previous kernel of the step wrote
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.num_stocks) return;
    const uint32_t fl = a.stock_flags[s];
    if (!(fl & 1u) || (fl & a.bit_delisted)) return; // Without<StockDelisted> (:142)
    const uint4 r = philox4x32_10(s, 0u, a.step, 0u, a.rk_price);
    const double change = dm_add(a.change_mean, dm_mul(a.change_std_dev, det_normal(r.x, r.y, r.z, r.w)));
    double p = dm_add(a.price[s], change);
    p = p > 0.0 ? p : 0.0; // .max(0.0) (:155)
    a.price[s] = p;
    if (p < 2.220446049250313e-16) a.stock_flags[s] = static_cast<uint8_t>(fl | a.bit_delisted); // f64::EPSILON (:157-161)
}

__global__ void fill_f64_kernel(double* p, double v, uint64_t n)
{
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += gstride) p[i] = v;
}
__global__ void iota_u64_kernel(uint64_t* p, uint64_t n)
{
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += gstride) p[i] = i;
}

inline int blocks_for(uint64_t items, int threads, int per_sm)
{
    uint64_t b = (items + threads - 1) / threads;
    const uint64_t cap = static_cast<uint64_t>(kNumSMs) * per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<int>(b);
}

} // namespace

void launch_traders_step(const TradersArgs& a, cudaStream_t st)
{
    const uint32_t nw = (a.num_stocks + 7) >> 3;
    const int blocks = blocks_for(a.n, kThreads, 16);
    if (nw <= 4)
        traders_step_kernel<4><<<blocks, kThreads, 0, st>>>(a);
    else if (nw <= 8)
        traders_step_kernel<8><<<blocks, kThreads, 0, st>>>(a);
    else if (nw <= 14)
    {
        static const int wave = wave_blocks(traders_step_kernel<14>, kThreads, 2.0);
        launch_pdl(traders_step_kernel<14>, dim3(std::min(blocks, wave)), dim3(kThreads), 0, st, a);
    }
    else
        traders_step_kernel<32><<<blocks, kThreads, 0, st>>>(a);
}
void launch_stocks_price_change(const TradersArgs& a, cudaStream_t st)
{
    launch_pdl(stocks_price_change_kernel, dim3((a.num_stocks + 127) / 128), dim3(128), 0, st, a);
}
void launch_fill_f64(double* p, double v, uint64_t n, cudaStream_t st)
{
    fill_f64_kernel<<<blocks_for(n, 256, 16), 256, 0, st>>>(p, v, n);
}
void launch_iota_u64(uint64_t* p, uint64_t n, cudaStream_t st)
{
    iota_u64_kernel<<<blocks_for(n, 256, 16), 256, 0, st>>>(p, n);
}

} // namespace incerto
