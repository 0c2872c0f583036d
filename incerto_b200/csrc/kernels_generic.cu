// Generic table kernels: filtered count / sum / min / max, radix select
// (Median / Percentile), counter systems, identifier lookup, compaction.
//
// Reference semantics restated here:
//   Simulation::count           src/simulation.rs:163-173
//   Minimum/Maximum/Mean        src/util/aggregate.rs:73-105,157-192
//   Median/Percentile           src/util/aggregate.rs:107-155 (k-th element of the sorted samples)
//   Simulation::sample          src/simulation.rs:43-65 (identifier scan)
#include <cfloat>

#include <algorithm>

#include "common.cuh"
#include "kernels.hpp"
#include "../../include/incerto_b200.h"

namespace incerto
{

namespace
{

constexpr int kReduceThreads = 256;
constexpr int kMaxReduceBlocks = kNumSMs * 8;

__device__ __forceinline__ bool row_passes(const uint8_t* flags, uint64_t i, uint8_t need_set, uint8_t need_clear)
{
    if (flags == nullptr) return true;
    const uint8_t f = flags[i];
    return (f & need_set) == need_set && (f & need_clear) == 0;
}

template <typename T>
struct Widen;
#define INC_WIDEN_UINT(T)                                                                   \
    template <>                                                                             \
    struct Widen<T>                                                                         \
    {                                                                                       \
        using Sum = uint64_t;                                                               \
        static __device__ uint64_t bits(T v) { return static_cast<uint64_t>(v); }           \
        static __device__ uint64_t sum_bits(Sum s) { return s; }                            \
        static __device__ uint64_t key(T v) { return static_cast<uint64_t>(v); }            \
        static __device__ T from_key(uint64_t k) { return static_cast<T>(k); }              \
        static __device__ bool is_nan(T) { return false; }                                  \
    };
#define INC_WIDEN_SINT(T, BITS)                                                             \
    template <>                                                                             \
    struct Widen<T>                                                                         \
    {                                                                                       \
        using Sum = int64_t;                                                                \
        static __device__ uint64_t bits(T v) { return static_cast<uint64_t>(static_cast<int64_t>(v)); } \
        static __device__ uint64_t sum_bits(Sum s) { return static_cast<uint64_t>(s); }     \
        static __device__ uint64_t key(T v)                                                 \
        {                                                                                   \
            return (static_cast<uint64_t>(static_cast<int64_t>(v)) ^ (1ull << (BITS - 1))) & \
                   (BITS == 64 ? ~0ull : ((1ull << (BITS % 64)) - 1));                      \
        }                                                                                   \
        static __device__ T from_key(uint64_t k) { return static_cast<T>(k ^ (1ull << (BITS - 1))); } \
        static __device__ bool is_nan(T) { return false; }                                  \
    };
INC_WIDEN_UINT(uint8_t)
INC_WIDEN_UINT(uint16_t)
INC_WIDEN_UINT(uint32_t)
INC_WIDEN_UINT(uint64_t)
INC_WIDEN_SINT(int8_t, 8)
INC_WIDEN_SINT(int16_t, 16)
INC_WIDEN_SINT(int32_t, 32)
INC_WIDEN_SINT(int64_t, 64)
template <>
struct Widen<float>
{
    using Sum = double;
    static __device__ uint64_t bits(float v) { return static_cast<uint64_t>(__double_as_longlong(static_cast<double>(v))); }
    static __device__ uint64_t sum_bits(Sum s) { return static_cast<uint64_t>(__double_as_longlong(s)); }
    static __device__ uint64_t key(float v)
    {
        const uint32_t b = __float_as_uint(v);
        return (b & 0x80000000u) ? static_cast<uint32_t>(~b) : (b | 0x80000000u);
    }
    static __device__ float from_key(uint64_t k)
    {
        const uint32_t b = static_cast<uint32_t>(k);
        return __uint_as_float((b & 0x80000000u) ? (b & 0x7fffffffu) : ~b);
    }
    static __device__ bool is_nan(float v) { return v != v; }
};
template <>
struct Widen<double>
{
    using Sum = double;
    static __device__ uint64_t bits(double v) { return static_cast<uint64_t>(__double_as_longlong(v)); }
    static __device__ uint64_t sum_bits(Sum s) { return static_cast<uint64_t>(__double_as_longlong(s)); }
    static __device__ uint64_t key(double v)
    {
        const uint64_t b = static_cast<uint64_t>(__double_as_longlong(v));
        return (b >> 63) ? ~b : (b | (1ull << 63));
    }
    static __device__ double from_key(uint64_t k)
    {
        return __longlong_as_double(static_cast<long long>((k >> 63) ? (k & ~(1ull << 63)) : ~k));
    }
    static __device__ bool is_nan(double v) { return v != v; }
};

template <typename T>
struct Local
{
    uint64_t count;
    typename Widen<T>::Sum sum;
    T mn, mx;
    uint32_t nan;
};

template <typename T>
__device__ __forceinline__ void merge(Local<T>& a, const Local<T>& b)
{
    if (b.count)
    {
        if (!a.count)
        {
            a.mn = b.mn;
            a.mx = b.mx;
        }
        else
        {
            a.mn = b.mn < a.mn ? b.mn : a.mn;
            a.mx = b.mx > a.mx ? b.mx : a.mx;
        }
    }
    a.count += b.count;
    a.sum += b.sum;
    a.nan |= b.nan;
}

template <typename T>
__device__ __forceinline__ Local<T> shfl_local(const Local<T>& v, int o)
{
    Local<T> r;
    r.count = __shfl_xor_sync(kFullMask, v.count, o);
    r.sum = __shfl_xor_sync(kFullMask, v.sum, o);
    r.mn = __shfl_xor_sync(kFullMask, v.mn, o);
    r.mx = __shfl_xor_sync(kFullMask, v.mx, o);
    r.nan = __shfl_xor_sync(kFullMask, v.nan, o);
    return r;
}
// 8/16-bit types are not shuffle-able directly
template <>
__device__ __forceinline__ Local<uint8_t> shfl_local(const Local<uint8_t>& v, int o)
{
    Local<uint8_t> r;
    r.count = __shfl_xor_sync(kFullMask, v.count, o);
    r.sum = __shfl_xor_sync(kFullMask, v.sum, o);
    r.mn = static_cast<uint8_t>(__shfl_xor_sync(kFullMask, static_cast<uint32_t>(v.mn), o));
    r.mx = static_cast<uint8_t>(__shfl_xor_sync(kFullMask, static_cast<uint32_t>(v.mx), o));
    r.nan = 0;
    return r;
}
template <>
__device__ __forceinline__ Local<int8_t> shfl_local(const Local<int8_t>& v, int o)
{
    Local<int8_t> r;
    r.count = __shfl_xor_sync(kFullMask, v.count, o);
    r.sum = __shfl_xor_sync(kFullMask, v.sum, o);
    r.mn = static_cast<int8_t>(__shfl_xor_sync(kFullMask, static_cast<int32_t>(v.mn), o));
    r.mx = static_cast<int8_t>(__shfl_xor_sync(kFullMask, static_cast<int32_t>(v.mx), o));
    r.nan = 0;
    return r;
}
template <>
__device__ __forceinline__ Local<uint16_t> shfl_local(const Local<uint16_t>& v, int o)
{
    Local<uint16_t> r;
    r.count = __shfl_xor_sync(kFullMask, v.count, o);
    r.sum = __shfl_xor_sync(kFullMask, v.sum, o);
    r.mn = static_cast<uint16_t>(__shfl_xor_sync(kFullMask, static_cast<uint32_t>(v.mn), o));
    r.mx = static_cast<uint16_t>(__shfl_xor_sync(kFullMask, static_cast<uint32_t>(v.mx), o));
    r.nan = 0;
    return r;
}
template <>
__device__ __forceinline__ Local<int16_t> shfl_local(const Local<int16_t>& v, int o)
{
    Local<int16_t> r;
    r.count = __shfl_xor_sync(kFullMask, v.count, o);
    r.sum = __shfl_xor_sync(kFullMask, v.sum, o);
    r.mn = static_cast<int16_t>(__shfl_xor_sync(kFullMask, static_cast<int32_t>(v.mn), o));
    r.mx = static_cast<int16_t>(__shfl_xor_sync(kFullMask, static_cast<int32_t>(v.mx), o));
    r.nan = 0;
    return r;
}

template <typename T>
__device__ __forceinline__ void to_record(const Local<T>& l, ReduceRecord* r)
{
    r->count = l.count;
    r->sum_bits = Widen<T>::sum_bits(l.sum);
    r->min_bits = l.count ? Widen<T>::bits(l.mn) : 0;
    r->max_bits = l.count ? Widen<T>::bits(l.mx) : 0;
    r->nan_seen = l.nan;
    r->pad = 0;
}

// Per-block partials are combined by the last block in block-index order, so
// the floating-point sum is reproducible for a given launch shape.
template <typename T>
__global__ void __launch_bounds__(kReduceThreads)
    reduce_column_kernel(const T* __restrict__ col, uint32_t stride, uint32_t lane, const uint8_t* __restrict__ flags,
                         uint8_t need_set, uint8_t need_clear, uint64_t n, ReduceRecord* out, Local<T>* partials,
                         uint32_t* ticket)
{
    Local<T> acc;
    acc.count = 0;
    acc.sum = 0;
    acc.mn = T(0);
    acc.mx = T(0);
    acc.nan = 0;
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += gstride)
    {
        if (!row_passes(flags, i, need_set, need_clear)) continue;
        const T v = col[i * stride + lane];
        Local<T> one;
        one.count = 1;
        one.sum = static_cast<typename Widen<T>::Sum>(v);
        one.mn = v;
        one.mx = v;
        one.nan = Widen<T>::is_nan(v) ? 1u : 0u;
        merge(acc, one);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        Local<T> other = shfl_local(acc, o);
        merge(acc, other);
    }
    __shared__ Local<T> s_part[kReduceThreads / 32];
    const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane_id == 0) s_part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        Local<T> b = s_part[0];
        for (int w = 1; w < kReduceThreads / 32; ++w) merge(b, s_part[w]);
        partials[blockIdx.x] = b;
    }
    if (last_block_arrives(ticket, gridDim.x))
    {
        if (threadIdx.x == 0)
        {
            Local<T> t = partials[0];
            for (uint32_t b = 1; b < gridDim.x; ++b) merge(t, partials[b]);
            to_record(t, out);
        }
    }
}

__global__ void __launch_bounds__(kReduceThreads)
    count_rows_kernel(const uint8_t* __restrict__ flags, uint8_t need_set, uint8_t need_clear, uint64_t n,
                      uint64_t* out, uint64_t* partials, uint32_t* ticket)
{
    // 16 rows per 128-bit load; the tail is handled bytewise
    uint64_t cnt = 0;
    const uint64_t nvec = n / 16;
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    const uint32_t set4 = 0x01010101u * need_set, clr4 = 0x01010101u * need_clear;
    for (uint64_t v = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < nvec; v += gstride)
    {
        const uint4 w = ld_stream_v4(flags + v * 16);
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            // byte passes iff ((f & set) ^ set) | (f & clr) == 0
            const uint32_t bad = ((ws[j] & set4) ^ set4) | (ws[j] & clr4);
            // zero-byte detector: bytes of `bad` equal to zero
            const uint32_t z = ~(((bad & 0x7f7f7f7fu) + 0x7f7f7f7fu) | bad) & 0x80808080u;
            cnt += __popc(z);
        }
    }
    const uint64_t tail0 = nvec * 16;
    for (uint64_t i = tail0 + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += gstride)
        cnt += row_passes(flags, i, need_set, need_clear) ? 1 : 0;

    uint64_t v[1] = {cnt};
    __shared__ uint64_t smem[32];
    block_sum<uint64_t, 1>(v, smem);
    if (threadIdx.x == 0) partials[blockIdx.x] = v[0];
    if (last_block_arrives(ticket, gridDim.x))
    {
        if (threadIdx.x == 0)
        {
            uint64_t t = 0;
            for (uint32_t b = 0; b < gridDim.x; ++b) t += partials[b];
            *out = t;
        }
    }
}

// ---------------------------------------------------------------- radix select
// state[0] = key prefix, state[1] = k, state[2] = shift of the current digit,
// state[3] = error flag (k >= count), state[4..260) = histogram
template <typename T>
__global__ void __launch_bounds__(kReduceThreads)
    select_hist_kernel(const T* __restrict__ col, uint32_t stride, uint32_t lane, const uint8_t* __restrict__ flags,
                       uint8_t need_set, uint8_t need_clear, uint64_t n, uint64_t* state)
{
    __shared__ uint32_t s_hist[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const uint64_t prefix = state[0];
    const uint32_t shift = static_cast<uint32_t>(state[2]);
    const uint32_t top = sizeof(T) * 8;
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += gstride)
    {
        if (!row_passes(flags, i, need_set, need_clear)) continue;
        const uint64_t key = Widen<T>::key(col[i * stride + lane]);
        const bool match = (shift + 8 >= top) ? true : ((key >> (shift + 8)) == (prefix >> (shift + 8)));
        if (match) atomicAdd(&s_hist[(key >> shift) & 0xff], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (s_hist[i]) atomicAdd(reinterpret_cast<unsigned long long*>(&state[4 + i]), static_cast<unsigned long long>(s_hist[i]));
}

__global__ void select_pick_kernel(uint64_t* state)
{
    // single thread: 256 bins
    if (threadIdx.x != 0) return;
    uint64_t k = state[1];
    const uint32_t shift = static_cast<uint32_t>(state[2]);
    uint64_t cum = 0;
    int bin = 255;
    for (int i = 0; i < 256; ++i)
    {
        const uint64_t c = state[4 + i];
        if (k < cum + c)
        {
            bin = i;
            break;
        }
        cum += c;
    }
    state[0] |= static_cast<uint64_t>(bin) << shift;
    state[1] = k - cum;
    state[2] = shift >= 8 ? shift - 8 : 0;
    for (int i = 0; i < 256; ++i) state[4 + i] = 0;
}

template <typename T>
__global__ void select_finish_kernel(uint64_t* state)
{
    if (threadIdx.x == 0) state[0] = Widen<T>::bits(Widen<T>::from_key(state[0]));
}

// mode 0: k = arg; mode 1 (Median, aggregate.rs:152): k = count / 2;
// mode 2 (Percentile<P>, aggregate.rs:131): k = floor((P / 100) * count) in f64.
__global__ void select_init_kernel(uint64_t* state, const ReduceRecord* rec, uint32_t mode, uint64_t arg,
                                   uint32_t top_shift)
{
    const int i = threadIdx.x;
    if (i == 0)
    {
        const uint64_t count = rec ? rec->count : ~0ull;
        uint64_t k = arg;
        if (mode == 1) k = count / 2;
        if (mode == 2) k = static_cast<uint64_t>(floor((static_cast<double>(arg) / 100.0) * static_cast<double>(count)));
        uint64_t err = 0;
        if (k >= count)
        {
            err = 1; // index == len: the reference indexes out of range here
            k = count ? count - 1 : 0;
        }
        state[0] = 0;
        state[1] = k;
        state[2] = top_shift;
        state[3] = err;
    }
    state[4 + i] = 0;
}

// ------------------------------------------------------------ counter systems
template <typename T>
__global__ void add_const_kernel(T* col, uint32_t stride, const uint8_t* __restrict__ flags, uint8_t need_set,
                                 uint64_t n, int64_t amount)
{
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += gstride)
    {
        if (!row_passes(flags, i, need_set, 0)) continue;
        col[i * stride] = static_cast<T>(col[i * stride] + static_cast<T>(amount));
    }
}

template <typename D, typename S>
__global__ void add_column_kernel(D* dst, const S* __restrict__ src, const uint8_t* __restrict__ flags, uint64_t n)
{
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += gstride)
    {
        if (!row_passes(flags, i, 1, 0)) continue;
        dst[i] = static_cast<D>(dst[i] + static_cast<D>(src[i]));
    }
}

template <typename T>
__global__ void find_id_kernel(const T* __restrict__ col, uint32_t stride, const uint8_t* __restrict__ flags,
                               uint64_t n, uint64_t id_bits, uint64_t* out)
{
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += gstride)
    {
        if (!row_passes(flags, i, 1, 0)) continue;
        if (Widen<T>::bits(col[i * stride]) == id_bits)
        {
            atomicAdd(reinterpret_cast<unsigned long long*>(&out[0]), 1ull);
            atomicMin(reinterpret_cast<unsigned long long*>(&out[1]), static_cast<unsigned long long>(i));
        }
    }
}

__global__ void iota_kernel(uint32_t* p, uint64_t n)
{
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += gstride)
        p[i] = static_cast<uint32_t>(i);
}

__global__ void flush_kernel(uint4* p, size_t nvec, uint32_t value)
{
    const size_t gstride = static_cast<size_t>(gridDim.x) * blockDim.x;
    const uint4 v = make_uint4(value, value, value, value);
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += gstride) p[i] = v;
}

// second half of the flush: reading lines that are not in the L2 evicts (writes back) the dirty lines the first half
// left behind, so the timed kernel starts with an L2 full of CLEAN foreign lines - a miss then costs a fetch, not a
// write-back plus a fetch
__global__ void flush_read_kernel(const uint4* p, size_t nvec, uint32_t* sink)
{
    const size_t gstride = static_cast<size_t>(gridDim.x) * blockDim.x;
    uint32_t acc = 0;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += gstride)
    {
        const uint4 v = ld_stream_v4(p + i);
        acc |= v.x ^ v.y ^ v.z ^ v.w; // all four words hold the same value: always 0
    }
    if (acc) *sink = acc;
}

__global__ void step_increment_kernel(uint32_t* d_step) { *d_step += 1; }
__global__ void set_u32_kernel(uint32_t* p, uint32_t v) { *p = v; }

// ------------------------------------------------------------------ compaction
// Pass 1: live rows per 1024-row block (ballot + popc). Pass 2: exclusive scan
// over the block counts (one block). Pass 3: each warp recomputes its ballot and
// scatters live rows to block_offset + warp prefix + lane prefix (stable).
constexpr int kCompactThreads = 1024;

__global__ void __launch_bounds__(kCompactThreads)
    compact_count_kernel(const uint8_t* __restrict__ flags, uint64_t n, uint32_t* block_counts)
{
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kCompactThreads + threadIdx.x;
    const bool live = i < n && (flags[i] & 1u);
    const unsigned b = __ballot_sync(kFullMask, live);
    __shared__ uint32_t s_w[32];
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    if (threadIdx.x < 32)
    {
        uint32_t v = warp_sum(s_w[threadIdx.x]);
        if (threadIdx.x == 0) block_counts[blockIdx.x] = v;
    }
}

__global__ void __launch_bounds__(1024) compact_scan_kernel(uint32_t* block_counts, uint32_t nblocks)
{
    // exclusive scan in place; block_counts[nblocks] receives the total
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nblocks; base += 1024)
    {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nblocks ? block_counts[i] : 0;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t y = __shfl_up_sync(kFullMask, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32)
        {
            uint32_t w = s_warp[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t y = __shfl_up_sync(kFullMask, w, o);
                if (threadIdx.x >= o) w += y;
            }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t warp_prefix = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0;
        const uint32_t incl = x + warp_prefix + s_carry;
        if (i < nblocks) block_counts[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_counts[nblocks] = s_carry;
}

__global__ void __launch_bounds__(kCompactThreads)
    compact_scatter_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint32_t row_bytes,
                           const uint8_t* __restrict__ flags, uint64_t n, const uint32_t* __restrict__ block_offsets)
{
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kCompactThreads + threadIdx.x;
    const bool live = i < n && (flags[i] & 1u);
    const unsigned b = __ballot_sync(kFullMask, live);
    __shared__ uint32_t s_w[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_w[warp] = __popc(b);
    __syncthreads();
    if (threadIdx.x < 32)
    {
        uint32_t w = s_w[threadIdx.x];
        uint32_t x = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t y = __shfl_up_sync(kFullMask, x, o);
            if (threadIdx.x >= o) x += y;
        }
        s_w[threadIdx.x] = x - w;
    }
    __syncthreads();
    if (live)
    {
        const uint64_t d = static_cast<uint64_t>(block_offsets[blockIdx.x]) + s_w[warp] + __popc(b & ((1u << lane) - 1u));
        if (row_bytes == 1)
            dst[d] = src[i];
        else if (row_bytes == 2)
            reinterpret_cast<uint16_t*>(dst)[d] = reinterpret_cast<const uint16_t*>(src)[i];
        else if (row_bytes == 4)
            reinterpret_cast<uint32_t*>(dst)[d] = reinterpret_cast<const uint32_t*>(src)[i];
        else if (row_bytes == 8)
            reinterpret_cast<uint64_t*>(dst)[d] = reinterpret_cast<const uint64_t*>(src)[i];
        else
            for (uint32_t k = 0; k < row_bytes; ++k) dst[d * row_bytes + k] = src[i * row_bytes + k];
    }
}

inline int blocks_for(uint64_t n, int threads, int per_thread = 1)
{
    uint64_t b = (n + static_cast<uint64_t>(threads) * per_thread - 1) / (static_cast<uint64_t>(threads) * per_thread);
    if (b < 1) b = 1;
    if (b > static_cast<uint64_t>(kMaxReduceBlocks)) b = kMaxReduceBlocks;
    return static_cast<int>(b);
}

} // namespace

#define INC_DISPATCH_DTYPE(dt, FN, ...)                          \
    switch (dt)                                                  \
    {                                                            \
    case INCERTO_U8: FN(uint8_t, __VA_ARGS__); break;            \
    case INCERTO_U16: FN(uint16_t, __VA_ARGS__); break;          \
    case INCERTO_U32: FN(uint32_t, __VA_ARGS__); break;          \
    case INCERTO_U64: FN(uint64_t, __VA_ARGS__); break;          \
    case INCERTO_I8: FN(int8_t, __VA_ARGS__); break;             \
    case INCERTO_I16: FN(int16_t, __VA_ARGS__); break;           \
    case INCERTO_I32: FN(int32_t, __VA_ARGS__); break;           \
    case INCERTO_I64: FN(int64_t, __VA_ARGS__); break;           \
    case INCERTO_F32: FN(float, __VA_ARGS__); break;             \
    case INCERTO_F64: FN(double, __VA_ARGS__); break;            \
    default: break;                                              \
    }

void launch_reduce_column(const void* col, int dtype, uint32_t stride, uint32_t lane, const uint8_t* flags,
                          uint8_t need_set, uint8_t need_clear, uint64_t n, ReduceRecord* out, void* scratch,
                          uint32_t* ticket, cudaStream_t st)
{
    const int blocks = blocks_for(n, kReduceThreads, 4);
#define INC_RED(T, ...)                                                                                       \
    reduce_column_kernel<T><<<blocks, kReduceThreads, 0, st>>>(static_cast<const T*>(col), stride, lane, flags, \
                                                               need_set, need_clear, n, out,                  \
                                                               static_cast<Local<T>*>(scratch), ticket)
    INC_DISPATCH_DTYPE(dtype, INC_RED, 0)
#undef INC_RED
}

void launch_count_rows(const uint8_t* flags, uint8_t need_set, uint8_t need_clear, uint64_t n, uint64_t* out,
                       void* scratch, uint32_t* ticket, cudaStream_t st)
{
    const int blocks = blocks_for(n, kReduceThreads, 64);
    count_rows_kernel<<<blocks, kReduceThreads, 0, st>>>(flags, need_set, need_clear, n, out,
                                                         static_cast<uint64_t*>(scratch), ticket);
}

void launch_radix_select(const void* col, int dtype, uint32_t stride, uint32_t lane, const uint8_t* flags,
                         uint8_t need_set, uint8_t need_clear, uint64_t n, const ReduceRecord* d_rec, uint32_t mode,
                         uint64_t arg, uint64_t* state, cudaStream_t st, int32_t (*between_passes)(void*), void* hook_ctx)
{
    size_t esize = 1;
    switch (dtype)
    {
    case INCERTO_U16:
    case INCERTO_I16: esize = 2; break;
    case INCERTO_U32:
    case INCERTO_I32:
    case INCERTO_F32: esize = 4; break;
    case INCERTO_U64:
    case INCERTO_I64:
    case INCERTO_F64: esize = 8; break;
    default: break;
    }
    select_init_kernel<<<1, 256, 0, st>>>(state, d_rec, mode, arg, static_cast<uint32_t>(esize * 8 - 8));
    const int blocks = blocks_for(n, kReduceThreads, 4);
    for (size_t pass = 0; pass < esize; ++pass)
    {
#define INC_SEL(T, ...)                                                                                    \
    select_hist_kernel<T><<<blocks, kReduceThreads, 0, st>>>(static_cast<const T*>(col), stride, lane, flags, \
                                                             need_set, need_clear, n, state)
        INC_DISPATCH_DTYPE(dtype, INC_SEL, 0)
#undef INC_SEL
        if (between_passes && between_passes(hook_ctx) != 0) return;
        select_pick_kernel<<<1, 32, 0, st>>>(state);
    }
#define INC_FIN(T, ...) select_finish_kernel<T><<<1, 32, 0, st>>>(state)
    INC_DISPATCH_DTYPE(dtype, INC_FIN, 0)
#undef INC_FIN
}

void launch_radix_select_multi(const SelectSource* srcs, size_t n_srcs, int dtype, const ReduceRecord* d_rec, uint32_t mode,
                               uint64_t arg, uint64_t* state, cudaStream_t st)
{
    size_t esize = 1;
    switch (dtype)
    {
    case INCERTO_U16:
    case INCERTO_I16: esize = 2; break;
    case INCERTO_U32:
    case INCERTO_I32:
    case INCERTO_F32: esize = 4; break;
    case INCERTO_U64:
    case INCERTO_I64:
    case INCERTO_F64: esize = 8; break;
    default: break;
    }
    select_init_kernel<<<1, 256, 0, st>>>(state, d_rec, mode, arg, static_cast<uint32_t>(esize * 8 - 8));
    for (size_t pass = 0; pass < esize; ++pass)
    {
        for (size_t k = 0; k < n_srcs; ++k)
        {
            const SelectSource& q = srcs[k];
            if (!q.n) continue;
            const int blocks = blocks_for(q.n, kReduceThreads, 4);
#define INC_SELM(T, ...)                                                                                  \
    select_hist_kernel<T><<<blocks, kReduceThreads, 0, st>>>(static_cast<const T*>(q.col), q.stride, 0u, q.flags, \
                                                             q.need_set, q.need_clear, q.n, state)
            INC_DISPATCH_DTYPE(dtype, INC_SELM, 0)
#undef INC_SELM
        }
        select_pick_kernel<<<1, 32, 0, st>>>(state);
    }
#define INC_FINM(T, ...) select_finish_kernel<T><<<1, 32, 0, st>>>(state)
    INC_DISPATCH_DTYPE(dtype, INC_FINM, 0)
#undef INC_FINM
}

void launch_add_const(void* col, int dtype, uint32_t stride, const uint8_t* flags, uint8_t need_set, uint64_t n,
                      int64_t amount, cudaStream_t st)
{
    const int blocks = blocks_for(n, 256, 4);
#define INC_ADD(T, ...) \
    add_const_kernel<T><<<blocks, 256, 0, st>>>(static_cast<T*>(col), stride, flags, need_set, n, amount)
    INC_DISPATCH_DTYPE(dtype, INC_ADD, 0)
#undef INC_ADD
}

namespace
{
template <typename D>
void add_column_dst(D* dst, const void* src, int src_dtype, const uint8_t* flags, uint64_t n, cudaStream_t st)
{
    const int blocks = blocks_for(n, 256, 4);
#define INC_ADDC(S, ...) add_column_kernel<D, S><<<blocks, 256, 0, st>>>(dst, static_cast<const S*>(src), flags, n)
    INC_DISPATCH_DTYPE(src_dtype, INC_ADDC, 0)
#undef INC_ADDC
}
} // namespace

void launch_add_column(void* dst, int dst_dtype, const void* src, int src_dtype, const uint8_t* flags, uint64_t n,
                       cudaStream_t st)
{
#define INC_ADDD(T, ...) add_column_dst<T>(static_cast<T*>(dst), src, src_dtype, flags, n, st)
    INC_DISPATCH_DTYPE(dst_dtype, INC_ADDD, 0)
#undef INC_ADDD
}

void launch_find_id(const void* col, int dtype, uint32_t stride, const uint8_t* flags, uint64_t n,
                    uint64_t id_bits, uint64_t* out, cudaStream_t st)
{
    const int blocks = blocks_for(n, 256, 4);
#define INC_FIND(T, ...) \
    find_id_kernel<T><<<blocks, 256, 0, st>>>(static_cast<const T*>(col), stride, flags, n, id_bits, out)
    INC_DISPATCH_DTYPE(dtype, INC_FIND, 0)
#undef INC_FIND
}

void launch_iota_u32(uint32_t* p, uint64_t n, cudaStream_t st)
{
    iota_kernel<<<blocks_for(n, 256, 4), 256, 0, st>>>(p, n);
}

void launch_flush(uint8_t* p, size_t bytes, uint32_t value, cudaStream_t st)
{
    flush_kernel<<<kNumSMs * 8, 256, 0, st>>>(reinterpret_cast<uint4*>(p), bytes / 16, value);
    // the first half of the buffer was written first and has long left the L2: read it back
    flush_read_kernel<<<kNumSMs * 8, 256, 0, st>>>(reinterpret_cast<const uint4*>(p), bytes / 32, reinterpret_cast<uint32_t*>(p));
}

void launch_step_increment(uint32_t* d_step, cudaStream_t st) { step_increment_kernel<<<1, 1, 0, st>>>(d_step); }
void launch_set_u32(uint32_t* p, uint32_t value, cudaStream_t st) { set_u32_kernel<<<1, 1, 0, st>>>(p, value); }

namespace
{
template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ src, T* __restrict__ dst, const uint32_t* __restrict__ perm, uint64_t n)
{
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += gstride) dst[i] = src[perm[i]];
}
__global__ void gather_rows_bytes_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint32_t row_bytes,
                                         const uint32_t* __restrict__ perm, uint64_t n)
{
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += gstride)
        for (uint32_t b = 0; b < row_bytes; ++b) dst[i * row_bytes + b] = src[static_cast<uint64_t>(perm[i]) * row_bytes + b];
}
} // namespace

// dst row i = src row perm[i]; uid_old == nullptr means the identity
void launch_gather_rows(const void* src, void* dst, uint32_t row_bytes, const uint32_t* perm, uint64_t n, cudaStream_t st)
{
    if (!n) return;
    const int blocks = static_cast<int>(std::min<uint64_t>((n + 255) / 256, static_cast<uint64_t>(kNumSMs) * 16));
    switch (row_bytes)
    {
    case 1: gather_rows_kernel<uint8_t><<<blocks, 256, 0, st>>>(static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), perm, n); break;
    case 2: gather_rows_kernel<uint16_t><<<blocks, 256, 0, st>>>(static_cast<const uint16_t*>(src), static_cast<uint16_t*>(dst), perm, n); break;
    case 4: gather_rows_kernel<uint32_t><<<blocks, 256, 0, st>>>(static_cast<const uint32_t*>(src), static_cast<uint32_t*>(dst), perm, n); break;
    case 8: gather_rows_kernel<uint2><<<blocks, 256, 0, st>>>(static_cast<const uint2*>(src), static_cast<uint2*>(dst), perm, n); break;
    case 16: gather_rows_kernel<uint4><<<blocks, 256, 0, st>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), perm, n); break;
    default:
        gather_rows_bytes_kernel<<<blocks, 256, 0, st>>>(static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), row_bytes, perm, n);
    }
}

void launch_compact_offsets(const uint8_t* flags, uint64_t n, uint32_t* d_block_counts, cudaStream_t st)
{
    const uint32_t nblocks = static_cast<uint32_t>((n + kCompactThreads - 1) / kCompactThreads);
    if (nblocks) compact_count_kernel<<<nblocks, kCompactThreads, 0, st>>>(flags, n, d_block_counts);
}
void launch_compact_scan(uint32_t* d_block_counts, uint32_t nblocks, cudaStream_t st)
{
    compact_scan_kernel<<<1, 1024, 0, st>>>(d_block_counts, nblocks);
}
void launch_compact_scatter(const void* src, void* dst, uint32_t row_bytes, const uint8_t* flags, uint64_t n,
                            const uint32_t* d_block_offsets, cudaStream_t st)
{
    const uint32_t nblocks = static_cast<uint32_t>((n + kCompactThreads - 1) / kCompactThreads);
    if (nblocks)
        compact_scatter_kernel<<<nblocks, kCompactThreads, 0, st>>>(static_cast<const uint8_t*>(src),
                                                                    static_cast<uint8_t*>(dst), row_bytes, flags, n,
                                                                    d_block_offsets);
}

} // namespace incerto
