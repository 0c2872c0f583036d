// Device helpers shared by the kernels: vector loads, warp/block reductions,
// the "last block combines" pattern used for every aggregate.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstdint>
#include <utility>

namespace incerto
{

constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ uint4 ld_stream_v4(const void* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_v4(void* p, const uint4& v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

// Programmatic dependent launch (see launch_pdl): let the next kernel of the stream be scheduled while this one drains,
// and wait - before the first access to anything the previous kernel wrote - until that kernel has completed.  Both
// are no-ops for a kernel launched the plain way.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}

// Block-wide sum of up to N values per thread; result valid in thread 0.
// smem must hold N * 32 elements of T.
template <typename T, int N>
__device__ __forceinline__ void block_sum(T (&v)[N], T* smem)
{
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int nwarps = (blockDim.x * blockDim.y + 31) >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0)
    {
#pragma unroll
        for (int i = 0; i < N; ++i) smem[i * 32 + warp] = v[i];
    }
    __syncthreads();
    if (warp == 0)
    {
#pragma unroll
        for (int i = 0; i < N; ++i)
        {
            T x = lane < nwarps ? smem[i * 32 + lane] : T(0);
            v[i] = warp_sum(x);
        }
    }
}

// Returns true in every thread of exactly one block: the last one to arrive.
// The caller must have written its partials and will read all partials after.
__device__ __forceinline__ bool last_block_arrives(uint32_t* ticket, uint32_t total_blocks)
{
    __shared__ uint32_t s_last;
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0)
    {
        // the barrier orders the block's writes before this thread; its fence is cumulative over them
        __threadfence();
        const uint32_t t = atomicAdd(ticket, 1u);
        s_last = (t == total_blocks - 1) ? 1u : 0u;
        if (s_last) *ticket = 0u; // leave the ticket ready for the next launch
    }
    __syncthreads();
    const bool last = s_last != 0;
    if (last) __threadfence();
    return last;
}

// Blocks of `waves` resident waves of `kernel` (occupancy x SMs): the grid of a grid-stride loop whose iterations cost
// about the same - every warp gets an equal share, where a fixed cap of k blocks per SM leaves a last wave that is partly
// empty.  How many waves is measured per kernel (one for the streaming kernels; two where a second wave hides the first
// one's tail).  INCERTO_B200_WAVES=<w> (experiments) overrides it everywhere.
template <typename K>
inline int wave_blocks(K kernel, int threads, double waves = 1.0, size_t smem = 0)
{
    int per_sm = 0, dev = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
    static const double forced = [] { const char* e = std::getenv("INCERTO_B200_WAVES"); return e ? std::atof(e) : 0.0; }();
    if (forced > 0) waves = forced;
    return std::max(1, static_cast<int>(per_sm * sms * waves));
}

// Launch `kernel` as a programmatic dependent of the previous launch on `st`: its CTAs may start as SM slots free up;
// the kernel must call pdl_wait() before it touches the previous kernel's results.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool plain = [] { const char* e = std::getenv("INCERTO_B200_NO_PDL"); return e && e[0] == '1'; }();
    cfg.attrs = at;
    cfg.numAttrs = plain ? 0 : 1; // INCERTO_B200_NO_PDL=1: ordinary stream order (A/B measurements)
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

} // namespace incerto
