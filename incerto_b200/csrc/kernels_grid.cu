// SpatialGrid<T, C> on the device: src/plugins/spatial_grid.rs:194-311.
// The reference keeps HashMap<GridPosition, HashSet<Entity>> and patches it for
// every Changed<GridPosition>; here the index is rebuilt from scratch by a
// counting sort on the linearised cell key:
//   cell_start[key] .. cell_start[key+1]  ->  rows of the entities in that cell.
// Algorithmic bytes per rebuild: ~16 B/entity (key read 4 + row write 4 +
// histogram/cursor atomics ~8) plus 12 B per cell for the count/scan/reset.
#include "common.cuh"
#include "kernels_grid.hpp"
#include "../../include/incerto_b200.h"

namespace incerto
{
namespace
{

__device__ __forceinline__ bool cell_key(const GridDesc& g, const int16_t* __restrict__ pos, uint32_t stride,
                                         uint64_t row, uint32_t& key)
{
    int x = pos[row * stride + 0] - g.bmin[0];
    int y = pos[row * stride + 1] - g.bmin[1];
    int z = g.dim == 3 ? pos[row * stride + 2] - g.bmin[2] : 0;
    if (x < 0 || y < 0 || z < 0) return false;
    x >>= g.shift;
    y >>= g.shift;
    z >>= g.shift;
    if (x >= g.ext[0] || y >= g.ext[1] || z >= g.ext[2]) return false;
    key = (static_cast<uint32_t>(x) * g.ext[1] + static_cast<uint32_t>(y)) * g.ext[2] + static_cast<uint32_t>(z);
    return true;
}

__global__ void grid_count_kernel(GridDesc g, const int16_t* __restrict__ pos, uint32_t stride,
                                  const uint8_t* __restrict__ flags, uint8_t need_set, uint64_t n,
                                  uint32_t* __restrict__ counts, uint32_t* error)
{
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += gstride)
    {
        if (flags && (flags[i] & need_set) != need_set) continue;
        uint32_t key;
        if (!cell_key(g, pos, stride, i, key))
        {
            // "entity at position {position:?} outside spatial grid bounds" (spatial_grid.rs:276-279)
            if (atomicCAS(error, 0u, static_cast<uint32_t>(INCERTO_ERR_OUT_OF_BOUNDS)) == 0u)
                error[1] = static_cast<uint32_t>(i);
            continue;
        }
        atomicAdd(&counts[key], 1u);
    }
}

__global__ void grid_scatter_kernel(GridDesc g, const int16_t* __restrict__ pos, uint32_t stride,
                                    const uint8_t* __restrict__ flags, uint8_t need_set, uint64_t n,
                                    const uint32_t* __restrict__ cell_start, uint32_t* __restrict__ cursor,
                                    uint32_t* __restrict__ sorted_rows)
{
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += gstride)
    {
        if (flags && (flags[i] & need_set) != need_set) continue;
        uint32_t key;
        if (!cell_key(g, pos, stride, i, key)) continue;
        const uint32_t slot = cell_start[key] + atomicAdd(&cursor[key], 1u);
        sorted_rows[slot] = static_cast<uint32_t>(i);
    }
}

// ---- exclusive scan of u32 (three phases; 2048 items per block)
constexpr int kScanThreads = 512;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const uint32_t* __restrict__ in, uint64_t n,
                                                                   uint32_t* block_sums)
{
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kScanTile;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
    {
        const uint64_t i = base + static_cast<uint64_t>(k) * kScanThreads + threadIdx.x;
        if (i < n) s += in[i];
    }
    uint32_t v[1] = {s};
    __shared__ uint32_t smem[32];
    block_sum<uint32_t, 1>(v, smem);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = v[0];
}

__global__ void __launch_bounds__(1024) scan_block_sums_kernel(uint32_t* block_sums, uint32_t nblocks)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nblocks; base += 1024)
    {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nblocks ? block_sums[i] : 0;
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
        if (i < nblocks) block_sums[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[nblocks] = s_carry;
}

__global__ void __launch_bounds__(kScanThreads)
    scan_apply_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint64_t n,
                      const uint32_t* __restrict__ block_sums, uint32_t nblocks)
{
    // thread t owns items [t*kScanItems, (t+1)*kScanItems) of the tile: contiguous -> 128-bit accesses
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kScanTile + static_cast<uint64_t>(threadIdx.x) * kScanItems;
    uint32_t v[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) v[k] = (base + k < n) ? in[base + k] : 0;
    uint32_t tsum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) tsum += v[k];
    // exclusive scan of tsum over the block
    uint32_t x = tsum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t y = __shfl_up_sync(kFullMask, x, o);
        if (lane >= o) x += y;
    }
    __shared__ uint32_t s_warp[kScanThreads / 32];
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0)
    {
        uint32_t w = lane < kScanThreads / 32 ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t y = __shfl_up_sync(kFullMask, w, o);
            if (lane >= o) w += y;
        }
        if (lane < kScanThreads / 32) s_warp[lane] = w;
    }
    __syncthreads();
    uint32_t run = block_sums[blockIdx.x] + (warp ? s_warp[warp - 1] : 0) + (x - tsum);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
    {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (blockIdx.x == nblocks - 1 && threadIdx.x == 0) out[n] = block_sums[nblocks]; // total
}

} // namespace

size_t scan_scratch_bytes(uint64_t n) { return ((n + kScanTile - 1) / kScanTile + 2) * sizeof(uint32_t); }

void launch_exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* scratch, cudaStream_t st)
{
    if (n == 0) return;
    const uint32_t nblocks = static_cast<uint32_t>((n + kScanTile - 1) / kScanTile);
    scan_reduce_kernel<<<nblocks, kScanThreads, 0, st>>>(in, n, scratch);
    scan_block_sums_kernel<<<1, 1024, 0, st>>>(scratch, nblocks);
    scan_apply_kernel<<<nblocks, kScanThreads, 0, st>>>(in, out, n, scratch, nblocks);
}

void launch_grid_build(const GridDesc& g, const int16_t* pos, uint32_t stride, const uint8_t* flags,
                       uint8_t need_set, uint64_t n, uint32_t* cell_start, uint32_t* cursor, uint32_t* sorted_rows,
                       uint32_t* scan_scratch, uint32_t* error, cudaStream_t st)
{
    const uint64_t n_cells = static_cast<uint64_t>(g.ext[0]) * g.ext[1] * g.ext[2];
    cudaMemsetAsync(cursor, 0, n_cells * sizeof(uint32_t), st);
    uint64_t b = (n + 255) / 256;
    if (b < 1) b = 1;
    if (b > static_cast<uint64_t>(kNumSMs) * 16) b = kNumSMs * 16;
    grid_count_kernel<<<static_cast<int>(b), 256, 0, st>>>(g, pos, stride, flags, need_set, n, cursor, error);
    launch_exclusive_scan_u32(cursor, cell_start, n_cells, scan_scratch, st);
    cudaMemsetAsync(cursor, 0, n_cells * sizeof(uint32_t), st);
    grid_scatter_kernel<<<static_cast<int>(b), 256, 0, st>>>(g, pos, stride, flags, need_set, n, cell_start, cursor,
                                                             sorted_rows);
}

} // namespace incerto
