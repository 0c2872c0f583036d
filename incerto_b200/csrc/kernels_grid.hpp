// Device spatial grid (counting sort by cell key). See kernels_grid.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.hpp"

namespace incerto
{

struct GridDesc
{
    uint32_t dim;    // 2 or 3
    int32_t bmin[3]; // inclusive lower bound (GridBounds::min, spatial_grid.rs:45-53)
    int32_t ext[3];  // max - min + 1 per axis (1 for the unused z of a 2-D grid), in units of 2^shift cells
    uint32_t shift = 0; // cells are blocked 2^shift per axis before linearising (row sorts by tile; 0 for real grids)
};

// key = ((x - minx) * ext_y + (y - miny)) * ext_z + (z - minz)
__host__ __device__ inline uint64_t grid_cell_key(const GridDesc& g, int x, int y, int z)
{
    return (static_cast<uint64_t>(x - g.bmin[0]) * g.ext[1] + static_cast<uint64_t>(y - g.bmin[1])) * g.ext[2] +
           static_cast<uint64_t>(z - g.bmin[2]);
}

size_t scan_scratch_bytes(uint64_t n);
void launch_exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* scratch, cudaStream_t st);
// positions: int16 lanes with `stride` elements per row. Rows failing (flags & need_set) == need_set are skipped.
void launch_grid_build(const GridDesc& g, const int16_t* pos, uint32_t stride, const uint8_t* flags,
                       uint8_t need_set, uint64_t n, uint32_t* cell_start, uint32_t* cursor, uint32_t* sorted_rows,
                       uint32_t* scan_scratch, uint32_t* error, cudaStream_t st);

} // namespace incerto
