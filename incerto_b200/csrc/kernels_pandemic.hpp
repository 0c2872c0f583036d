// pandemic kernels (examples/pandemic.rs). See kernels_pandemic.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.hpp"

namespace incerto
{

struct PandemicArgs
{
    uint32_t* pos;        // x | y << 16 per row
    uint8_t* flags;       // bit 0 alive, marker bits
    const uint32_t* uid;  // nullptr: uid == row
    uint64_t n;
    uint32_t grid_size;
    uint32_t step;
    uint32_t bit_person, bit_infected;
    uint32_t do_move, do_infect, do_die, do_recover;
    uint32_t fate_two_level;
    uint64_t thr_infect, thr_die, thr_recover, thr_start_infected;
    PhiloxRoundKeys rk_spawn, rk_move, rk_fate, rk_infect;
    // scratch: per-cell lists of infected people
    uint32_t* cell_bits;     // 1 bit per cell: some infected person stands here this step
    uint32_t* cell_bits_clear; // the next step's bitmap, zeroed by this step's interact kernel (bits_vec4 x 16 B)
    uint64_t bits_vec4;
    uint32_t* cell_head;     // (step & 0xff) << 24 | row of the list head; slot = cell, or hashed (head_shift)
    uint32_t link_mask;      // 0x00ffffff: 24-bit row links + the step stamp in the heads (stale heads read as empty, the
                             // table is cleared every 256 steps); 0xffffffff for >= 2^24 - 1 rows: 32-bit links, no stamp,
                             // the heads are cleared before every step.  link_mask is also the end-of-list marker.
    uint32_t head_shift;     // 0: one slot per cell; else slot = (cell * 0x9E3779B1) >> head_shift
    uint32_t* next_infected; // per row: next infected row in the same cell, 0xffffff = end
};

void launch_pandemic_spawn(const PandemicArgs& a, cudaStream_t st);
void launch_pandemic_move(const PandemicArgs& a, cudaStream_t st);
void launch_pandemic_interact(const PandemicArgs& a, cudaStream_t st);

} // namespace incerto
