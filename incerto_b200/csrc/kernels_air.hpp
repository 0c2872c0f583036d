// air_traffic_3d kernels (examples/air_traffic_3d.rs). See kernels_air.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.hpp"

namespace incerto
{

struct AirArgs
{
    uint2* pos;            // GridPosition3D as int16 x, y, z, pad per row (one 8-byte word)
    const int16_t* target; // Aircraft.target_altitude
    const uint8_t* flags;  // nullptr: every row alive
    const uint32_t* uid;   // nullptr: uid == row
    uint64_t n;
    int32_t bmin[3], ext[3]; // spatial grid bounds (GridBounds3D, inclusive): min and max - min + 1
    int32_t size, height;    // AIRSPACE_SIZE / AIRSPACE_HEIGHT
    uint32_t step;
    uint32_t do_check, do_move;
    uint64_t thr_move;
    PhiloxRoundKeys rk_move, rk_spawn;
    // occupancy index of the step-start snapshot: open addressing, slot = tag:8 | cell key:32 | count:24
    unsigned long long* slots;
    uint32_t slot_mask; // slots - 1 (power of two)
    uint32_t slot_shift; // 32 - log2(slots)
    uint32_t tag;       // 1..255, changes every step; slots with another tag are empty
    uint64_t* acc;      // [0] conflicts of this step (each pair counted twice), [1] aircraft indexed, [2] row + 1 of an aircraft that left the bounds
    // occupancy pre-filter: one bit per hashed cell (no false negatives); only aircraft that see a set bit
    // in one of their six face-neighbour cells enter the exact index above
    uint32_t* bits;
    uint32_t* bits_next; // bitmap of the next step, filled by the step kernel from the moved positions (nullptr: not wanted)
    uint32_t bits_shift; // 32 - log2(number of bits)
    uint8_t* cand;       // per row: which of the six neighbour cells (-x +x -y +y -z +z) showed a set bit
    uint32_t* error;
};

void launch_air_spawn(const AirArgs& a, int16_t* target_out, uint8_t* flags_out, cudaStream_t st);
void launch_air_bits(const AirArgs& a, cudaStream_t st);
void launch_air_candidates(const AirArgs& a, cudaStream_t st);
void launch_air_step(const AirArgs& a, cudaStream_t st);

} // namespace incerto
