// air_traffic_3d kernels (examples/air_traffic_3d.rs). See kernels_air.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.hpp"

namespace incerto
{

struct AirMailbox;

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
    // ---- domain decomposition by x (one world over several GPUs; module_air.cu): this rank owns the aircraft whose
    // grid x lies in [x_lo, x_hi); every rank keeps the full table, "alive" means "owned here"
    uint32_t part;        // 0: the whole airspace lives here
    int32_t x_lo, x_hi;
    uint8_t* flags_rw;    // the same flags, writable: an aircraft that leaves the slab is marked dead here
    AirMailbox* out_lo;   // what the lower / upper neighbour needs after this step (nullptr: no neighbour there)
    AirMailbox* out_hi;
    const AirMailbox* in_lo; // what they sent after the previous step
    const AirMailbox* in_hi;
    const uint32_t* row_of_uid; // nullptr: row == uid
};

// One direction of the per-step exchange with a neighbour rank, a fixed-size blob (the counts travel in the header):
// the aircraft that crossed into the neighbour's slab (uid + new position: the neighbour revives its copy of the row)
// and the positions of the aircraft now standing in this rank's boundary column (the neighbour's ghosts: it needs them
// for the +-x probes of its own boundary aircraft).
constexpr uint32_t kAirMailHeader = 16;
struct AirMailbox
{
    uint32_t n_migrants, n_ghosts, capacity, overflow;
    // followed by capacity x uint4 {uid, pos.x, pos.y, 0} and capacity x uint2 positions
    __host__ __device__ uint4* migrants() { return reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(this) + kAirMailHeader); }
    __host__ __device__ const uint4* migrants() const { return reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(this) + kAirMailHeader); }
    __host__ __device__ uint2* ghosts() { return reinterpret_cast<uint2*>(migrants() + capacity); }
    __host__ __device__ const uint2* ghosts() const { return reinterpret_cast<const uint2*>(migrants() + capacity); }
    static size_t bytes(uint32_t cap) { return kAirMailHeader + static_cast<size_t>(cap) * (16 + 8); }
};

void launch_air_spawn(const AirArgs& a, int16_t* target_out, uint8_t* flags_out, cudaStream_t st);
void launch_air_bits(const AirArgs& a, cudaStream_t st);
void launch_air_candidates(const AirArgs& a, cudaStream_t st);
void launch_air_step(const AirArgs& a, cudaStream_t st);
// domain decomposition
void launch_air_own_init(const AirArgs& a, AirMailbox* in_lo, AirMailbox* in_hi, cudaStream_t st);
void launch_air_ghosts(const AirArgs& a, const AirMailbox* sent_lo, const AirMailbox* sent_hi, cudaStream_t st);
void launch_air_unpack(const AirArgs& a, AirMailbox* in_lo, AirMailbox* in_hi, cudaStream_t st);
void launch_air_mail_reset(AirMailbox* a, AirMailbox* b, AirMailbox* c, AirMailbox* d, cudaStream_t st);
void launch_invert_permutation(const uint32_t* perm, uint32_t* inv, uint64_t n, cudaStream_t st);

} // namespace incerto
