// pandemic_spatial kernels (examples/pandemic_spatial.rs). See kernels_spatial.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.hpp"

namespace incerto
{

constexpr uint32_t kSpatialHistorySlots = 20; // ContactHistory capacity (cleared when it would exceed 20, :555)
constexpr uint32_t kSpatialRowBits = 26;      // per-cell source lists link rows with 26 bits + 6-bit step tag

struct SpatialArgs
{
    uint32_t* pos;     // GridPosition2D: x | y << 16
    uint8_t* disease;  // Person.disease_state (INCERTO_DISEASE_*)
    uint8_t* flags;    // bit 0 alive, marker bits (Person.social_distancing)
    uint8_t* quar;     // Quarantined.remaining_duration, 0 = component absent
    uint8_t* aux;      // module-owned: ContactHistory length (bits 0-4) | pending op (bits 5-7)
    uint32_t* hist;    // ContactHistory: slot k of row i at [k * n_cap + i], packed like pos
    const uint32_t* uid;
    uint64_t n, n_cap;
    int32_t G;
    uint32_t step, systems; // bit k: system kind 40 + k
    uint32_t bit_social;
    uint32_t radius, incubation, crowding, quarantine_duration, history_cap;
    int32_t zone_x, zone_y, zone_radius;
    uint64_t thr_infect1, thr_infect2, thr_recover, thr_die, thr_attempt, thr_social, thr_start;
    PhiloxRoundKeys rk_spawn, rk_move, rk_transmit, rk_fate;
    uint32_t* count;    // persons per cell (all live persons), as of the last PreUpdate
    uint32_t* src_bits; // 1 bit per cell: a non-quarantined infectious person stands here this step
    uint32_t* src_head; // tag:6 | row:26 of the list head
    uint32_t* src_next; // per row: next source in the same cell
    uint32_t tag;       // 1..63
    unsigned long long* mark; // per cell: markers (non-quarantined people remembering it) << 32 | sum of their uid + 1
    uint8_t* marked;          // 1 bit per row: this person's remembered cells are currently in `mark`
    uint64_t* stats_acc; // healthy, exposed, infectious, recovered of the post-flush state (zeroed by prepare, summed by interact)
    uint32_t* error;
};

void launch_spatial_spawn(const SpatialArgs& a, cudaStream_t st);
void launch_spatial_prepare(const SpatialArgs& a, cudaStream_t st);
void launch_spatial_interact(const SpatialArgs& a, cudaStream_t st);
void launch_spatial_stats_finalize(const uint64_t* acc, uint64_t initial_population, uint64_t* out6, cudaStream_t st);
// PandemicStats (examples/pandemic_spatial.rs:107-138): out6 = healthy, exposed, infectious, recovered, dead, total
void launch_spatial_stats(const uint8_t* disease, const uint8_t* flags, uint64_t n, uint64_t initial_population,
                          uint64_t* out6, void* scratch, uint32_t* ticket, cudaStream_t st);

} // namespace incerto
