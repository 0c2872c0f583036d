// Launch wrappers of the sm_100a kernels (definitions in kernels_*.cu).
// Everything here takes raw device pointers and a stream; no engine types.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "philox.cuh"

namespace incerto
{

constexpr int kNumSMs = 148;

// ------------------------------------------------------------------ generic
struct ReduceRecord
{
    uint64_t count;
    uint64_t sum_bits; // i64 / u64 (wrapping) or f64 bits
    uint64_t min_bits; // value widened to 64 bit (i64 / u64 / f64 bits)
    uint64_t max_bits;
    uint32_t nan_seen;
    uint32_t pad;
};

// count / sum / min / max of lane `lane` of a strided column under a row filter.
// out: one ReduceRecord in device memory. scratch: >= max_blocks * sizeof(ReduceRecord).
void launch_reduce_column(const void* col, int dtype, uint32_t stride, uint32_t lane, const uint8_t* flags,
                          uint8_t need_set, uint8_t need_clear, uint64_t n, ReduceRecord* out, void* scratch,
                          uint32_t* ticket, cudaStream_t st);
// rows matching the filter; out u64.
void launch_count_rows(const uint8_t* flags, uint8_t need_set, uint8_t need_clear, uint64_t n, uint64_t* out,
                       void* scratch, uint32_t* ticket, cudaStream_t st);

// k-th smallest (0-based) of the filtered column, by 8-bit MSD radix select on
// order-preserving keys. state: 4 + 256 u64 of device scratch. The value (widened
// to 64 bit like ReduceRecord::min_bits) lands in state[0], an index-out-of-range
// flag in state[3].  k comes from `arg` (mode 0) or is derived on the device from
// d_rec->count: mode 1 = Median (count/2), mode 2 = Percentile<arg>.
// `between_passes` (optional) runs after every histogram pass, before the digit is picked: the hook of the distributed
// select, which sums the 256 bins at state + 4 over all ranks (d_rec->count is then the global count).
void launch_radix_select(const void* col, int dtype, uint32_t stride, uint32_t lane, const uint8_t* flags,
                         uint8_t need_set, uint8_t need_clear, uint64_t n, const ReduceRecord* d_rec, uint32_t mode,
                         uint64_t arg, uint64_t* state, cudaStream_t st, int32_t (*between_passes)(void*) = nullptr,
                         void* hook_ctx = nullptr);
// the same over several columns of one dtype (entities of several archetype tables): every pass accumulates the
// digit histogram of all sources before the digit is picked; `d_rec` holds the combined count
struct SelectSource
{
    const void* col;
    uint32_t stride;
    const uint8_t* flags;
    uint8_t need_set, need_clear;
    uint64_t n;
};
void launch_radix_select_multi(const SelectSource* srcs, size_t n_srcs, int dtype, const ReduceRecord* d_rec, uint32_t mode,
                               uint64_t arg, uint64_t* state, cudaStream_t st);
constexpr size_t kSelectStateWords = 4 + 256;

void launch_add_const(void* col, int dtype, uint32_t stride, const uint8_t* flags, uint8_t need_set, uint64_t n,
                      int64_t amount, cudaStream_t st);
void launch_add_column(void* dst, int dst_dtype, const void* src, int src_dtype, const uint8_t* flags, uint64_t n,
                       cudaStream_t st);
// rows whose identifier equals id_bits: out[0] = matches, out[1] = first matching row.
void launch_find_id(const void* col, int dtype, uint32_t stride, const uint8_t* flags, uint64_t n,
                    uint64_t id_bits, uint64_t* out, cudaStream_t st);
void launch_iota_u32(uint32_t* p, uint64_t n, cudaStream_t st);
void launch_flush(uint8_t* p, size_t bytes, uint32_t value, cudaStream_t st);
// step counter kept on the device so that captured step programs need no new arguments
void launch_step_increment(uint32_t* d_step, cudaStream_t st);
void launch_set_u32(uint32_t* p, uint32_t value, cudaStream_t st);

// dst row i = src row perm[i] (row reordering of a column)
void launch_gather_rows(const void* src, void* dst, uint32_t row_bytes, const uint32_t* perm, uint64_t n, cudaStream_t st);

// stream compaction of a table by its alive bit (ballot + prefix sum)
// d_block_counts: >= ceil(n/1024)+1 u32; returns nothing; new_n is read back from d_block_counts[nblocks].
void launch_compact_offsets(const uint8_t* flags, uint64_t n, uint32_t* d_block_counts, cudaStream_t st);
void launch_compact_scan(uint32_t* d_block_counts, uint32_t nblocks, cudaStream_t st);
void launch_compact_scatter(const void* src, void* dst, uint32_t row_bytes, const uint8_t* flags, uint64_t n,
                            const uint32_t* d_block_offsets, cudaStream_t st);

// ---------------------------------------------------------------- coin toss
// examples/coin_toss.rs:67-81. heads/tosses are u32 columns; `steps` >= 1 steps
// are applied back to back in registers (entities are independent).
void launch_coin_toss(uint32_t* heads, uint32_t* tosses, const uint32_t* uid, const uint8_t* flags, uint64_t n,
                      PhiloxKey key, uint64_t threshold, uint32_t first_step, uint32_t steps, cudaStream_t st);
void launch_coin_toss_u64(uint64_t* heads, uint64_t* tosses, const uint32_t* uid, const uint8_t* flags, uint64_t n,
                          PhiloxKey key, uint64_t threshold, uint32_t first_step, uint32_t steps, cudaStream_t st);
// examples/coin_toss.rs:33-49: out[0] = f64 sum of heads/tosses, out[1] = count (as u64 bits)
void launch_coin_odds(const void* heads, const void* tosses, int dtype, const uint8_t* flags, uint8_t need_set,
                      uint8_t need_clear, uint64_t n, double* out, void* scratch, uint32_t* ticket,
                      cudaStream_t st);

// -------------------------------------------------------------- forest fire
struct ForestArgs
{
    const uint8_t* cur; // W*H cells, x-major (uid = x*H + y), rows [x_begin-?]: see kernels_forest.cu
    uint8_t* next;
    int32_t W, H;           // global grid
    int32_t x_begin, x_end; // slab owned by this process
    const uint8_t* halo_lo; // H cells of column x_begin-1 (nullptr: out of bounds)
    const uint8_t* halo_hi; // H cells of column x_end
    // overlay entities (the extra igniters of forest_fire.rs:260-278)
    uint32_t n_overlay;
    const int16_t* ov_pos;   // n_overlay * 2 (x, y), device
    const uint32_t* near_bits;  // device bitmap, one bit per CTA tile with an overlay entity within one cell
    const uint8_t* ov_cur;
    uint8_t* ov_next;
    uint32_t ov_uid_base;    // uid of overlay entity 0 (= W*H)
    uint32_t ov_owned;       // overlay entities whose position lies in this slab
    // enabled systems and parameters
    uint32_t do_spread, do_progress, do_regrow;
    uint64_t thr_spread;
    uint64_t thr_regrow;        // floor(p * 2^32)
    uint64_t thr_regrow_l2;     // floor(256 * p * 2^32), used when regrow_two_level
    uint32_t regrow_two_level;  // p * 256 <= 1
    uint32_t burn_duration;
    PhiloxKey key_spread, key_regrow;
    uint32_t l2_resident;       // hint: consecutive steps of a grid whose two buffers stay in the L2 (graph replay)
    // StepNumber of the launch = (d_step ? *d_step : 0) + step_base: by value for single launches, device counter +
    // offset inside a captured graph of steps (which then replays without new arguments)
    const uint32_t* d_step;
    uint32_t step_base;
    // fused FireStats (forest_fire.rs:104-134): per-step record + series ring
    uint64_t* stats_run;      // 12 u64: totals (healthy, burning, empty), last published step, two delta slots of
                              // this slab + owned overlays (layout: kernels_forest.cu, forest_publish)
    uint32_t* error;          // device error flag
    uint32_t* ticket;         // (unused by the step kernel since the stamps left it; kept for layout)
    uint64_t* stats_out;      // 5 u64: healthy, burning, burned, empty, total (state after this step)
    uint8_t* series_values;   // nullptr or ring of incerto_fire_stats
    uint64_t* series_counts;
    uint64_t series_interval, series_capacity;
};
// Halo transport fused into the step kernel (row partition, neighbours' slabs mapped through CUDA IPC).  A stamp k in
// a rank's mailbox says "I finished step k: the buffer that step wrote is complete".  The warps that read a halo column
// wait for the neighbour's stamp of the previous step (read over NVLink) and then PULL its boundary column.  The writer's
// side is one plain store into its OWN mailbox by the first thread of the launch that FOLLOWS step k (the next step's
// grid, or the publish kernel at the end of a run): the kernel boundary has made step k's cells visible; a step never
// writes a neighbour's memory (a grid that does completes ~5 us later, kernels_forest.cu: forest_stamp).
// The CTAs of the two boundary column groups are scheduled first, so that the wait and the NVLink round trip happen
// while the rest of the grid works.  The neighbour cannot overwrite what is being pulled: the warps that write its
// boundary column in step k+2 wait for MY stamp k+1, which leaves after my step k+1 (and its pulls) completed.
struct ForestPeer
{
    uint32_t* my_mail = nullptr;         // my mailbox (slot 4: the last step of mine that is complete; slot 5: rendezvous epoch)
    uint32_t need_lo = 0, need_hi = 0;   // stamp the neighbour's buffer must carry (0: first step, halos spawned locally)
    const uint8_t* src_lo = nullptr;     // lower neighbour's last own column in the buffer this step reads
    const uint8_t* src_hi = nullptr;     // upper neighbour's first own column
    const uint32_t* mail_lo = nullptr;   // the lower / upper neighbour's mailbox, read over NVLink
    const uint32_t* mail_hi = nullptr;
    uint32_t stamp = 0;                  // step to announce at the start of this launch (0: none yet)
    uint32_t debug = 0;                  // timing experiments (INCERTO_B200_DEBUG_HALO): 1 no boundary-first order, 2 no pulls
};
cudaError_t launch_forest_step(const ForestArgs& a, cudaStream_t st);
// put step (d_step ? *d_step : 0) + step_base on record (idempotent) and add `bump` to the device StepNumber
cudaError_t launch_forest_publish(const ForestArgs& a, uint32_t bump, cudaStream_t st, const ForestPeer* peer = nullptr);
// pdl: launch as a programmatic dependent of the previous launch on the stream (captured runs of steps on one GPU)
cudaError_t launch_forest_step_ex(const ForestArgs& a, const uint8_t* ov_lo, const uint8_t* ov_hi, const uint8_t* ov_owner,
                                  cudaStream_t st, const ForestPeer* peer = nullptr, bool pdl = false);
int forest_step_grid_blocks(int32_t slab_w, int32_t H);
uint32_t forest_pitch(int32_t H);            // bytes per column: H rounded up to whole CTA tiles
uint32_t forest_alloc_cols(int32_t slab_w);  // columns allocated: slab_w rounded up to whole CTA tiles + 2 halos
void forest_tile_shape(uint32_t* cols, uint32_t* rows);
// halo transport over peer memory: store my two boundary columns (+ overlay states) into the neighbours' halo columns
// and then publish `stamp` in their mailboxes; and its counterpart, wait until both neighbours published `need`
void launch_forest_halo_push(const uint8_t* col_lo, uint8_t* peer_lo, const uint8_t* col_hi, uint8_t* peer_hi,
                             uint32_t col_bytes, const uint8_t* ov, uint8_t* peer_ov_lo, uint8_t* peer_ov_hi, uint32_t n_ov,
                             uint32_t* peer_mail_lo, uint32_t* peer_mail_hi, uint32_t stamp, cudaStream_t st);
void launch_forest_halo_wait(const uint32_t* mail, uint32_t need_lo, uint32_t need_hi, uint32_t* error, cudaStream_t st);
// neighbour rendezvous over the peer mailboxes (slot 5), epoch strictly increasing
void launch_forest_rendezvous(uint32_t* mail, const uint32_t* peer_mail_lo, const uint32_t* peer_mail_hi, uint32_t epoch, uint32_t* error, cudaStream_t st);
void launch_forest_stamp_wait(const uint32_t* peer_mail_lo, const uint32_t* peer_mail_hi, uint32_t need, uint32_t* error, cudaStream_t st);
void launch_forest_spawn(uint8_t* cells, int32_t W, int32_t H, int32_t x_begin, int32_t x_end, PhiloxKey key,
                         uint64_t thr_density, cudaStream_t st);
// FireStats of the owned columns of a pitched slab (standalone sample_aggregate)
void launch_forest_stats(const uint8_t* slab, int32_t H, int32_t slab_w, uint64_t* out5, void* scratch,
                         uint32_t* ticket, cudaStream_t st);

} // namespace incerto
