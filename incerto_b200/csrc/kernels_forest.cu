// forest_fire: examples/forest_fire.rs
//   fire_spread_system      :282-329
//   burn_progression_system :332-351
//   regrowth_system         :354-365
//   FireStats aggregate     :104-134 (fused: counted on the state this step produced)
//
// One launch advances every cell one step through all three systems in the
// pinned order spread -> progression -> regrowth, reading the
// previous state buffer and writing the next one (double buffer), so the only
// HBM traffic is 1 byte read + 1 byte written per cell.
//
// Layout: the slab owned by this process holds grid columns x in
// [x_begin, x_end) ("column" = all cells with the same x, contiguous in y, as
// the reference spawns them: uid = x*H + y, forest_fire.rs:239-257) plus one
// halo column on either side:  col 0 = x_begin-1, cols 1..slab_w, col slab_w+1 = x_end.
// Each column is `pitch` = round_up(H, 512) bytes (whole warp spans; columns are padded to whole CTA tiles too); pad bytes and out-of-grid halo
// columns hold BURNED (0), which is inert under all three systems.
//
// Cell encoding (u8): BURNED 0, BURNING{t} = t in 1..63, HEALTHY 0x40, EMPTY 0x80.
//
// Draw contract:
//   spread, healthy target T, burning source S on the base grid at direction
//     d (N(0,-1)=0, W(-1,0)=1, E(+1,0)=2, S(0,+1)=3; spatial_grid.rs:65-69):
//     word d of philox(key[SPREAD], (uid_T, 0, step, 0)) < thr_spread
//   spread, source S an overlay entity (uid >= W*H):
//     word 0 of philox(key[SPREAD], (uid_T, uid_S, step, 1)) < thr_spread
//   regrow, empty entity uid, p*256 <= 1 (two-level, exact product of probabilities):
//     uid is a winner of the block lottery of block uid >> 8 at this step (lottery.cuh: 256 independent
//     Bernoulli(2^-8) draws realised by ONE call philox(key[REGROW], (uid >> 8, 0, step, 0)))
//     AND word (uid & 3) of philox(key[REGROW], (uid >> 2, 0, step, 1)) < floor(256*p*2^32)
//   regrow, p*256 > 1: word (uid & 3) of philox(key[REGROW], (uid >> 2, 0, step, 1)) < floor(p*2^32)
#include <cstdlib>

#include "common.cuh"
#include "kernels.hpp"
#include "lottery.cuh"

namespace incerto
{
namespace
{

constexpr int kTX = 32;  // chunks (16 cells) along y per CTA
#ifndef INCERTO_FOREST_TY
#define INCERTO_FOREST_TY 4
#endif
constexpr int kTY = INCERTO_FOREST_TY;   // column groups (warps) per CTA
#ifndef INCERTO_FOREST_COLS
#define INCERTO_FOREST_COLS 7
#endif
constexpr int kColsPerThread = INCERTO_FOREST_COLS; // adjacent columns per thread
#ifndef INCERTO_FOREST_MIN_BLOCKS
#define INCERTO_FOREST_MIN_BLOCKS 8
#endif
constexpr int kMinBlocks = INCERTO_FOREST_MIN_BLOCKS;

__device__ __forceinline__ uint32_t burn_bits(uint32_t w)
{
    // 0x01 in every byte whose low six bits are non-zero (BURNING)
    return (((w & 0x3f3f3f3fu) + 0x3f3f3f3fu) >> 6) & 0x01010101u;
}
__device__ __forceinline__ uint32_t zero_bytes(uint32_t v)
{
    // 0x80 in every byte of v that is zero
    return ~(((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v) & 0x80808080u;
}

#ifdef INCERTO_FOREST_TRACE
// developer instrumentation (tools/trace_forest.py): per warp {start, loads consumed, end} in globaltimer ns + SM id
__device__ unsigned long long g_forest_trace[8 * (1 << 16)];
__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t smid()
{
    uint32_t v;
    asm volatile("mov.u32 %0, %smid;" : "=r"(v));
    return v;
}
#endif

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Fused halo transport, stamp side.  "Step k of this rank is complete" is announced by the launch that FOLLOWS step k
// (the first thread of step k+1's grid, or the publish kernel at the end of a run): the kernel boundary has already made
// every cell and overlay state of step k visible, so the stamp is one plain store - into this rank's OWN mailbox (slot
// kStampSlot), where the neighbours read it over NVLink.  No remote write, no fence, no ticket inside a step: measured
// by switching the halves of the transport off separately, ONE 4-byte store into a neighbour's memory - wherever in the
// kernel, whatever its ordering - made every step of a row-split grid 5 us longer (28.0 against 23.2 us; a grid that
// wrote peer memory completes later), as did the round-1 design (device-scope fence + ticket in each boundary CTA,
// system-scope fence + remote stamp by the last of them).  Remote READS cost nothing of the kind.
constexpr uint32_t kStampSlot = 4, kEpochSlot = 5;
__device__ __forceinline__ void forest_stamp(const ForestPeer& pr)
{
    if (pr.stamp && pr.my_mail) *reinterpret_cast<volatile uint32_t*>(pr.my_mail + kStampSlot) = pr.stamp;
}

constexpr int kEventQueue = 512;   // rare events of one warp tile handled per round
constexpr uint16_t kNullEvent = 0xffffu;
static_assert(INCERTO_FOREST_COLS <= 7, "a lottery event carries its column in three bits, and 0xffff must stay free");
constexpr int kMaxNear = 64;       // overlay entities within one cell of a CTA tile

struct ForestConst
{
    ForestArgs a;
    uint32_t pitch, chunks_per_col, slab_w;
    uint32_t regrow_two_level;
    uint64_t thr_regrow_l2;
    const uint8_t* ov_lo;    // overlay states received from the lower / upper neighbour rank
    const uint8_t* ov_hi;
    const uint8_t* ov_owner; // 0 = this rank, 1 = lower neighbour, 2 = upper neighbour, 3 = elsewhere
    const uint32_t* near_bits; // one bit per CTA tile: an overlay entity lies within one cell of the tile
    uint32_t* error;
    PhiloxRoundKeys rk_spread, rk_regrow; // key schedules of the two streams, as constant operands
    ForestPeer peer;         // fused halo transport (all null without it)
};

__device__ __forceinline__ uint32_t ignited_value(const ForestArgs& a)
{
    // Burning{burn_duration} (forest_fire.rs:324-326) then, when progression is
    // enabled, one progression step in the same update (:340-349)
    if (!a.do_progress) return a.burn_duration;
    return a.burn_duration > 1 ? a.burn_duration - 1 : 0;
}

__device__ __forceinline__ bool regrow_draw(const ForestConst& c, uint32_t uid, uint32_t step)
{
    if (c.regrow_two_level && !lottery256_wins(c.rk_regrow, uid, 0u, step)) return false;
    const uint4 l2 = philox4x32_10(uid >> 2, 0u, step, 1u, c.rk_regrow);
    return bernoulli(word_of(l2, uid & 3u), c.thr_regrow_l2);
}

// FireStats category of a state code as a delta vector (healthy, burning, empty); burned is implicit
struct Delta
{
    int32_t h = 0, b = 0, e = 0;
    __device__ __forceinline__ void add(uint32_t s, int32_t sign)
    {
        h += (s & 0x40u) ? sign : 0;
        e += (s & 0x80u) ? sign : 0;
        b += (s & 0x3fu) ? sign : 0;
    }
};

// Overlay entities near this CTA's tile, staged in shared memory.  Overlay entities are the igniters the
// reference spawns on top of existing cells (forest_fire.rs:260-278) — or any other entity sharing a
// cell of the dense grid.  The vector path ignores them; the thread that owns a cell an overlay can
// influence re-evaluates that cell with `marked`, a scalar restatement of fire_spread_system for one
// position (:296-316).
struct NearOverlays
{
    int16_t x[kMaxNear], y[kMaxNear];
    uint16_t k[kMaxNear];
    uint8_t old[kMaxNear], owner[kMaxNear];
    uint32_t n;
};

__device__ __forceinline__ bool overlay_marked(const ForestConst& c, const NearOverlays& nr, uint32_t step, int px, int py)
{
    const ForestArgs& a = c.a;
    if (!a.do_spread) return false;
    const uint8_t* pc = a.cur + static_cast<size_t>(px - a.x_begin + 1) * c.pitch + py;
    // old states of the base cell and of its four orthogonal neighbours (halo columns hold the neighbour rank's cells)
    const uint32_t s_c = pc[0];
    const uint32_t s_n = py > 0 ? pc[-1] : 0u;
    const uint32_t s_w = px > 0 ? pc[-static_cast<ptrdiff_t>(c.pitch)] : 0u;
    const uint32_t s_e = px + 1 < a.W ? pc[c.pitch] : 0u;
    const uint32_t s_s = py + 1 < a.H ? pc[1] : 0u;
    const uint32_t base_dirs = ((s_n & 0x3f) ? 1u : 0u) | ((s_w & 0x3f) ? 2u : 0u) | ((s_e & 0x3f) ? 4u : 0u) | ((s_s & 0x3f) ? 8u : 0u);
    const uint32_t n = nr.n < kMaxNear ? nr.n : kMaxNear;
    // healthy entities at p: the base cell (e == -1) and overlays
    for (int e = -1; e < static_cast<int>(n); ++e)
    {
        uint32_t uid_t;
        if (e < 0)
        {
            if (s_c != 0x40u) continue;
            uid_t = static_cast<uint32_t>(px) * a.H + py;
        }
        else
        {
            if (nr.x[e] != px || nr.y[e] != py || nr.old[e] != 0x40) continue;
            uid_t = a.ov_uid_base + nr.k[e];
        }
        if (base_dirs)
        {
            const uint4 r = philox4x32_10(uid_t, 0u, step, 0u, c.rk_spread);
            if (((base_dirs & 1u) && bernoulli(r.x, a.thr_spread)) || ((base_dirs & 2u) && bernoulli(r.y, a.thr_spread)) ||
                ((base_dirs & 4u) && bernoulli(r.z, a.thr_spread)) || ((base_dirs & 8u) && bernoulli(r.w, a.thr_spread)))
                return true;
        }
        for (uint32_t o = 0; o < n; ++o)
        {
            if (!(nr.old[o] & 0x3f)) continue;
            const int dx = nr.x[o] - px, dy = nr.y[o] - py;
            if (dx * dx + dy * dy != 1) continue; // burning overlay on an orthogonally adjacent position
            const uint4 rp = philox4x32_10(uid_t, a.ov_uid_base + nr.k[o], step, 1u, c.rk_spread);
            if (bernoulli(rp.x, a.thr_spread)) return true;
        }
    }
    return false;
}

// The vectorised body: each thread owns kColsPerThread (7) adjacent columns x 16 cells (nine 128-bit loads incl. the
// W/E neighbour columns), a CTA of 4 warps 28 columns x 512 cells, 8 CTAs per SM at 64 registers: a 4096^2 grid is ONE
// resident wave (1176 CTAs on 1184 slots).  Width and registers have to be tuned together (profiles/r01_forest_cols_sweep.txt,
// profiles/r02_forest_lottery_sweeps.txt).
//  * streaming path, fully unrolled and register resident: 128-bit loads of the chunk and its W/E neighbours, one
//    warp-wide OR that says which of the warp's columns have fire in or next to them (progression / spread work is done
//    for those only; by far most warps have none and copy), the 128-bit store;
//  * rare events - healthy cells next to fire (packed into one event word per column), the Empty cells among the
//    winners of the step's regrowth lottery (queued directly) - are handled afterwards by a single rolled loop that
//    patches single bytes of the freshly written buffer;
//  * FireStats are kept incrementally: only state transitions are counted (one warp-aggregated atomic
//    when a warp saw any) into the step's delta slot; see forest_publish for who puts the step on record.
// FireStats bookkeeping (forest_fire.rs:104-134, sampled every step by the fused series, time_series.rs:68-87).
// A step only adds the deltas of its state transitions to one of two delta slots (fire-and-forget reductions, no
// fence, no grid-wide ticket on the step's critical path).  The totals are brought up to date — and the record +
// series slot of that step written — by ONE thread afterwards: the first thread of the next step's grid, which runs
// while that grid's loads are in flight, or forest_publish_kernel after the last step of a run.
// stats_run layout (u64): [0..2] totals healthy / burning / empty, [3] last published step,
//                         [4..6] deltas of even steps, [8..10] deltas of odd steps.
__device__ __forceinline__ void forest_publish(const ForestArgs& a, uint32_t slab_w, uint32_t p)
{
    unsigned long long* st = reinterpret_cast<unsigned long long*>(a.stats_run);
    if (p == 0 || __ldcg(st + 3) >= p) return; // nothing ran yet, or that step is already on record
    unsigned long long* d = st + 4 + 4 * (p & 1u);
    const uint64_t healthy = __ldcg(st + 0) + __ldcg(d + 0); // deltas are two's complement: wrapping adds
    const uint64_t burning = __ldcg(st + 1) + __ldcg(d + 1);
    const uint64_t empty = __ldcg(st + 2) + __ldcg(d + 2);
    st[0] = healthy;
    st[1] = burning;
    st[2] = empty;
    st[3] = p;
    d[0] = d[1] = d[2] = 0ull; // the slot is used again two steps later
    const uint64_t total = static_cast<uint64_t>(slab_w) * a.H + a.ov_owned;
    const uint64_t burned = total - healthy - burning - empty;
    a.stats_out[0] = healthy;
    a.stats_out[1] = burning;
    a.stats_out[2] = burned;
    a.stats_out[3] = empty;
    a.stats_out[4] = total;
    if (a.series_values && a.series_interval && (p % a.series_interval) == 0)
    {
        // slot = number of sampling instants before this one (time_series.rs:68-87)
        const uint64_t slot = p / a.series_interval - 1;
        if (slot < a.series_capacity)
        {
            uint64_t* rec = reinterpret_cast<uint64_t*>(a.series_values) + slot * 5;
            rec[0] = healthy;
            rec[1] = burning;
            rec[2] = burned;
            rec[3] = empty;
            rec[4] = total;
            a.series_counts[slot] = total;
        }
    }
}

// After the last step of a run (and at the end of every captured graph of steps): put that step on record and,
// for a graph, advance the device-side StepNumber (src/plugins/step_number.rs:20-23) by the steps it held.
__global__ void forest_publish_kernel(ForestArgs a, uint32_t slab_w, uint32_t bump, ForestPeer pr)
{
    forest_stamp(pr); // row-split grids: the last step of the run is complete
    const uint32_t p = (a.d_step ? *a.d_step : 0u) + a.step_base;
    forest_publish(a, slab_w, p);
    if (bump && a.d_step) *const_cast<uint32_t*>(a.d_step) += bump;
}

// Regrowth level 1 is the block lottery (lottery.cuh): one Philox call per 256 cells, made by up to 3 lanes per column
// (the blocks that intersect the warp's 512 cells of it) while the loads are in flight; the winners go straight into
// the warp's rare-event queue.  (Until round 2 every thread made one call per 16 cells and turned the 16 bytes into
// masks: 45 % of the kernel's instructions.)
// PEER: the fused halo transport is compiled in (row-partitioned grids with peer memory); the single-GPU kernel
// carries none of its registers or branches.
// PDL: the launch is a node of a captured run of steps with programmatic dependencies (run(n) on one GPU).  The grid lets
// its successor's CTAs be scheduled as soon as SM slots free up, and everything that does not depend on the previous
// step - the launch itself, the index arithmetic, the regrowth lottery of the step - runs before `griddepcontrol.wait`,
// i.e. under the previous step's tail; the wait returns when the previous grid has completed and its cells are visible.
template <bool PEER, bool PDL>
__global__ void __launch_bounds__(kTX* kTY, kMinBlocks) forest_step_kernel(const __grid_constant__ ForestConst c)
{
    if (PDL) asm volatile("griddepcontrol.launch_dependents;");
    const ForestArgs& a = c.a;
    // The slab is padded to whole CTA tiles (pitch a multiple of 512 B, columns a multiple of 32 plus the
    // two halo columns; pad bytes are BURNED = inert), so no load or store below needs a bounds predicate.
    const uint32_t cy = blockIdx.x * kTX + threadIdx.x;
    // column group of this CTA.  With the fused halo transport the two boundary groups come first in launch order
    // (groups 1 and last swapped): they wait for the neighbours while the rest of the grid works.
    uint32_t by = blockIdx.y;
    if (PEER && c.peer.my_mail && gridDim.y > 2 && !(c.peer.debug & 1u)) by = by == 1 ? gridDim.y - 1 : (by == gridDim.y - 1 ? 1 : by);
    const uint32_t lx0 = 1 + (by * kTY + threadIdx.y) * kColsPerThread;
    const uint32_t lane = threadIdx.x; // blockDim.x == 32: one warp per threadIdx.y
    const uint32_t tid = threadIdx.y * kTX + threadIdx.x;
    const uint32_t y0 = cy * 16;
    const uint32_t H = static_cast<uint32_t>(a.H);
    constexpr bool chunk_ok = true;

#ifdef INCERTO_FOREST_TRACE
    const unsigned long long tr_start = gtime();
#endif
    __shared__ uint32_t s_done; // warps of this block that finished (see the end of the kernel)
    __shared__ int32_t s_delta[3];
    __shared__ NearOverlays nr;
    __shared__ uint16_t s_evq[kTY][kEventQueue];                                   // rare events of a warp's tile
    // row-split grid: the neighbours learn that my previous step is complete - first thing, before this very thread may
    // have to wait for THEIR stamp below (both ranks stamp, then wait)
    if (PEER && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) forest_stamp(c.peer);
    if (tid < 3) s_delta[tid] = 0;
    if (tid == 0)
    {
        s_done = 0;
        nr.n = 0;
    }

    // StepNumber of this launch: by value, or (captured graphs) the device counter + the launch's offset in the graph
    // (the counter only changes between graphs)
    const uint32_t step = (a.d_step ? *a.d_step : 0u) + a.step_base;
    uint32_t lot_k = 0, lot_z = 0, lot_w = 0, lot_off = 0, lot_total = 0;
    auto draw_lottery = [&]() {
        // ---- level-1 regrowth draws (block lottery): nothing here depends on the cells, so the call runs while the loads
        // are in flight.  Lane 3(j-1)+i draws the i-th block of 256 uids that intersects the warp's cells of column j and
        // keeps the number of winners and the head of the position stream; the events are made after the stores.
        if (a.do_regrow && c.regrow_two_level)
        {
            const uint32_t j = lane / 3 + 1, bi = lane - 3 * (j - 1);
            const uint32_t yw = blockIdx.x * (kTX * 16);
            const uint32_t lo = (static_cast<uint32_t>(a.x_begin) + lx0 - 2 + j) * H + yw;
            const uint32_t len = yw < H ? min(H - yw, static_cast<uint32_t>(kTX * 16)) : 0u;
            const uint32_t blk = (lo >> 8) + bi;
            if (lane < 3 * kColsPerThread && lx0 - 1 + j <= c.slab_w && len != 0 && blk <= ((lo + len - 1) >> 8))
            {
                const uint4 r = philox4x32_10(blk, 0u, step, 0u, c.rk_regrow);
                lot_k = lottery_count((static_cast<uint64_t>(r.x) << 32) | r.y);
                lot_z = r.z;
                lot_w = r.w;
            }
            // queue slots of the lane's winners: exclusive prefix sum of the counts (a winner outside the span leaves a
            // null item), so that the lanes can write their events later without talking to each other
            uint32_t incl = lot_k;
    #pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t up = __shfl_up_sync(kFullMask, incl, o);
                if (lane >= static_cast<uint32_t>(o)) incl += up;
            }
            lot_total = __shfl_sync(kFullMask, incl, 31);
            lot_off = incl - lot_k;
        }
    };
    if (PDL)
    {
        draw_lottery();
        asm volatile("griddepcontrol.wait;" ::: "memory"); // the previous step is complete: its cells may be read
    }

    // ---- loads first.  With the fused halo transport the two warps (per boundary CTA) whose W / E neighbour column is
    // the neighbour rank's boundary column do not read it from the local halo column: they pull it over NVLink below,
    // after their other loads are in flight, straight into the register it is needed in.
    const uint32_t hi_off = (c.slab_w - 1u) % (kTY * kColsPerThread);   // position of the last real column in its group
    const bool pull_lo = PEER && c.peer.my_mail && c.peer.need_lo && by == 0 && threadIdx.y == 0 && !(c.peer.debug & 2u);                // my v[0]
    const bool pull_hi = PEER && c.peer.my_mail && c.peer.need_hi && by == gridDim.y - 1 && threadIdx.y == hi_off / kColsPerThread && !(c.peer.debug & 2u);
    const uint32_t jh = hi_off % kColsPerThread + 2u;                                                          // my v[jh]
    const uint8_t* p0 = a.cur + static_cast<size_t>(lx0 - 1) * c.pitch + y0;
    uint4 v[kColsPerThread + 2];
#pragma unroll
    for (int j = 0; j < kColsPerThread + 2; ++j) v[j] = *reinterpret_cast<const uint4*>(p0 + static_cast<size_t>(j) * c.pitch);
    // (a pulling warp's load of the stale local halo column is simply overwritten below: no extra predicate or register
    // on everybody else's path)
    // ---- fused halo transport, pull side (warp-uniform: threadIdx.y is).  The pulling warp waits for the neighbour's
    // stamp of the previous step (nothing else in this grid depends on it, so it may spin), fetches its chunk of the
    // neighbour's boundary column and also leaves it in the local halo column: the rare-event and overlay paths of this
    // CTA read single cells from there after the barrier below.  The other warps of the CTA never wait for NVLink.
    if (PEER && (pull_lo || pull_hi))
    {
        if (lane == 0)
        {
            // Acquire loads of the stamp, not a volatile poll + system-scope fence: the fence also waited for this
            // warp's own cell loads (just issued, microseconds from a cold L2) before the pull could even start, which
            // put the boundary warps - and with them the end of the step - 3-5 us behind everybody else.
            const long long t0 = clock64();
            // stamps only grow; wrap-around-safe comparison (the stamp words live in the neighbours' memory)
            while ((pull_lo && static_cast<int32_t>(ld_acquire_sys(c.peer.mail_lo + kStampSlot) - c.peer.need_lo) < 0) ||
                   (pull_hi && static_cast<int32_t>(ld_acquire_sys(c.peer.mail_hi + kStampSlot) - c.peer.need_hi) < 0))
            {
                __nanosleep(64);
                if (clock64() - t0 > 20000000000ll) // ~10 s: a neighbour died; never hang the device
                {
                    if (c.error) atomicCAS(c.error, 0u, 39u); // INCERTO_ERR_COMM
                    break;
                }
            }
        }
        __syncwarp(); // orders the other lanes' pulls after lane 0's acquire
        uint8_t* mine = const_cast<uint8_t*>(a.cur);
        if (pull_lo)
        {
            v[0] = __ldcv(reinterpret_cast<const uint4*>(c.peer.src_lo + y0));
            *reinterpret_cast<uint4*>(mine + y0) = v[0];
        }
        if (pull_hi)
        {
            const uint4 q = __ldcv(reinterpret_cast<const uint4*>(c.peer.src_hi + y0));
#pragma unroll
            for (int j = 1; j < kColsPerThread + 2; ++j)
                if (static_cast<uint32_t>(j) == jh) v[j] = q;
            *reinterpret_cast<uint4*>(mine + static_cast<size_t>(c.slab_w + 1) * c.pitch + y0) = q;
        }
        __syncwarp();
    }
    // the word just outside the warp's y span, per column: lane 0 needs the cell y0-1 (top byte of the word
    // before its chunk), lane 31 the cell y0+16 (bottom byte of the word after it)
    // (bit j-1 of `edge`: that cell of my column j is burning).  The 2 x kColsPerThread words are fetched by as many
    // lanes with ONE load instruction - lane i < C the word above column i+1, lane C+i the word below it - and handed to
    // lanes 0 and 31 by a vote.
    uint32_t edge;
    {
        const bool top = lane < kColsPerThread, bot = !top && lane < 2 * kColsPerThread;
        const uint32_t jj = top ? lane + 1 : lane + 1 - kColsPerThread;
        bool burning = false;
        if ((top && blockIdx.x > 0) || (bot && blockIdx.x + 1 < gridDim.x))
        {
            const uint8_t* span = p0 - lane * 16 + static_cast<size_t>(jj) * c.pitch; // the warp's first cell of column jj
            burning = (*reinterpret_cast<const uint32_t*>(span + (top ? -4 : kTX * 16)) & (top ? 0x3f000000u : 0x0000003fu)) != 0;
        }
        const uint32_t eb = __ballot_sync(kFullMask, burning);
        edge = lane == 0 ? (eb & ((1u << kColsPerThread) - 1u)) : (lane == 31 ? ((eb >> kColsPerThread) & ((1u << kColsPerThread) - 1u)) : 0u);
    }
    if (!PDL) draw_lottery();
    // the previous step's FireStats go on record now, off every critical path (see forest_publish)
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) forest_publish(a, c.slab_w, step - 1);

    // ---- overlays that can touch this CTA's tile (usually none): one bit per tile, prepared by the host.
    // The staging into shared memory (one more barrier) only happens in the CTAs that have one nearby.  Overlay k's
    // position, owner and old state are fetched by thread k right here, in the shadow of the cell loads: the staging
    // below then has no dependent global load left (it used to add two memory latencies to the three CTAs next to the
    // example's igniters, i.e. to the end of every step: 4096^2 cold 19.0 -> 16.9 us without overlays).
    const uint32_t nov = a.n_overlay;
    uint32_t n_near = 0;
    bool any_near = false;
    int pf_x = 0, pf_y = 0;
    uint32_t pf_owner = 0, pf_old = 0;
    if (nov)
    {
        const uint32_t tile = by * gridDim.x + blockIdx.x;
        any_near = (c.near_bits[tile >> 5] >> (tile & 31)) & 1u;
        if (tid < nov)
        {
            pf_x = a.ov_pos[2 * tid];
            pf_y = a.ov_pos[2 * tid + 1];
            pf_owner = c.ov_owner ? c.ov_owner[tid] : 0;
            // (a neighbour rank's overlay state lives in ITS memory: fetched below, by the few CTAs that have the entity
            // nearby - as a prefetch in every CTA it was an NVLink round trip no warp 0 of the grid could retire before)
            pf_old = pf_owner == 0 ? a.ov_cur[tid] : 0u;
        }
    }
    __syncthreads(); // s_done and nr.n are initialised; all warps are still in step here, so this barrier is cheap
    if (any_near) // CTA-uniform
    {
        const int tx0 = a.x_begin + static_cast<int>(by * kTY * kColsPerThread);
        const int ty0 = static_cast<int>(blockIdx.x * kTX * 16);
        for (uint32_t k = tid; k < nov; k += kTX * kTY)
        {
            int ox = pf_x, oy = pf_y;
            uint32_t o = pf_owner, old = pf_old;
            if (k == tid && (o == 1 || o == 2)) old = o == 1 ? (c.ov_lo ? c.ov_lo[k] : 0) : (c.ov_hi ? c.ov_hi[k] : 0);
            if (k != tid) // more overlays than threads: the rest is fetched here
            {
                ox = a.ov_pos[2 * k];
                oy = a.ov_pos[2 * k + 1];
                o = c.ov_owner ? c.ov_owner[k] : 0;
                old = o == 0 ? a.ov_cur[k] : (o == 1 ? (c.ov_lo ? c.ov_lo[k] : 0) : (o == 2 ? (c.ov_hi ? c.ov_hi[k] : 0) : 0));
            }
            if (ox >= tx0 - 1 && ox <= tx0 + kTY * kColsPerThread && oy >= ty0 - 1 && oy <= ty0 + kTX * 16)
            {
                const uint32_t i = atomicAdd(&nr.n, 1u);
                if (i < kMaxNear)
                {
                    nr.x[i] = static_cast<int16_t>(ox);
                    nr.y[i] = static_cast<int16_t>(oy);
                    nr.k[i] = static_cast<uint16_t>(k);
                    nr.owner[i] = static_cast<uint8_t>(o);
                    nr.old[i] = static_cast<uint8_t>(old);
                }
                else if (c.error)
                    atomicCAS(c.error, 0u, 37u); // INCERTO_ERR_UNSUPPORTED: too many overlay entities around one tile
            }
        }
        __syncthreads();
        n_near = nr.n < kMaxNear ? nr.n : kMaxNear;
    }

    // ---- is there any fire in what this warp loaded, and in which columns?  Bit j of `wf` (one warp-wide OR): some
    // lane holds a burning cell in column j of its eight (the edge cells count for their column).
    uint32_t fl = edge << 1;
#pragma unroll
    for (int j = 0; j < kColsPerThread + 2; ++j) fl |= (((v[j].x | v[j].y | v[j].z | v[j].w) & 0x3f3f3f3fu) ? 1u : 0u) << j;
    const uint32_t wf = __reduce_or_sync(kFullMask, fl);
    const bool fire = wf != 0;
#ifdef INCERTO_FOREST_TRACE
    const unsigned long long tr_loaded = gtime();
#endif

    const uint32_t ign = ignited_value(a);
    Delta dl;
    uint32_t ev[kColsPerThread] = {};           // per column: bit 8b+k = spread candidate, bit 8b+4+k = regrowth
                                                // draw owed, for cell 4k+b of the chunk
    uint8_t* q0 = a.next + (p0 - a.cur);
    if (!fire) // warp-uniform, and by far the common case: no progression, no spread — the chunk is copied
    {
#pragma unroll
        for (int j = 1; j <= kColsPerThread; ++j)
            if (lx0 - 1 + j <= c.slab_w) *reinterpret_cast<uint4*>(q0 + static_cast<size_t>(j) * c.pitch) = v[j];
    }
    else
    {
#pragma unroll
        for (int j = 1; j <= kColsPerThread; ++j)
        {
            if (!((wf >> (j - 1)) & 7u)) // warp-uniform: no fire in this column nor in its W / E neighbours - copied
            {
                if (lx0 - 1 + j <= c.slab_w) *reinterpret_cast<uint4*>(q0 + static_cast<size_t>(j) * c.pitch) = v[j];
                continue;
            }
            const uint32_t cw[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
            uint32_t B[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) B[k] = burn_bits(cw[k]);
            uint32_t prev = __shfl_up_sync(kFullMask, B[3] >> 24, 1);
            uint32_t next = __shfl_down_sync(kFullMask, B[0] & 1u, 1);
            if (lane == 0) prev = (edge >> (j - 1)) & 1u;
            if (lane == 31) next = (edge >> (j - 1)) & 1u;
            if (lx0 - 1 + j > c.slab_w) continue; // warp-uniform; only in the last, partial column group
            uint32_t nw[4] = {cw[0], cw[1], cw[2], cw[3]};
            // ---- progression of cells already burning (:340-349); Burning{1} -> Burned leaves the burning count
            if (a.do_progress)
            {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                {
                    nw[k] = cw[k] - B[k];
                    dl.b -= __popc(zero_bytes((cw[k] & 0x3f3f3f3fu) ^ 0x01010101u));
                }
            }
            // ---- spread (:282-329): healthy cells with a burning orthogonal neighbour take draws below
            if (a.do_spread)
            {
                const uint32_t ww[4] = {v[j - 1].x, v[j - 1].y, v[j - 1].z, v[j - 1].w};
                const uint32_t ew[4] = {v[j + 1].x, v[j + 1].y, v[j + 1].z, v[j + 1].w};
                uint32_t evj = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                {
                    const uint32_t bn = __funnelshift_l(k == 0 ? (prev << 24) : B[k - 1], B[k], 8);
                    const uint32_t bs = __funnelshift_r(B[k], k == 3 ? next : B[k + 1], 8);
                    const uint32_t src = bn | burn_bits(ww[k]) | burn_bits(ew[k]) | bs;
                    evj |= (src & (cw[k] >> 6) & 0x01010101u) << k;
                }
                ev[j - 1] = evj;
            }
            *reinterpret_cast<uint4*>(q0 + static_cast<size_t>(j) * c.pitch) = make_uint4(nw[0], nw[1], nw[2], nw[3]);
        }
    }

    // ---- regrowth (:354-365): only EMPTY cells draw.  Two-level: the lottery winners are already queued; a
    // probability above 1/256 has no level 1 and every Empty cell takes its draw in the rare-event loop.
    if (a.do_regrow && !c.regrow_two_level)
    {
#pragma unroll
        for (int j = 1; j <= kColsPerThread; ++j)
        {
            const uint32_t e0 = v[j].x & 0x80808080u, e1 = v[j].y & 0x80808080u, e2 = v[j].z & 0x80808080u,
                           e3 = v[j].w & 0x80808080u;
            if (lx0 - 1 + j <= c.slab_w) ev[j - 1] |= (e0 >> 3) | (e1 >> 2) | (e2 >> 1) | e3;
        }
    }

#ifdef INCERTO_FOREST_TRACE
    const unsigned long long tr_stored = gtime();
#endif
    // ---- regrowth events of the lottery drawn above: every drawing lane walks its position stream - skipping a byte
    // that repeats an earlier one (the first eight bytes sit in two registers: byte-parallel compare) - and writes its
    // winners into its own slots of the warp's event queue (a null item for a winner outside the span).  No loads, no
    // votes; whether the cell is Empty is looked at when the event is handled.  No out-of-line path either: a rarely
    // executed stretch of code costs a small grid microseconds of instruction fetches.  This half sits after the
    // stores (with the walk in the shadow of the loads the compiler spilled two of the eight cell vectors, i.e. made
    // the warp wait for them first).
    uint32_t queued = lot_total; // warp-uniform
    {
        const uint32_t j = lane / 3 + 1, bi = lane - 3 * (j - 1);
        const uint32_t yw = blockIdx.x * (kTX * 16);
        const uint32_t lo = (static_cast<uint32_t>(a.x_begin) + lx0 - 2 + j) * H + yw; // uid of the span's first cell
        const uint32_t len = yw < H ? min(H - yw, static_cast<uint32_t>(kTX * 16)) : 0u;
        const uint32_t blk = (lo >> 8) + bi;
        uint32_t need = lot_k; // stream bytes this lane still has to look at: winners left + repeats met
        // regrowth event of a winner: bit 15 | column << 9 | offset in the warp's span (null when outside the span)
        auto winner = [&](uint32_t b) {
            const uint32_t yo = ((blk << 8) | b) - lo;
            s_evq[threadIdx.y][lot_off++] = yo < len ? static_cast<uint16_t>(0x8000u | ((j - 1) << 9) | yo) : kNullEvent;
        };
        // the first four bytes (one register) with compile-time masks: all there is for 99 % of the blocks
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if (static_cast<uint32_t>(t) < need)
            {
                const uint32_t b = (lot_z >> (8 * t)) & 0xffu;
                if (t > 0 && (zero_bytes(lot_z ^ (b * 0x01010101u)) & ((1u << (8 * t)) - 1u)))
                    ++need;
                else
                    winner(b);
            }
        for (uint32_t t = 4; t < need; ++t)
        {
            uint32_t b;
            bool rep;
            if (t < 8)
            {
                b = (lot_w >> (8 * (t & 3u))) & 0xffu;
                const uint32_t r4 = b * 0x01010101u;
                rep = (zero_bytes(lot_z ^ r4) | (zero_bytes(lot_w ^ r4) & ((1u << (8 * (t - 4))) - 1u))) != 0;
            }
            else // beyond the first eight bytes (about one block in a million)
            {
                b = lottery_stream_byte(c.rk_regrow, blk, 0u, step, lot_z, lot_w, t);
                rep = false;
                for (uint32_t h = 0; h < t; ++h) rep = rep || lottery_stream_byte(c.rk_regrow, blk, 0u, step, lot_z, lot_w, h) == b;
            }
            if (rep)
                ++need;
            else
                winner(b);
        }
    }
#ifdef INCERTO_FOREST_TRACE
    const unsigned long long tr_lot = gtime();
#endif
    // `queued` lottery winners are waiting in the queue (first round only).  Warps without fire (nearly all) have nothing
    // else: they skip the compaction of the event words.
    bool packing = fire || (a.do_regrow && !c.regrow_two_level); // warp-uniform: event words may hold something
    for (;;)
    {
        uint32_t total = queued;
        if (packing)
        {
            uint32_t mine = 0;
#pragma unroll
            for (int j = 0; j < kColsPerThread; ++j) mine += __popc(ev[j]);
            // exclusive prefix sum of the per-lane event counts
            uint32_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t up = __shfl_up_sync(kFullMask, incl, o);
                if (lane >= static_cast<uint32_t>(o)) incl += up;
            }
            total += __shfl_sync(kFullMask, incl, 31);
            uint32_t slot = incl - mine + queued;
#pragma unroll
            for (int j = 0; j < kColsPerThread; ++j)
                while (ev[j] && slot < kEventQueue)
                {
                    const uint32_t bit = __ffs(ev[j]) - 1;
                    ev[j] &= ev[j] - 1;
                    s_evq[threadIdx.y][slot++] = static_cast<uint16_t>((lane << 10) | (static_cast<uint32_t>(j) << 5) | bit);
                }
            packing = total > kEventQueue; // leftovers go through another round
        }
        queued = 0;
        if (!total) break;
        __syncwarp(); // also orders the owners' vector stores of `next` before the byte patches below
        const uint32_t n_ev = total < kEventQueue ? total : kEventQueue;
        for (uint32_t k = lane; k < n_ev; k += 32)
        {
            const uint32_t item = s_evq[threadIdx.y][k];
            if (item == kNullEvent) continue;
            // spread / one-level regrowth event: lane << 10 | column << 5 | bit of the event word; lottery winner: see above
            const bool won = (item & 0x8000u) != 0;
            const uint32_t bit = won ? 4u : (item & 31u);
            const uint32_t lx = lx0 + (won ? ((item >> 9) & 7u) : ((item >> 5) & 31u));
            const uint32_t y = blockIdx.x * (kTX * 16) + (won ? (item & 511u) : ((item >> 10) * 16 + 4 * (bit & 3) + ((bit >> 3) & 3)));
            const size_t off = static_cast<size_t>(lx) * c.pitch + y;
            const uint32_t uid = (static_cast<uint32_t>(a.x_begin) + lx - 1) * H + y;
            if (!(bit & 4))
            {
                // (a) spread draws of a healthy cell at the fire front: word d of one call, d = N, W, E, S
                const uint8_t* pc = a.cur + off;
                uint32_t dirs = 0;
                if (y > 0 && (pc[-1] & 0x3f)) dirs |= 1u;
                if (pc[-static_cast<ptrdiff_t>(c.pitch)] & 0x3f) dirs |= 2u;
                if (pc[c.pitch] & 0x3f) dirs |= 4u;
                if (y + 1 < H && (pc[1] & 0x3f)) dirs |= 8u;
                const uint4 r = philox4x32_10(uid, 0u, step, 0u, c.rk_spread);
                const bool hit = ((dirs & 1u) && bernoulli(r.x, a.thr_spread)) || ((dirs & 2u) && bernoulli(r.y, a.thr_spread)) ||
                                 ((dirs & 4u) && bernoulli(r.z, a.thr_spread)) || ((dirs & 8u) && bernoulli(r.w, a.thr_spread));
                if (hit)
                {
                    a.next[off] = static_cast<uint8_t>(ign);
                    dl.h -= 1;
                    dl.b += ign ? 1 : 0;
                }
            }
            else
            {
                // (b) regrowth draw still owed by an Empty cell: level 2 (or the only level)
                const uint4 l2 = philox4x32_10(uid >> 2, 0u, step, 1u, c.rk_regrow);
                if (a.cur[off] == 0x80u && bernoulli(word_of(l2, uid & 3u), c.thr_regrow_l2))
                {
                    a.next[off] = 0x40;
                    dl.e -= 1;
                    dl.h += 1;
                }
            }
        }
        __syncwarp();
    }
#ifdef INCERTO_FOREST_TRACE
    const unsigned long long tr_events = gtime();
#endif
    // (c) cells an overlay entity can influence, and the overlay entities living in this thread's tile
    if (n_near && chunk_ok)
    {
        const int gx0 = a.x_begin + static_cast<int>(lx0) - 1; // my columns: gx0 .. gx0 + kColsPerThread - 1
        for (uint32_t i = 0; i < n_near; ++i)
        {
            const int ox = nr.x[i], oy = nr.y[i];
            if (ox < gx0 - 1 || ox > gx0 + kColsPerThread || oy < static_cast<int>(y0) - 1 || oy > static_cast<int>(y0) + 16)
                continue;
            const bool src_burning = (nr.old[i] & 0x3f) != 0;
            for (int d = 0; d < 5; ++d)
            {
                // the overlay's own cell, then N, W, E, S of it; a neighbour is only affected while the overlay burns
                if (d > 0 && !src_burning) break;
                const int px = ox + (d == 2 ? -1 : (d == 3 ? 1 : 0));
                const int py = oy + (d == 1 ? -1 : (d == 4 ? 1 : 0));
                if (px < gx0 || px >= gx0 + kColsPerThread || px >= a.x_end || py < static_cast<int>(y0) ||
                    py >= static_cast<int>(y0) + 16 || py >= a.H)
                    continue;
                // several overlays may make a thread visit the same cell twice: the patch is idempotent
                const bool m = overlay_marked(c, nr, step, px, py);
                if (m)
                {
                    // every entity at a marked position becomes Burning{duration} (:320-328)
                    uint8_t* nb = a.next + static_cast<size_t>(px - a.x_begin + 1) * c.pitch + py;
                    dl.add(*nb, -1);
                    *nb = static_cast<uint8_t>(ign);
                    dl.add(ign, +1);
                }
                if (d == 0 && nr.owner[i] == 0)
                {
                    // the overlay entity itself: progression / regrowth, overridden by ignition
                    const uint32_t s0 = nr.old[i];
                    uint32_t s1 = s0;
                    if (a.do_progress && (s1 & 0x3f)) s1 -= 1;
                    if (a.do_regrow && s1 == 0x80 && regrow_draw(c, a.ov_uid_base + nr.k[i], step)) s1 = 0x40;
                    if (m) s1 = ign;
                    a.ov_next[nr.k[i]] = static_cast<uint8_t>(s1);
                    dl.add(s0, -1);
                    dl.add(s1, +1);
                }
            }
        }
    }

#ifdef INCERTO_FOREST_TRACE
    const unsigned long long tr_overlays = gtime();
#endif
    // ---- incremental FireStats: warps that saw a transition add their deltas to the block's counters
    if (__any_sync(kFullMask, (dl.h | dl.b | dl.e) != 0))
    {
        const int32_t h = __reduce_add_sync(kFullMask, dl.h), b = __reduce_add_sync(kFullMask, dl.b), e = __reduce_add_sync(kFullMask, dl.e);
        if (lane == 0)
        {
            if (h) atomicAdd(&s_delta[0], h);
            if (b) atomicAdd(&s_delta[1], b);
            if (e) atomicAdd(&s_delta[2], e);
        }
    }
    // Completion without a trailing block barrier (warps finish at very different times because of the
    // rare-event loops): every warp arrives on a shared counter; the last warp of the block sends the
    // block's deltas to the step's delta slot (forest_publish explains who puts the step on record).
#ifdef INCERTO_FOREST_TRACE
    if (lane == 0)
    {
        const uint32_t w = ((blockIdx.y * gridDim.x + blockIdx.x) * kTY + threadIdx.y) & ((1u << 16) - 1);
        g_forest_trace[8 * w + 0] = tr_start;
        g_forest_trace[8 * w + 1] = tr_loaded;
        g_forest_trace[8 * w + 2] = gtime();
        g_forest_trace[8 * w + 3] = smid() | (static_cast<unsigned long long>(fire) << 32);
        g_forest_trace[8 * w + 4] = tr_stored;
        g_forest_trace[8 * w + 5] = tr_lot;
        g_forest_trace[8 * w + 6] = tr_events;
        g_forest_trace[8 * w + 7] = tr_overlays;
    }
#endif
    if (lane != 0) return;
    __threadfence_block();
    if (atomicAdd(&s_done, 1u) != kTY - 1) return;
    {
        // the block's deltas go to this step's delta slot: reductions nobody waits for (the kernel boundary orders
        // them before the thread that publishes the step)
        unsigned long long* d = reinterpret_cast<unsigned long long*>(a.stats_run) + 4 + 4 * (step & 1u);
        const int32_t h = atomicAdd(&s_delta[0], 0), b = atomicAdd(&s_delta[1], 0), e = atomicAdd(&s_delta[2], 0);
        if (h) atomicAdd(d + 0, static_cast<unsigned long long>(static_cast<long long>(h)));
        if (b) atomicAdd(d + 1, static_cast<unsigned long long>(static_cast<long long>(b)));
        if (e) atomicAdd(d + 2, static_cast<unsigned long long>(static_cast<long long>(e)));
    }
}

__global__ void forest_spawn_kernel(uint8_t* cells, int32_t H, uint32_t pitch, int32_t x_begin, uint32_t slab_w,
                                    PhiloxKey key, uint64_t thr)
{
    // forest_fire.rs:239-257: Healthy with INITIAL_FOREST_DENSITY else Empty;
    // draw = word (uid & 3) of philox(key[FOREST_SPAWN], (uid >> 2, 0, 0, 0))
    const uint64_t ncells = static_cast<uint64_t>(slab_w) * H;
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < ncells; i += gstride)
    {
        const uint32_t lx = static_cast<uint32_t>(i / H), y = static_cast<uint32_t>(i % H);
        const uint32_t uid = (static_cast<uint32_t>(x_begin) + lx) * H + y;
        const uint4 r = philox4x32_10(uid >> 2, 0u, 0u, 0u, key);
        cells[static_cast<size_t>(lx + 1) * pitch + y] = bernoulli(word_of(r, uid & 3u), thr) ? 0x40 : 0x80;
    }
}

__global__ void __launch_bounds__(256)
    forest_stats_kernel(const uint8_t* __restrict__ slab, uint32_t H, uint32_t pitch, uint32_t slab_w, uint64_t* out5,
                        uint64_t* partials, uint32_t* ticket)
{
    uint64_t c[4] = {0, 0, 0, 0};
    const uint64_t n = static_cast<uint64_t>(slab_w) * H;
    const uint64_t gstride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += gstride)
    {
        const uint64_t lx = i / H, y = i % H;
        const uint8_t s = slab[(lx + 1) * pitch + y];
        if (s & 0x40)
            ++c[0];
        else if (s & 0x80)
            ++c[3];
        else if (s & 0x3f)
            ++c[1];
        else
            ++c[2];
    }
    __shared__ uint64_t smem[4 * 32];
    block_sum<uint64_t, 4>(c, smem);
    if (threadIdx.x == 0)
        for (int i = 0; i < 4; ++i) partials[4 * blockIdx.x + i] = c[i];
    if (last_block_arrives(ticket, gridDim.x))
    {
        if (threadIdx.x == 0)
        {
            uint64_t t[4] = {0, 0, 0, 0};
            for (uint32_t b = 0; b < gridDim.x; ++b)
                for (int i = 0; i < 4; ++i) t[i] += partials[4 * b + i];
            out5[0] = t[0];
            out5[1] = t[1];
            out5[2] = t[2];
            out5[3] = t[3];
            out5[4] = n;
        }
    }
}

} // namespace

// Extended launch: the engine passes the neighbour overlay arrays separately.
cudaError_t launch_forest_step_ex(const ForestArgs& a, const uint8_t* ov_lo, const uint8_t* ov_hi, const uint8_t* ov_owner,
                                  cudaStream_t st, const ForestPeer* peer, bool pdl)
{
    ForestConst c;
    if (peer) c.peer = *peer;
    c.a = a;
    c.slab_w = static_cast<uint32_t>(a.x_end - a.x_begin);
    c.pitch = forest_pitch(a.H);
    c.chunks_per_col = c.pitch / 16;
    c.regrow_two_level = a.regrow_two_level;
    c.thr_regrow_l2 = a.regrow_two_level ? a.thr_regrow_l2 : a.thr_regrow;
    c.ov_lo = ov_lo;
    c.ov_hi = ov_hi;
    c.ov_owner = ov_owner;
    c.near_bits = a.near_bits;
    c.error = a.error;
    c.rk_spread = make_round_keys(a.key_spread);
    c.rk_regrow = make_round_keys(a.key_regrow);
    dim3 block(kTX, kTY);
    dim3 grid((c.chunks_per_col + kTX - 1) / kTX, (c.slab_w + kTY * kColsPerThread - 1) / (kTY * kColsPerThread));
    if (peer != nullptr)
        forest_step_kernel<true, false><<<grid, block, 0, st>>>(c);
    else if (pdl && static_cast<uint64_t>(grid.x) * grid.y <= 4ull * kNumSMs * kMinBlocks)
    {
        // (grids of many waves gain nothing from the overlap - 16384^2: 92.9 us without, 94.2 us with - and keep the
        // plain edge)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = block;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, forest_step_kernel<false, true>, c);
    }
    else
        forest_step_kernel<false, false><<<grid, block, 0, st>>>(c);
    return cudaGetLastError();
}

cudaError_t launch_forest_publish(const ForestArgs& a, uint32_t bump, cudaStream_t st, const ForestPeer* peer)
{
    ForestPeer pr;
    if (peer) pr = *peer;
    forest_publish_kernel<<<1, 1, 0, st>>>(a, static_cast<uint32_t>(a.x_end - a.x_begin), bump, pr);
    return cudaGetLastError();
}

cudaError_t launch_forest_step(const ForestArgs& a, cudaStream_t st) { return launch_forest_step_ex(a, nullptr, nullptr, nullptr, st); }

uint32_t forest_pitch(int32_t H) { return (static_cast<uint32_t>(H) + kTX * 16 - 1) / (kTX * 16) * (kTX * 16); }
uint32_t forest_alloc_cols(int32_t slab_w)
{
    return (static_cast<uint32_t>(slab_w) + kTY * kColsPerThread - 1) / (kTY * kColsPerThread) * (kTY * kColsPerThread) + 2;
}
void forest_tile_shape(uint32_t* cols, uint32_t* rows)
{
    *cols = kTY * kColsPerThread;
    *rows = kTX * 16;
}

namespace
{
// block 0 serves the lower neighbour, block 1 the upper one.  The stores go over NVLink into the neighbour's HBM; the
// system-scope fence orders them before the mailbox stamp the neighbour's wait kernel polls.
__global__ void __launch_bounds__(256)
    forest_halo_push_kernel(const uint8_t* col_lo, uint8_t* peer_lo, const uint8_t* col_hi, uint8_t* peer_hi, uint32_t col_bytes,
                            const uint8_t* ov, uint8_t* peer_ov_lo, uint8_t* peer_ov_hi, uint32_t n_ov, uint32_t* peer_mail_lo,
                            uint32_t* peer_mail_hi, uint32_t stamp)
{
    const uint8_t* src = blockIdx.x == 0 ? col_lo : col_hi;
    uint8_t* dst = blockIdx.x == 0 ? peer_lo : peer_hi;
    uint8_t* ov_dst = blockIdx.x == 0 ? peer_ov_lo : peer_ov_hi;
    // the lower neighbour knows me as its upper one (mailbox slot 1) and vice versa (slot 0)
    uint32_t* mail = blockIdx.x == 0 ? peer_mail_lo + 1 : peer_mail_hi;
    if (!dst) return; // no neighbour on this side
    for (uint32_t i = threadIdx.x; i < col_bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
    for (uint32_t i = threadIdx.x; i < n_ov; i += blockDim.x) ov_dst[i] = ov[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
    {
        *reinterpret_cast<volatile uint32_t*>(mail) = stamp;
        __threadfence_system();
    }
}

__global__ void forest_halo_wait_kernel(const uint32_t* mail, uint32_t need_lo, uint32_t need_hi, uint32_t* error)
{
    const volatile uint32_t* m = mail;
    const long long t0 = clock64();
    // stamps only grow; wrap-around-safe comparison
    while (static_cast<int32_t>(m[0] - need_lo) < 0 || static_cast<int32_t>(m[1] - need_hi) < 0)
    {
        __nanosleep(200);
        if (clock64() - t0 > 20000000000ll) // ~10 s: a neighbour died; never hang the device
        {
            if (atomicCAS(error, 0u, static_cast<uint32_t>(39 /* INCERTO_ERR_COMM */)) == 0u) error[1] = m[0] | (m[1] << 16);
            return;
        }
    }
    __threadfence_system();
}
} // namespace

void launch_forest_halo_push(const uint8_t* col_lo, uint8_t* peer_lo, const uint8_t* col_hi, uint8_t* peer_hi,
                             uint32_t col_bytes, const uint8_t* ov, uint8_t* peer_ov_lo, uint8_t* peer_ov_hi, uint32_t n_ov,
                             uint32_t* peer_mail_lo, uint32_t* peer_mail_hi, uint32_t stamp, cudaStream_t st)
{
    forest_halo_push_kernel<<<2, 256, 0, st>>>(col_lo, peer_lo, col_hi, peer_hi, col_bytes, ov, peer_ov_lo, peer_ov_hi, n_ov,
                                               peer_mail_lo, peer_mail_hi, stamp);
}
namespace
{
// Neighbour rendezvous over the peer mailboxes (slot kEpochSlot: the owner's epoch, read by its neighbours over NVLink):
// the ranks of a row-split grid leave it within one NVLink latency of each other.  Bench hygiene (run_cold): the flush
// before a timed step takes a different time on every GPU.
__global__ void forest_rendezvous_kernel(uint32_t* mail, const uint32_t* peer_mail_lo, const uint32_t* peer_mail_hi, uint32_t epoch, uint32_t* error)
{
    *reinterpret_cast<volatile uint32_t*>(mail + kEpochSlot) = epoch;
    __threadfence_system();
    const long long t0 = clock64();
    while ((peer_mail_lo && static_cast<int32_t>(ld_acquire_sys(peer_mail_lo + kEpochSlot) - epoch) < 0) ||
           (peer_mail_hi && static_cast<int32_t>(ld_acquire_sys(peer_mail_hi + kEpochSlot) - epoch) < 0))
    {
        __nanosleep(64);
        if (clock64() - t0 > 20000000000ll) // ~10 s: a neighbour died; never hang the device
        {
            atomicCAS(error, 0u, static_cast<uint32_t>(39 /* INCERTO_ERR_COMM */));
            return;
        }
    }
}
// wait until both neighbours announced step `need` (fused transport: reset must not respawn under a neighbour's pull)
__global__ void forest_stamp_wait_kernel(const uint32_t* peer_mail_lo, const uint32_t* peer_mail_hi, uint32_t need, uint32_t* error)
{
    const long long t0 = clock64();
    while ((peer_mail_lo && static_cast<int32_t>(ld_acquire_sys(peer_mail_lo + kStampSlot) - need) < 0) ||
           (peer_mail_hi && static_cast<int32_t>(ld_acquire_sys(peer_mail_hi + kStampSlot) - need) < 0))
    {
        __nanosleep(200);
        if (clock64() - t0 > 20000000000ll)
        {
            atomicCAS(error, 0u, static_cast<uint32_t>(39 /* INCERTO_ERR_COMM */));
            return;
        }
    }
}
} // namespace
void launch_forest_stamp_wait(const uint32_t* peer_mail_lo, const uint32_t* peer_mail_hi, uint32_t need, uint32_t* error, cudaStream_t st)
{
    forest_stamp_wait_kernel<<<1, 1, 0, st>>>(peer_mail_lo, peer_mail_hi, need, error);
}
void launch_forest_rendezvous(uint32_t* mail, const uint32_t* peer_mail_lo, const uint32_t* peer_mail_hi, uint32_t epoch, uint32_t* error, cudaStream_t st)
{
    forest_rendezvous_kernel<<<1, 1, 0, st>>>(mail, peer_mail_lo, peer_mail_hi, epoch, error);
}

void launch_forest_halo_wait(const uint32_t* mail, uint32_t need_lo, uint32_t need_hi, uint32_t* error, cudaStream_t st)
{
    forest_halo_wait_kernel<<<1, 1, 0, st>>>(mail, need_lo, need_hi, error);
}

int forest_step_grid_blocks(int32_t slab_w, int32_t H)
{
    const uint32_t chunks = forest_pitch(H) / 16;
    return static_cast<int>(((chunks + kTX - 1) / kTX) *
                            ((static_cast<uint32_t>(slab_w) + kTY * kColsPerThread - 1) / (kTY * kColsPerThread)));
}

void launch_forest_spawn(uint8_t* cells, int32_t W, int32_t H, int32_t x_begin, int32_t x_end, PhiloxKey key,
                         uint64_t thr_density, cudaStream_t st)
{
    (void)W;
    const uint32_t pitch = forest_pitch(H);
    forest_spawn_kernel<<<kNumSMs * 8, 256, 0, st>>>(cells, H, pitch, x_begin, static_cast<uint32_t>(x_end - x_begin),
                                                     key, thr_density);
}

void launch_forest_stats(const uint8_t* slab, int32_t H, int32_t slab_w, uint64_t* out5, void* scratch,
                         uint32_t* ticket, cudaStream_t st)
{
    const uint64_t n = static_cast<uint64_t>(slab_w) * H;
    uint64_t b = (n + 256 * 16 - 1) / (256 * 16);
    if (b < 1) b = 1;
    if (b > static_cast<uint64_t>(kNumSMs) * 8) b = kNumSMs * 8;
    const uint32_t pitch = forest_pitch(H);
    forest_stats_kernel<<<static_cast<int>(b), 256, 0, st>>>(slab, static_cast<uint32_t>(H), pitch,
                                                             static_cast<uint32_t>(slab_w), out5,
                                                             static_cast<uint64_t*>(scratch), ticket);
}

} // namespace incerto

#ifdef INCERTO_FOREST_TRACE
extern "C" int incerto_debug_forest_trace(unsigned long long* out, unsigned long long n_words)
{
    return static_cast<int>(cudaMemcpyFromSymbol(out, incerto::g_forest_trace, n_words * sizeof(unsigned long long)));
}
#endif
