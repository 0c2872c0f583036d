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
// Each column is `pitch` = round_up(H,16) bytes; pad bytes and out-of-grid halo
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
//     byte (uid & 15) of philox(key[REGROW], (uid >> 4, 0, step, 0)) == 0
//     AND word (uid & 3) of philox(key[REGROW], (uid >> 2, 0, step, 1)) < floor(256*p*2^32)
//   regrow, p*256 > 1: word (uid & 3) of philox(key[REGROW], (uid >> 2, 0, step, 1)) < floor(p*2^32)
#include "common.cuh"
#include "kernels.hpp"

namespace incerto
{
namespace
{

constexpr int kTX = 32;  // chunks (16 cells) along y per CTA
constexpr int kTY = 8;   // column quads per CTA
constexpr int kColsPerThread = 4;

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

struct ForestConst
{
    ForestArgs a;
    uint32_t pitch, chunks_per_col, slab_w;
    uint32_t regrow_two_level;
    uint64_t thr_regrow_l2;
    const uint8_t* ov_lo;  // overlay states received from the lower / upper neighbour rank
    const uint8_t* ov_hi;
    const uint8_t* ov_owner; // 0 = this rank, 1 = lower neighbour, 2 = upper neighbour, 3 = elsewhere
};

__device__ __forceinline__ uint8_t ignited_value(const ForestArgs& a)
{
    // Burning{burn_duration} (forest_fire.rs:324-326) then, when progression is
    // enabled, one progression step in the same update (:340-349)
    if (!a.do_progress) return static_cast<uint8_t>(a.burn_duration);
    return a.burn_duration > 1 ? static_cast<uint8_t>(a.burn_duration - 1) : 0;
}

__device__ __forceinline__ bool regrow_draw(const ForestConst& c, uint32_t uid, uint32_t step)
{
    if (c.regrow_two_level)
    {
        const uint4 l1 = philox4x32_10(uid >> 4, 0u, step, 0u, c.a.key_regrow);
        const uint32_t w = word_of(l1, (uid >> 2) & 3u);
        if ((w >> (8 * (uid & 3u))) & 0xffu) return false;
    }
    const uint4 l2 = philox4x32_10(uid >> 2, 0u, step, 1u, c.a.key_regrow);
    return bernoulli(word_of(l2, uid & 3u), c.thr_regrow_l2);
}

// The vectorised body: each thread owns 4 adjacent columns x 16 cells.
template <bool ALIGNED>
__global__ void __launch_bounds__(kTX* kTY) forest_step_kernel(const ForestConst c)
{
    const ForestArgs& a = c.a;
    const uint32_t step = *a.d_step;
    const uint32_t cy = blockIdx.x * kTX + threadIdx.x;
    const uint32_t lx0 = 1 + (blockIdx.y * kTY + threadIdx.y) * kColsPerThread;
    const bool chunk_ok = cy < c.chunks_per_col;
    const uint32_t lane = threadIdx.x; // blockDim.x == 32: one warp per threadIdx.y

    uint32_t n_healthy = 0, n_burning = 0, n_empty = 0;

    // columns lx0-1 .. lx0+4
    uint4 v[kColsPerThread + 2];
#pragma unroll
    for (int j = 0; j < kColsPerThread + 2; ++j)
    {
        const uint32_t lx = lx0 - 1 + j;
        if (chunk_ok && lx <= c.slab_w + 1)
            v[j] = *reinterpret_cast<const uint4*>(a.cur + static_cast<size_t>(lx) * c.pitch + static_cast<size_t>(cy) * 16);
        else
            v[j] = make_uint4(0, 0, 0, 0);
    }

    const uint8_t ign = ignited_value(a);

#pragma unroll
    for (int j = 1; j <= kColsPerThread; ++j)
    {
        const uint32_t lx = lx0 - 1 + j;
        const bool col_ok = chunk_ok && lx <= c.slab_w; // uniform across the warp except chunk_ok
        uint32_t cw[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
        uint32_t B[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) B[k] = burn_bits(cw[k]);

        // burning flags of the y-1 / y+1 cells living in the adjacent chunks
        uint32_t prev = __shfl_up_sync(kFullMask, B[3] >> 24, 1);
        uint32_t next = __shfl_down_sync(kFullMask, B[0] & 1u, 1);
        if (lane == 0)
        {
            prev = 0;
            if (col_ok && cy > 0)
            {
                const uint8_t s = a.cur[static_cast<size_t>(lx) * c.pitch + static_cast<size_t>(cy) * 16 - 1];
                prev = (s & 0x3f) ? 1u : 0u;
            }
        }
        if (lane == 31)
        {
            next = 0;
            if (col_ok && cy + 1 < c.chunks_per_col)
            {
                const uint8_t s = a.cur[static_cast<size_t>(lx) * c.pitch + static_cast<size_t>(cy) * 16 + 16];
                next = (s & 0x3f) ? 1u : 0u;
            }
        }
        if (!col_ok) continue;

        const uint32_t gx = static_cast<uint32_t>(a.x_begin) + lx - 1;
        const uint32_t uid0 = gx * static_cast<uint32_t>(a.H) + cy * 16;

        uint32_t nw[4];
        // ---- progression of cells already burning (:340-349)
#pragma unroll
        for (int k = 0; k < 4; ++k) nw[k] = a.do_progress ? cw[k] - B[k] : cw[k];

        // ---- spread (:282-329)
        if (a.do_spread)
        {
            const uint32_t ww[4] = {v[j - 1].x, v[j - 1].y, v[j - 1].z, v[j - 1].w};
            const uint32_t ew[4] = {v[j + 1].x, v[j + 1].y, v[j + 1].z, v[j + 1].w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
            {
                const uint32_t bn = __funnelshift_l(k == 0 ? (prev << 24) : B[k - 1], B[k], 8);
                const uint32_t bs = __funnelshift_r(B[k], k == 3 ? next : B[k + 1], 8);
                const uint32_t src = bn | (burn_bits(ww[k]) << 1) | (burn_bits(ew[k]) << 2) | (bs << 3);
                const uint32_t healthy = (cw[k] >> 6) & 0x01010101u;
                uint32_t cand = src & (healthy * 0x0fu);
                while (cand)
                {
                    const int bit = __ffs(cand) - 1;
                    const int byte = bit >> 3;
                    const uint32_t dirs = (cand >> (8 * byte)) & 0x0fu;
                    cand &= ~(0xffu << (8 * byte));
                    const uint32_t uid = uid0 + 4 * k + byte;
                    const uint4 r = philox4x32_10(uid, 0u, step, 0u, a.key_spread);
                    const bool hit = ((dirs & 1u) && bernoulli(r.x, a.thr_spread)) ||
                                     ((dirs & 2u) && bernoulli(r.y, a.thr_spread)) ||
                                     ((dirs & 4u) && bernoulli(r.z, a.thr_spread)) ||
                                     ((dirs & 8u) && bernoulli(r.w, a.thr_spread));
                    if (hit) nw[k] = (nw[k] & ~(0xffu << (8 * byte))) | (static_cast<uint32_t>(ign) << (8 * byte));
                }
            }
        }

        // ---- regrowth (:354-365): only EMPTY cells draw
        if (a.do_regrow)
        {
            uint32_t E[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) E[k] = cw[k] & 0x80808080u;
            if (E[0] | E[1] | E[2] | E[3])
            {
                if (ALIGNED && c.regrow_two_level)
                {
                    // uid0 % 16 == 0: one level-1 call covers the 16 cells of this chunk
                    const uint4 l1 = philox4x32_10(uid0 >> 4, 0u, step, 0u, a.key_regrow);
                    const uint32_t lw[4] = {l1.x, l1.y, l1.z, l1.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                    {
                        uint32_t pass = zero_bytes(lw[k]) & E[k];
                        if (pass)
                        {
                            const uint4 l2 = philox4x32_10((uid0 >> 2) + k, 0u, step, 1u, a.key_regrow);
                            const uint32_t l2w[4] = {l2.x, l2.y, l2.z, l2.w};
#pragma unroll
                            for (int byte = 0; byte < 4; ++byte)
                                if ((pass >> (8 * byte + 7)) & 1u)
                                    if (bernoulli(l2w[byte], c.thr_regrow_l2)) nw[k] ^= 0xc0u << (8 * byte);
                        }
                    }
                }
                else
                {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                    {
                        uint32_t e = E[k];
                        while (e)
                        {
                            const int byte = (__ffs(e) - 1) >> 3;
                            e &= ~(0xffu << (8 * byte));
                            if (regrow_draw(c, uid0 + 4 * k + byte, step)) nw[k] ^= 0xc0u << (8 * byte);
                        }
                    }
                }
            }
        }

#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            n_healthy += __popc(nw[k] & 0x40404040u);
            n_empty += __popc(nw[k] & 0x80808080u);
            n_burning += __popc(burn_bits(nw[k]));
        }
        *reinterpret_cast<uint4*>(a.next + static_cast<size_t>(lx) * c.pitch + static_cast<size_t>(cy) * 16) =
            make_uint4(nw[0], nw[1], nw[2], nw[3]);
    }

    // ---- fused FireStats: block partials, combined by the last block
    uint32_t sv[3] = {n_healthy, n_burning, n_empty};
    __shared__ uint32_t s_red[3 * 32];
    {
        const int warp = threadIdx.y;
#pragma unroll
        for (int i = 0; i < 3; ++i) sv[i] = warp_sum(sv[i]);
        if (lane == 0)
        {
#pragma unroll
            for (int i = 0; i < 3; ++i) s_red[i * 32 + warp] = sv[i];
        }
    }
    __syncthreads();
    const uint32_t bid = blockIdx.y * gridDim.x + blockIdx.x;
    const uint32_t nblocks = gridDim.x * gridDim.y;
    if (threadIdx.y == 0 && lane < 3)
    {
        uint32_t t = 0;
        for (int w = 0; w < kTY; ++w) t += s_red[lane * 32 + w];
        a.stats_partials[static_cast<size_t>(bid) * 4 + lane] = t;
    }
    if (!last_block_arrives(a.ticket, nblocks)) return;

    // ================= last block: combine + overlay entities =================
    const uint32_t tid = threadIdx.y * kTX + threadIdx.x;
    uint64_t tot[3] = {0, 0, 0};
    for (uint32_t b = tid; b < nblocks; b += kTX * kTY)
    {
        tot[0] += a.stats_partials[static_cast<size_t>(b) * 4 + 0];
        tot[1] += a.stats_partials[static_cast<size_t>(b) * 4 + 1];
        tot[2] += a.stats_partials[static_cast<size_t>(b) * 4 + 2];
    }
    __shared__ uint64_t s_tot[3 * 32];
    block_sum<uint64_t, 3>(tot, s_tot);
    if (tid != 0) return;

    int64_t healthy = static_cast<int64_t>(tot[0]), burning = static_cast<int64_t>(tot[1]),
            empty = static_cast<int64_t>(tot[2]);
    const int64_t owned = static_cast<int64_t>(c.slab_w) * a.H;
    int64_t burned = owned - healthy - burning - empty;

    // ---- overlay entities (the igniters spawned on top of existing cells,
    // forest_fire.rs:260-278): a scalar restatement of the three systems for the
    // positions they can influence.
    const uint32_t nov = a.n_overlay;
    auto ov_old = [&](uint32_t k) -> uint8_t {
        const uint8_t o = c.ov_owner ? c.ov_owner[k] : 0;
        if (o == 0) return a.ov_cur[k];
        if (o == 1) return c.ov_lo ? c.ov_lo[k] : 0;
        if (o == 2) return c.ov_hi ? c.ov_hi[k] : 0;
        return 0;
    };
    auto in_grid = [&](int x, int y) { return x >= 0 && x < a.W && y >= 0 && y < a.H; };
    auto visible = [&](int x) { return x >= a.x_begin - 1 && x <= a.x_end; }; // own columns + halos
    auto base_ptr = [&](const uint8_t* buf, int x, int y) -> const uint8_t* {
        return buf + static_cast<size_t>(x - a.x_begin + 1) * c.pitch + y;
    };
    const int DX[4] = {0, -1, 1, 0}, DY[4] = {-1, 0, 0, 1};
    auto category_adjust = [&](uint8_t s, int64_t delta) {
        if (s & 0x40)
            healthy += delta;
        else if (s & 0x80)
            empty += delta;
        else if (s & 0x3f)
            burning += delta;
        else
            burned += delta;
    };
    // is position (px,py) marked by fire_spread this step?  (:296-316)
    auto marked = [&](int px, int py) -> bool {
        if (!a.do_spread) return false;
        // healthy entities at p: base cell and overlays
        for (int e = -1; e < static_cast<int>(nov); ++e)
        {
            uint32_t uid_t;
            if (e < 0)
            {
                if (*base_ptr(a.cur, px, py) != 0x40) continue;
                uid_t = static_cast<uint32_t>(px) * a.H + py;
            }
            else
            {
                if (a.ov_pos[2 * e] != px || a.ov_pos[2 * e + 1] != py || ov_old(e) != 0x40) continue;
                uid_t = a.ov_uid_base + e;
            }
            uint4 r = make_uint4(0, 0, 0, 0);
            bool have_r = false;
            for (int d = 0; d < 4; ++d)
            {
                const int qx = px + DX[d], qy = py + DY[d];
                if (!in_grid(qx, qy)) continue;
                if (visible(qx) && ((*base_ptr(a.cur, qx, qy)) & 0x3f) && !((*base_ptr(a.cur, qx, qy)) & 0xc0))
                {
                    if (!have_r)
                    {
                        r = philox4x32_10(uid_t, 0u, step, 0u, a.key_spread);
                        have_r = true;
                    }
                    if (bernoulli(word_of(r, d), a.thr_spread)) return true;
                }
                for (uint32_t o = 0; o < nov; ++o)
                {
                    if (a.ov_pos[2 * o] != qx || a.ov_pos[2 * o + 1] != qy) continue;
                    const uint8_t so = ov_old(o);
                    if (!(so & 0x3f) || (so & 0xc0)) continue;
                    const uint4 rp = philox4x32_10(uid_t, a.ov_uid_base + o, step, 1u, a.key_spread);
                    if (bernoulli(rp.x, a.thr_spread)) return true;
                }
            }
        }
        return false;
    };

    if (nov)
    {
        // overlays first take their own progression / regrowth; marking overrides below
        for (uint32_t k = 0; k < nov; ++k)
        {
            uint8_t s = ov_old(k);
            if (a.do_progress && (s & 0x3f) && !(s & 0xc0)) s = static_cast<uint8_t>(s - 1);
            if (a.do_regrow && s == 0x80 && regrow_draw(c, a.ov_uid_base + k, step)) s = 0x40;
            a.ov_next[k] = s;
        }
        // affected positions: every overlay position and its 4 orthogonal neighbours
        for (uint32_t k = 0; k < nov; ++k)
        {
            for (int d = -1; d < 4; ++d)
            {
                const int px = a.ov_pos[2 * k] + (d < 0 ? 0 : DX[d]);
                const int py = a.ov_pos[2 * k + 1] + (d < 0 ? 0 : DY[d]);
                if (!in_grid(px, py) || px < a.x_begin || px >= a.x_end) continue;
                // dedupe against positions already handled
                bool seen = false;
                for (uint32_t k2 = 0; k2 <= k && !seen; ++k2)
                    for (int d2 = -1; d2 < 4 && !seen; ++d2)
                    {
                        if (k2 == k && d2 >= d) break;
                        const int qx = a.ov_pos[2 * k2] + (d2 < 0 ? 0 : DX[d2]);
                        const int qy = a.ov_pos[2 * k2 + 1] + (d2 < 0 ? 0 : DY[d2]);
                        seen = (qx == px && qy == py);
                    }
                if (seen) continue;
                if (!marked(px, py)) continue;
                // every entity at a marked position becomes Burning{duration} (:320-328)
                uint8_t* nb = const_cast<uint8_t*>(base_ptr(a.next, px, py));
                category_adjust(*nb, -1);
                *nb = ign;
                category_adjust(ign, +1);
                for (uint32_t o = 0; o < nov; ++o)
                    if (a.ov_pos[2 * o] == px && a.ov_pos[2 * o + 1] == py) a.ov_next[o] = ign;
            }
        }
        // overlay entities this rank owns are counted here
        for (uint32_t k = 0; k < nov; ++k)
        {
            const uint8_t o = c.ov_owner ? c.ov_owner[k] : 0;
            if (o == 0) category_adjust(a.ov_next[k], +1);
        }
    }

    uint64_t total = static_cast<uint64_t>(owned);
    for (uint32_t k = 0; k < nov; ++k)
        if ((c.ov_owner ? c.ov_owner[k] : 0) == 0) ++total;

    a.stats_out[0] = static_cast<uint64_t>(healthy);
    a.stats_out[1] = static_cast<uint64_t>(burning);
    a.stats_out[2] = static_cast<uint64_t>(burned);
    a.stats_out[3] = static_cast<uint64_t>(empty);
    a.stats_out[4] = total;
    if (a.series_values && a.series_interval && (step % a.series_interval) == 0)
    {
        // slot = number of sampling instants before this one (time_series.rs:68-87)
        const uint64_t slot = step / a.series_interval - 1;
        if (slot < a.series_capacity)
        {
            uint64_t* rec = reinterpret_cast<uint64_t*>(a.series_values) + slot * 5;
            rec[0] = static_cast<uint64_t>(healthy);
            rec[1] = static_cast<uint64_t>(burning);
            rec[2] = static_cast<uint64_t>(burned);
            rec[3] = static_cast<uint64_t>(empty);
            rec[4] = total;
            a.series_counts[slot] = total;
        }
    }
    // StepNumber += 1 (src/plugins/step_number.rs:20-23): the device-side counter
    // lets a captured step program replay without new arguments
    if (a.bump_step) *const_cast<uint32_t*>(a.d_step) = step + 1;
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
void launch_forest_step_ex(const ForestArgs& a, const uint8_t* ov_lo, const uint8_t* ov_hi, const uint8_t* ov_owner,
                           cudaStream_t st)
{
    ForestConst c;
    c.a = a;
    c.slab_w = static_cast<uint32_t>(a.x_end - a.x_begin);
    c.pitch = (static_cast<uint32_t>(a.H) + 15u) & ~15u;
    c.chunks_per_col = c.pitch / 16;
    c.regrow_two_level = a.regrow_two_level;
    c.thr_regrow_l2 = a.regrow_two_level ? a.thr_regrow_l2 : a.thr_regrow;
    c.ov_lo = ov_lo;
    c.ov_hi = ov_hi;
    c.ov_owner = ov_owner;
    dim3 block(kTX, kTY);
    dim3 grid((c.chunks_per_col + kTX - 1) / kTX, (c.slab_w + kTY * kColsPerThread - 1) / (kTY * kColsPerThread));
    if ((a.H % 16) == 0)
        forest_step_kernel<true><<<grid, block, 0, st>>>(c);
    else
        forest_step_kernel<false><<<grid, block, 0, st>>>(c);
}

void launch_forest_step(const ForestArgs& a, cudaStream_t st) { launch_forest_step_ex(a, nullptr, nullptr, nullptr, st); }

int forest_step_grid_blocks(int32_t slab_w, int32_t H)
{
    const uint32_t chunks = ((static_cast<uint32_t>(H) + 15u) & ~15u) / 16;
    return static_cast<int>(((chunks + kTX - 1) / kTX) *
                            ((static_cast<uint32_t>(slab_w) + kTY * kColsPerThread - 1) / (kTY * kColsPerThread)));
}

void launch_forest_spawn(uint8_t* cells, int32_t W, int32_t H, int32_t x_begin, int32_t x_end, PhiloxKey key,
                         uint64_t thr_density, cudaStream_t st)
{
    (void)W;
    const uint32_t pitch = (static_cast<uint32_t>(H) + 15u) & ~15u;
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
    const uint32_t pitch = (static_cast<uint32_t>(H) + 15u) & ~15u;
    forest_stats_kernel<<<static_cast<int>(b), 256, 0, st>>>(slab, static_cast<uint32_t>(H), pitch,
                                                             static_cast<uint32_t>(slab_w), out5,
                                                             static_cast<uint64_t*>(scratch), ticket);
}

} // namespace incerto
