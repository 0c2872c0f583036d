// Block lottery: the level-1 draw of a rare per-entity Bernoulli event, sampled for 256 consecutive ids at once
// (DESIGN.md §2.4).  The reference asks every Empty cell `rng.random_bool(p)` (forest_fire.rs:354-365): independent
// Bernoulli draws.  For p <= 1/256 the contract factors p = 2^-8 * (256 p) and realises the 2^-8 factor for a whole
// block of 256 ids with ONE Philox call: the ids that pass ("winners") are exactly distributed like 256 independent
// Bernoulli(2^-8) draws, sampled as
//   K ~ Binomial(256, 2^-8) by inverse CDF on a 64-bit uniform (thresholds floor(CDF(k) 2^64), k = 0..20), then
//   K distinct positions: bytes of the stream below, in order, skipping bytes already taken.
// Stream of block B at `step`: r = philox(key, (B, 0, step, 0)):  U = r.x << 32 | r.y;  bytes 0..3 = r.z (least
// significant first), 4..7 = r.w; bytes 8 + 16 c + i (c = 0, 1, ...) = byte i of the words x, y, z, w of
// philox(key, (B, 0, step, 2 + c)).  (Counter word 3 = 1 is the level-2 draw of the entities.)
#pragma once
#include "philox.cuh"

namespace incerto
{

// floor(P(Binomial(256, 1/256) <= k) * 2^64), k = 0..20 (exact rational arithmetic; tests/test_lottery.py recomputes them)
static __device__ __constant__ const uint64_t kLotteryCdf[21] = {
    0x5dfe2e83aaa155dfull, 0xbc5ab99263fc066eull, 0xeb88ff19c0a95eb6ull, 0xfb334c65d160bff4ull, 0xff1602b52bdc8a86ull,
    0xffda9ccd49cad7e7ull, 0xfffadd91a9b98cecull, 0xffff61fa9ca0e619ull, 0xffffef2105923822ull, 0xfffffe61be083611ull,
    0xffffffdbf6da03dfull, 0xfffffffd2270a8adull, 0xffffffffca5278cdull, 0xfffffffffc5d60d0ull, 0xffffffffffc5618bull,
    0xfffffffffffc8d07ull, 0xffffffffffffcf48ull, 0xfffffffffffffd78ull, 0xffffffffffffffe0ull, 0xfffffffffffffffeull,
    0xffffffffffffffffull};
constexpr int kLotteryMaxWinners = 21;

__device__ __forceinline__ uint32_t lottery_count(uint64_t u)
{
    // number of thresholds <= u; the first four as immediates (P(K >= 4) = 1.9 %)
    uint32_t k = (u >= 0x5dfe2e83aaa155dfull ? 1u : 0u) + (u >= 0xbc5ab99263fc066eull ? 1u : 0u) +
                 (u >= 0xeb88ff19c0a95eb6ull ? 1u : 0u) + (u >= 0xfb334c65d160bff4ull ? 1u : 0u);
    if (k == 4)
        while (k < kLotteryMaxWinners && u >= kLotteryCdf[k]) ++k;
    return k;
}

// byte i of the position stream of a block whose first call returned r
__device__ __forceinline__ uint32_t lottery_byte(const PhiloxRoundKeys& rk, uint32_t block, uint32_t step, const uint4& r, uint32_t i)
{
    if (i < 8) return ((i < 4 ? r.z : r.w) >> (8 * (i & 3u))) & 0xffu;
    const uint32_t j = i - 8;
    const uint4 e = philox4x32_10(block, 0u, step, 2u + (j >> 4), rk); // rare (K >= 8 or repeated bytes): recomputed per byte
    return (word_of(e, (j >> 2) & 3u) >> (8 * (j & 3u))) & 0xffu;
}

// the same from the two words that hold the first eight bytes
__device__ __forceinline__ uint32_t lottery_stream_byte(const PhiloxRoundKeys& rk, uint32_t block, uint32_t step, uint32_t z, uint32_t w, uint32_t i)
{
    uint4 r;
    r.z = z;
    r.w = w;
    return lottery_byte(rk, block, step, r, i);
}

// Calls emit(position) for every winner of the block.  `taken` is scratch for the positions already chosen
// (>= kLotteryMaxWinners bytes, private to the calling thread; shared or local memory).
template <typename Emit>
__device__ __forceinline__ void lottery256(const PhiloxRoundKeys& rk, uint32_t block, uint32_t step, uint8_t* taken, Emit&& emit)
{
    const uint4 r = philox4x32_10(block, 0u, step, 0u, rk);
    const uint32_t k = lottery_count((static_cast<uint64_t>(r.x) << 32) | r.y);
    uint32_t n = 0;
    for (uint32_t i = 0; n < k; ++i)
    {
        const uint32_t b = lottery_byte(rk, block, step, r, i);
        bool dup = false;
        for (uint32_t t = 0; t < n; ++t) dup = dup || taken[t] == b;
        if (dup) continue;
        taken[n++] = static_cast<uint8_t>(b);
        emit(b);
    }
}

// scalar form: did id `uid` win the lottery of its block?
__device__ inline bool lottery256_wins(const PhiloxRoundKeys& rk, uint32_t uid, uint32_t step)
{
    uint8_t taken[kLotteryMaxWinners];
    bool won = false;
    const uint32_t pos = uid & 255u;
    lottery256(rk, uid >> 8, step, taken, [&](uint32_t b) { won = won || b == pos; });
    return won;
}

} // namespace incerto
