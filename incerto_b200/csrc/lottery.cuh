// Block lottery: the level-1 draw of a rare per-entity Bernoulli event, sampled for 256 consecutive ids at once
// (DESIGN.md §2.4).  The reference asks every candidate `rng.random_bool(p)` (forest_fire.rs:360, pandemic.rs:201,218,
// traders.rs:178,214): independent Bernoulli draws.  For p <= 2^-k (k = 5..8) the contract factors p = 2^-k * (2^k p)
// and realises the 2^-k factor for a whole block of 256 ids with ONE Philox call: the ids that pass ("winners") are
// exactly distributed like 256 independent Bernoulli(2^-k) draws, sampled as
//   K ~ Binomial(256, 2^-k) by inverse CDF on a 64-bit uniform (thresholds floor(CDF(j) 2^64), j = 0, 1, ...), then
//   K distinct positions: bytes of the stream below, in order, skipping bytes already taken.
// Stream of block B, sub-stream s, at `step`: r = philox(key, (B, s, step, 0)):  U = r.x << 32 | r.y;  bytes 0..3 = r.z
// (least significant first), 4..7 = r.w; bytes 8 + 16 c + i (c = 0, 1, ...) = byte i of the words x, y, z, w of
// philox(key, (B, s, step, 2 + c)).  (Counter word 3 = 1 is the level-2 draw of the entities.)
#pragma once
#include "philox.cuh"

namespace incerto
{

// floor(P(Binomial(256, 2^-k) <= j) * 2^64), j = 0 .. until 2^64 - 1 is reached; row k - 5.  Exact rational arithmetic:
// tools/gen_lottery_tables.py; tests/test_lottery.py recomputes them.
constexpr int kLotteryRows = 45;
static __device__ __constant__ const uint64_t kLotteryCdf[4][kLotteryRows] = {
    // k = 5: 45 thresholds
    {
        0x0013599425f6a185ull, 0x00b3248d1d673234ull, 0x03445af871eff8e6ull, 0x0a4752b23a89515full, 0x1895a2d8a456d667ull,
        0x2fd7d92e379474e1ull, 0x4f3ad56d65806222ull, 0x7363bab04901711cull, 0x97b1f37479d070a9ull, 0xb7f77b3f4fddc599ull,
        0xd1ae0e418bc2683dull, 0xe43ac662b035cc4eull, 0xf0724aaa10798964ull, 0xf7d7dffe91e7ca53ull, 0xfbfc1b6bd21ddaa0ull,
        0xfe23e2433ba78c20ull, 0xff2ffc2adde6d680ull, 0xffaa14983da664b2ull, 0xffde603447012806ull, 0xfff381cc6d8e24ceull,
        0xfffb95a8dc15cacdull, 0xfffe834c87e3dedeull, 0xffff859b130af94dull, 0xffffda6131a1ae51ull, 0xfffff4edb0e37575ull,
        0xfffffce04158ecdcull, 0xffffff275cc62f29ull, 0xffffffc7984e56caull, 0xfffffff1de477ebdull, 0xfffffffc96e539b5ull,
        0xffffffff34d15e42ull, 0xffffffffd25d5900ull, 0xfffffffff61939e0ull, 0xfffffffffdec47d4ull, 0xffffffffff9413a5ull,
        0xffffffffffeac9f9ull, 0xfffffffffffbf5e1ull, 0xffffffffffff4108ull, 0xffffffffffffddc8ull, 0xfffffffffffffa0cull,
        0xfffffffffffffefeull, 0xffffffffffffffd5ull, 0xfffffffffffffff9ull, 0xfffffffffffffffeull, 0xffffffffffffffffull},
    // k = 6: 33 thresholds
    {
        0x048b052a991c311cull, 0x1700f1866cdbf49cull, 0x3c5d4fdebdf86dfaull, 0x6e93037cb67d82dcull, 0xa0fbb8d5371b8076ull,
        0xc94f7d1c0466b1beull, 0xe416b03401351174ull, 0xf344db198c753c19ull, 0xfac4cec4dbebcb13ull, 0xfe0c93321a3807cdull,
        0xff55d12c2bbc5200ull, 0xffcab0deb4b96903ull, 0xfff091117414e95cull, 0xfffbd9cd1d4e7df0ull, 0xfffef5ad39a4e2f4ull,
        0xffffc17cf7375a8eull, 0xfffff23785cf7398ull, 0xfffffd22f10eef86ull, 0xffffff70197ef7cfull, 0xffffffe53dfb0b6cull,
        0xfffffffb46abfb9bull, 0xffffffff34df4850ull, 0xffffffffdf79e592ull, 0xfffffffffb06f0a8ull, 0xffffffffff45d159ull,
        0xffffffffffe5eaa5ull, 0xfffffffffffc7e9full, 0xffffffffffff8c29ull, 0xfffffffffffff19eull, 0xfffffffffffffe48ull,
        0xffffffffffffffcdull, 0xfffffffffffffffaull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull,
        0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull,
        0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull},
    // k = 7: 26 thresholds
    {
        0x225ff3841d92c634ull, 0x67aa6f84592f8c28ull, 0xad3ac1ace5a9a991ull, 0xdb9af87298a5bd2cull, 0xf2b3b4fc5d2041e6ull,
        0xfbde2dc42bb40dc1ull, 0xfee3191418bc6c1dull, 0xffbc744dfe17c1afull, 0xfff1b946c3a8586bull, 0xfffd482122eee98eull,
        0xffff8797663e8d3eull, 0xffffecece96494cfull, 0xfffffd374c963397ull, 0xffffff9fa33175b2ull, 0xfffffff3df685ee9ull,
        0xfffffffe92cc7950ull, 0xffffffffd7b26cd2ull, 0xfffffffffbd04175ull, 0xffffffffff96e6ebull, 0xfffffffffff63ea4ull,
        0xffffffffffff240dull, 0xffffffffffffed94ull, 0xfffffffffffffe87ull, 0xffffffffffffffe3ull, 0xfffffffffffffffdull,
        0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull,
        0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull,
        0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull,
        0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull},
    // k = 8: 21 thresholds
    {
        0x5dfe2e83aaa155dfull, 0xbc5ab99263fc066eull, 0xeb88ff19c0a95eb6ull, 0xfb334c65d160bff4ull, 0xff1602b52bdc8a86ull,
        0xffda9ccd49cad7e7ull, 0xfffadd91a9b98cecull, 0xffff61fa9ca0e619ull, 0xffffef2105923822ull, 0xfffffe61be083611ull,
        0xffffffdbf6da03dfull, 0xfffffffd2270a8adull, 0xffffffffca5278cdull, 0xfffffffffc5d60d0ull, 0xffffffffffc5618bull,
        0xfffffffffffc8d07ull, 0xffffffffffffcf48ull, 0xfffffffffffffd78ull, 0xffffffffffffffe0ull, 0xfffffffffffffffeull,
        0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull,
        0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull,
        0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull,
        0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull,
        0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull}
};
constexpr int kLotteryMaxWinners = 21; // k = 8 (the longest list a caller of lottery256 must hold)

// number of winners: thresholds <= u.  K8: the k = 8 table's first four entries as immediates (P(K >= 4) = 1.9 %).
__device__ __forceinline__ uint32_t lottery_count(uint64_t u)
{
    uint32_t k = (u >= 0x5dfe2e83aaa155dfull ? 1u : 0u) + (u >= 0xbc5ab99263fc066eull ? 1u : 0u) +
                 (u >= 0xeb88ff19c0a95eb6ull ? 1u : 0u) + (u >= 0xfb334c65d160bff4ull ? 1u : 0u);
    if (k == 4)
        while (k < 21 && u >= kLotteryCdf[3][k]) ++k;
    return k;
}
// any k in 5..8 (table walk; the thresholds of a row end with 2^64 - 1, which no u exceeds... but may equal)
__device__ __forceinline__ uint32_t lottery_count_k(uint64_t u, uint32_t kexp)
{
    uint32_t k = 0;
    while (k < kLotteryRows && u >= kLotteryCdf[kexp - 5][k]) ++k;
    return k;
}

// byte i of the position stream of a block whose first call returned r
__device__ __forceinline__ uint32_t lottery_byte(const PhiloxRoundKeys& rk, uint32_t block, uint32_t sub, uint32_t step, const uint4& r, uint32_t i)
{
    if (i < 8) return ((i < 4 ? r.z : r.w) >> (8 * (i & 3u))) & 0xffu;
    const uint32_t j = i - 8;
    const uint4 e = philox4x32_10(block, sub, step, 2u + (j >> 4), rk); // rare: recomputed per byte
    return (word_of(e, (j >> 2) & 3u) >> (8 * (j & 3u))) & 0xffu;
}
// the same from the two words that hold the first eight bytes
__device__ __forceinline__ uint32_t lottery_stream_byte(const PhiloxRoundKeys& rk, uint32_t block, uint32_t sub, uint32_t step, uint32_t z, uint32_t w, uint32_t i)
{
    uint4 r;
    r.z = z;
    r.w = w;
    return lottery_byte(rk, block, sub, step, r, i);
}

// Walks the position stream of a block whose first call gave (count, z, w): calls emit(position) for every winner.
// The first eight bytes are compared byte-parallel in registers; no scratch memory.
template <typename Emit>
__device__ __forceinline__ void lottery_walk(const PhiloxRoundKeys& rk, uint32_t block, uint32_t sub, uint32_t step, uint32_t count, uint32_t z,
                                             uint32_t w, Emit&& emit)
{
    auto zero_bytes = [](uint32_t v) { return ~(((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v) & 0x80808080u; };
    uint32_t need = count; // stream bytes still to look at: winners left + repeats met
    uint4 e = make_uint4(0u, 0u, 0u, 0u); // bytes 8..23: the first continuation call, made once when the walk gets there
    for (uint32_t t = 0; t < need; ++t)
    {
        uint32_t b;
        bool rep;
        if (t < 8)
        {
            b = ((t < 4 ? z : w) >> (8 * (t & 3u))) & 0xffu;
            const uint32_t r4 = b * 0x01010101u;
            const uint32_t mz = t >= 4 ? 0xffffffffu : ((1u << (8 * t)) - 1u), mw = t > 4 ? ((1u << (8 * (t - 4))) - 1u) : 0u;
            rep = ((zero_bytes(z ^ r4) & mz) | (zero_bytes(w ^ r4) & mw)) != 0;
        }
        else if (t < 24)
        {
            if (t == 8) e = philox4x32_10(block, sub, step, 2u, rk);
            const uint32_t j = t - 8, ws = j >> 2, part = (1u << (8 * (j & 3u))) - 1u; // word of the byte, earlier bytes of that word
            b = (word_of(e, ws) >> (8 * (j & 3u))) & 0xffu;
            const uint32_t r4 = b * 0x01010101u;
            uint32_t hit = zero_bytes(z ^ r4) | zero_bytes(w ^ r4);
            hit |= zero_bytes(e.x ^ r4) & (ws > 0 ? 0xffffffffu : part);
            hit |= zero_bytes(e.y ^ r4) & (ws > 1 ? 0xffffffffu : (ws == 1 ? part : 0u));
            hit |= zero_bytes(e.z ^ r4) & (ws > 2 ? 0xffffffffu : (ws == 2 ? part : 0u));
            hit |= zero_bytes(e.w ^ r4) & (ws == 3 ? part : 0u);
            rep = hit != 0;
        }
        else // 24 stream bytes used up: out of reach in practice
        {
            b = lottery_stream_byte(rk, block, sub, step, z, w, t);
            rep = false;
            for (uint32_t h = 0; h < t; ++h) rep = rep || lottery_stream_byte(rk, block, sub, step, z, w, h) == b;
        }
        if (rep)
            ++need;
        else
            emit(b);
    }
}

// one block, k = 8: emit(position) for every winner
template <typename Emit>
__device__ __forceinline__ void lottery256(const PhiloxRoundKeys& rk, uint32_t block, uint32_t sub, uint32_t step, Emit&& emit)
{
    const uint4 r = philox4x32_10(block, sub, step, 0u, rk);
    lottery_walk(rk, block, sub, step, lottery_count((static_cast<uint64_t>(r.x) << 32) | r.y), r.z, r.w, emit);
}
// any k in 5..8
template <typename Emit>
__device__ __forceinline__ void lottery256_k(const PhiloxRoundKeys& rk, uint32_t kexp, uint32_t block, uint32_t sub, uint32_t step, Emit&& emit)
{
    const uint4 r = philox4x32_10(block, sub, step, 0u, rk);
    lottery_walk(rk, block, sub, step, lottery_count_k((static_cast<uint64_t>(r.x) << 32) | r.y, kexp), r.z, r.w, emit);
}

// scalar form: did id `uid` win the k = 8 lottery of its block?
__device__ inline bool lottery256_wins(const PhiloxRoundKeys& rk, uint32_t uid, uint32_t sub, uint32_t step)
{
    bool won = false;
    const uint32_t pos = uid & 255u;
    lottery256(rk, uid >> 8, sub, step, [&](uint32_t b) { won = won || b == pos; });
    return won;
}

} // namespace incerto
