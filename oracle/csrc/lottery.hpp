// ORACLE — test infrastructure, not product code.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may build or call anything in oracle/.
//
// Block lottery: the 2^-k factor (k = 5..8) of a rare per-entity Bernoulli draw (the reference's
// `rng.random_bool(p)`, e.g. forest_fire.rs:360, pandemic.rs:201,218, traders.rs:178,214) sampled for the 256 ids of
// block B = id >> 8 with one call: the number of winners K ~ Binomial(256, 2^-k) by inverse CDF on a 64-bit uniform,
// then K distinct byte positions taken in order from the block's byte stream.  Independent CPU statement; the
// thresholds are recomputed with exact rationals by tests/test_lottery.py.
#pragma once
#include <cstdint>
#include <vector>

#include "philox.hpp"

namespace oracle
{

// floor(P(Binomial(256, 2^-k) <= j) * 2^64), j = 0 .. until 2^64 - 1 is reached
static const uint64_t LOTTERY_CDF_K5[45] = {
    0x0013599425f6a185ull, 0x00b3248d1d673234ull, 0x03445af871eff8e6ull, 0x0a4752b23a89515full, 0x1895a2d8a456d667ull,
    0x2fd7d92e379474e1ull, 0x4f3ad56d65806222ull, 0x7363bab04901711cull, 0x97b1f37479d070a9ull, 0xb7f77b3f4fddc599ull,
    0xd1ae0e418bc2683dull, 0xe43ac662b035cc4eull, 0xf0724aaa10798964ull, 0xf7d7dffe91e7ca53ull, 0xfbfc1b6bd21ddaa0ull,
    0xfe23e2433ba78c20ull, 0xff2ffc2adde6d680ull, 0xffaa14983da664b2ull, 0xffde603447012806ull, 0xfff381cc6d8e24ceull,
    0xfffb95a8dc15cacdull, 0xfffe834c87e3dedeull, 0xffff859b130af94dull, 0xffffda6131a1ae51ull, 0xfffff4edb0e37575ull,
    0xfffffce04158ecdcull, 0xffffff275cc62f29ull, 0xffffffc7984e56caull, 0xfffffff1de477ebdull, 0xfffffffc96e539b5ull,
    0xffffffff34d15e42ull, 0xffffffffd25d5900ull, 0xfffffffff61939e0ull, 0xfffffffffdec47d4ull, 0xffffffffff9413a5ull,
    0xffffffffffeac9f9ull, 0xfffffffffffbf5e1ull, 0xffffffffffff4108ull, 0xffffffffffffddc8ull, 0xfffffffffffffa0cull,
    0xfffffffffffffefeull, 0xffffffffffffffd5ull, 0xfffffffffffffff9ull, 0xfffffffffffffffeull, 0xffffffffffffffffull};
static const uint64_t LOTTERY_CDF_K6[33] = {
    0x048b052a991c311cull, 0x1700f1866cdbf49cull, 0x3c5d4fdebdf86dfaull, 0x6e93037cb67d82dcull, 0xa0fbb8d5371b8076ull,
    0xc94f7d1c0466b1beull, 0xe416b03401351174ull, 0xf344db198c753c19ull, 0xfac4cec4dbebcb13ull, 0xfe0c93321a3807cdull,
    0xff55d12c2bbc5200ull, 0xffcab0deb4b96903ull, 0xfff091117414e95cull, 0xfffbd9cd1d4e7df0ull, 0xfffef5ad39a4e2f4ull,
    0xffffc17cf7375a8eull, 0xfffff23785cf7398ull, 0xfffffd22f10eef86ull, 0xffffff70197ef7cfull, 0xffffffe53dfb0b6cull,
    0xfffffffb46abfb9bull, 0xffffffff34df4850ull, 0xffffffffdf79e592ull, 0xfffffffffb06f0a8ull, 0xffffffffff45d159ull,
    0xffffffffffe5eaa5ull, 0xfffffffffffc7e9full, 0xffffffffffff8c29ull, 0xfffffffffffff19eull, 0xfffffffffffffe48ull,
    0xffffffffffffffcdull, 0xfffffffffffffffaull, 0xffffffffffffffffull};
static const uint64_t LOTTERY_CDF_K7[26] = {
    0x225ff3841d92c634ull, 0x67aa6f84592f8c28ull, 0xad3ac1ace5a9a991ull, 0xdb9af87298a5bd2cull, 0xf2b3b4fc5d2041e6ull,
    0xfbde2dc42bb40dc1ull, 0xfee3191418bc6c1dull, 0xffbc744dfe17c1afull, 0xfff1b946c3a8586bull, 0xfffd482122eee98eull,
    0xffff8797663e8d3eull, 0xffffecece96494cfull, 0xfffffd374c963397ull, 0xffffff9fa33175b2ull, 0xfffffff3df685ee9ull,
    0xfffffffe92cc7950ull, 0xffffffffd7b26cd2ull, 0xfffffffffbd04175ull, 0xffffffffff96e6ebull, 0xfffffffffff63ea4ull,
    0xffffffffffff240dull, 0xffffffffffffed94ull, 0xfffffffffffffe87ull, 0xffffffffffffffe3ull, 0xfffffffffffffffdull,
    0xffffffffffffffffull};
static const uint64_t LOTTERY_CDF_K8[21] = {
    0x5dfe2e83aaa155dfull, 0xbc5ab99263fc066eull, 0xeb88ff19c0a95eb6ull, 0xfb334c65d160bff4ull, 0xff1602b52bdc8a86ull,
    0xffda9ccd49cad7e7ull, 0xfffadd91a9b98cecull, 0xffff61fa9ca0e619ull, 0xffffef2105923822ull, 0xfffffe61be083611ull,
    0xffffffdbf6da03dfull, 0xfffffffd2270a8adull, 0xffffffffca5278cdull, 0xfffffffffc5d60d0ull, 0xffffffffffc5618bull,
    0xfffffffffffc8d07ull, 0xffffffffffffcf48ull, 0xfffffffffffffd78ull, 0xffffffffffffffe0ull, 0xfffffffffffffffeull,
    0xffffffffffffffffull};

inline int lottery_count(uint64_t u, int kexp)
{
    const uint64_t* t = kexp == 5 ? LOTTERY_CDF_K5 : kexp == 6 ? LOTTERY_CDF_K6 : kexp == 7 ? LOTTERY_CDF_K7 : LOTTERY_CDF_K8;
    const int n = kexp == 5 ? 45 : kexp == 6 ? 33 : kexp == 7 ? 26 : 21;
    int k = 0;
    while (k < n && u >= t[k]) ++k;
    return k;
}

// winners (positions 0..255) of block `block`, sub-stream `sub`, at `step`, in the order they are drawn
inline std::vector<uint8_t> lottery256(const Stream& rng, uint32_t block, uint32_t sub, uint32_t step, int kexp = 8)
{
    const U4 first = rng.block(block, sub, step, 0);
    const int k = lottery_count((static_cast<uint64_t>(first.w[0]) << 32) | first.w[1], kexp);
    // byte stream: words 2, 3 of the first call, then all four words of the calls with counter word 3 = 2, 3, ...
    std::vector<uint8_t> winners;
    bool seen[256] = {};
    uint32_t extra = 0;
    U4 cur = first;
    int word = 2, byte = 0;
    while (static_cast<int>(winners.size()) < k)
    {
        if (word == 4)
        {
            cur = rng.block(block, sub, step, 2 + extra++);
            word = 0;
        }
        const uint8_t b = static_cast<uint8_t>(cur.w[word] >> (8 * byte));
        if (++byte == 4)
        {
            byte = 0;
            ++word;
        }
        if (seen[b]) continue;
        seen[b] = true;
        winners.push_back(b);
    }
    return winners;
}

// membership test with a one-block cache (systems walk the entities in id order)
struct LotteryCache
{
    uint32_t block = 0xffffffffu, sub = 0xffffffffu, step = 0xffffffffu;
    int kexp = 0;
    bool win[256] = {};
    bool wins(const Stream& rng, uint32_t id, uint32_t at_sub, uint32_t at_step, int at_kexp = 8)
    {
        return wins_at(rng, id >> 8, id & 255u, at_sub, at_step, at_kexp);
    }
    bool wins_at(const Stream& rng, uint32_t at_block, uint32_t pos, uint32_t at_sub, uint32_t at_step, int at_kexp = 8)
    {
        if (at_block != block || at_sub != sub || at_step != step || at_kexp != kexp)
        {
            block = at_block;
            sub = at_sub;
            step = at_step;
            kexp = at_kexp;
            for (bool& w : win) w = false;
            for (uint8_t b : lottery256(rng, block, sub, step, kexp)) win[b] = true;
        }
        return win[pos];
    }
};

} // namespace oracle
