// ORACLE — test infrastructure, not product code.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may build or call anything in oracle/.
//
// Block lottery: the 2^-8 factor of a rare per-entity Bernoulli draw (the reference's
// `rng.random_bool(p)`, e.g. forest_fire.rs:360) sampled for the 256 ids of block B = id >> 8 with one call:
// the number of winners K ~ Binomial(256, 2^-8) by inverse CDF on a 64-bit uniform, then K distinct byte positions
// taken in order from the block's byte stream.  Independent CPU statement; the thresholds are recomputed with exact
// rationals by tests/test_lottery.py.
#pragma once
#include <cstdint>
#include <vector>

#include "philox.hpp"

namespace oracle
{

// floor(P(Binomial(256, 1/256) <= k) * 2^64), k = 0..20
static const uint64_t LOTTERY_CDF[21] = {
    0x5dfe2e83aaa155dfull, 0xbc5ab99263fc066eull, 0xeb88ff19c0a95eb6ull, 0xfb334c65d160bff4ull, 0xff1602b52bdc8a86ull,
    0xffda9ccd49cad7e7ull, 0xfffadd91a9b98cecull, 0xffff61fa9ca0e619ull, 0xffffef2105923822ull, 0xfffffe61be083611ull,
    0xffffffdbf6da03dfull, 0xfffffffd2270a8adull, 0xffffffffca5278cdull, 0xfffffffffc5d60d0ull, 0xffffffffffc5618bull,
    0xfffffffffffc8d07ull, 0xffffffffffffcf48ull, 0xfffffffffffffd78ull, 0xffffffffffffffe0ull, 0xfffffffffffffffeull,
    0xffffffffffffffffull};

// winners (positions 0..255) of block `block` at `step`, in the order they are drawn
inline std::vector<uint8_t> lottery256(const Stream& rng, uint32_t block, uint32_t step)
{
    const U4 first = rng.block(block, 0, step, 0);
    const uint64_t u = (static_cast<uint64_t>(first.w[0]) << 32) | first.w[1];
    int k = 0;
    while (k < 21 && u >= LOTTERY_CDF[k]) ++k;
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
            cur = rng.block(block, 0, step, 2 + extra++);
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
    uint32_t block = 0xffffffffu, step = 0xffffffffu;
    bool win[256] = {};
    bool wins(const Stream& rng, uint32_t id, uint32_t at_step)
    {
        if ((id >> 8) != block || at_step != step)
        {
            block = id >> 8;
            step = at_step;
            for (bool& w : win) w = false;
            for (uint8_t b : lottery256(rng, block, step)) win[b] = true;
        }
        return win[id & 255u];
    }
};

} // namespace oracle
