// Philox4x32-10 counter-based RNG and the draw contract of DESIGN.md §2.
//
// The reference draws from rand::rng() (OS-seeded ChaCha; e.g.
// examples/coin_toss.rs:68-73) which is not reproducible.  The engine replaces
// every draw site by a keyed Philox draw so that results do not depend on the
// launch configuration, the iteration order or the GPU count.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace incerto
{

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

// Stream tags: one per draw site, xor-ed into the high key word.
enum StreamTag : uint32_t
{
    TAG_COIN_TOSS = 0x434F494Eu,      // 'COIN'
    TAG_FOREST_SPAWN = 0x46535057u,   // 'FSPW'
    TAG_FOREST_IGNITE = 0x4649474Eu,  // 'FIGN'
    TAG_FOREST_SPREAD = 0x46535052u,  // 'FSPR'
    TAG_FOREST_REGROW = 0x46524547u,  // 'FREG'
    TAG_PANDEMIC_SPAWN = 0x50535057u, // 'PSPW'
    TAG_PANDEMIC_MOVE = 0x504D4F56u,  // 'PMOV'
    TAG_PANDEMIC_INFECT = 0x50494E46u, // 'PINF'
    TAG_PANDEMIC_FATE = 0x50464154u,  // 'PFAT'  (die: word 0, recover: word 1)
    TAG_SPATIAL_SPAWN = 0x53535057u,  // 'SSPW'
    TAG_SPATIAL_MOVE = 0x534D4F56u,   // 'SMOV'
    TAG_SPATIAL_TRANSMIT = 0x5354524Eu, // 'STRN'
    TAG_SPATIAL_FATE = 0x53464154u,   // 'SFAT'
    TAG_TRADERS_BUY = 0x54425559u,    // 'TBUY'
    TAG_TRADERS_SELL = 0x5453454Cu,   // 'TSEL'
    TAG_STOCKS_PRICE = 0x54505243u,   // 'TPRC'
    TAG_AIR_SPAWN = 0x41535057u,      // 'ASPW'
    TAG_AIR_MOVE = 0x414D4F56u,       // 'AMOV'
};

struct PhiloxKey
{
    uint32_t k0, k1;
};

// key = (seed_lo ^ replica, seed_hi ^ tag)
__host__ __device__ __forceinline__ PhiloxKey make_key(uint64_t seed, uint32_t replica, uint32_t tag)
{
    PhiloxKey k;
    k.k0 = static_cast<uint32_t>(seed) ^ replica;
    k.k1 = static_cast<uint32_t>(seed >> 32) ^ tag;
    return k;
}

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, PhiloxKey key)
{
    uint32_t k0 = key.k0, k1 = key.k1;
#pragma unroll
    for (int r = 0; r < 10; ++r)
    {
        const uint64_t p0 = static_cast<uint64_t>(kPhiloxM0) * c0;
        const uint64_t p1 = static_cast<uint64_t>(kPhiloxM1) * c2;
        const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
        c1 = static_cast<uint32_t>(p1);
        c3 = static_cast<uint32_t>(p0);
        c0 = n0;
        c2 = n2;
        k0 += kPhiloxW0;
        k1 += kPhiloxW1;
    }
    return make_uint4(c0, c1, c2, c3);
}

// The ten round keys of a stream, precomputed on the host.  Kernels keep them in their parameter
// block (constant bank), which turns the key schedule into constant operands of the xor instructions.
struct PhiloxRoundKeys
{
    uint32_t k[20];
};
__host__ __device__ __forceinline__ PhiloxRoundKeys make_round_keys(PhiloxKey key)
{
    PhiloxRoundKeys r;
    uint32_t k0 = key.k0, k1 = key.k1;
    for (int i = 0; i < 10; ++i)
    {
        r.k[2 * i] = k0;
        r.k[2 * i + 1] = k1;
        k0 += kPhiloxW0;
        k1 += kPhiloxW1;
    }
    return r;
}
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxRoundKeys& rk)
{
#pragma unroll
    for (int r = 0; r < 10; ++r)
    {
        const uint64_t p0 = static_cast<uint64_t>(kPhiloxM0) * c0;
        const uint64_t p1 = static_cast<uint64_t>(kPhiloxM1) * c2;
        const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ rk.k[2 * r];
        const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ rk.k[2 * r + 1];
        c1 = static_cast<uint32_t>(p1);
        c3 = static_cast<uint32_t>(p0);
        c0 = n0;
        c2 = n2;
    }
    return make_uint4(c0, c1, c2, c3);
}

// bernoulli(p): u < floor(p * 2^32), threshold precomputed on the host as u64
// (p >= 1 -> 2^32 -> always true; p <= 0 -> 0 -> never).
__host__ __device__ __forceinline__ bool bernoulli(uint32_t u, uint64_t threshold)
{
    return static_cast<uint64_t>(u) < threshold;
}

// uniform integer in [0, n): mulhi32(u, n)
__host__ __device__ __forceinline__ uint32_t uniform_below(uint32_t u, uint32_t n)
{
    return static_cast<uint32_t>((static_cast<uint64_t>(u) * n) >> 32);
}

__host__ __device__ __forceinline__ uint32_t word_of(const uint4& v, uint32_t i)
{
    return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

} // namespace incerto
