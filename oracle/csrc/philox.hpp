// ORACLE — test infrastructure, not product code.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may build or call anything in oracle/.
//
// Independent CPU statement of the draw contract.  The reference draws from
// rand::rng() (thread-local, OS-seeded; e.g. examples/coin_toss.rs:68, forest_fire.rs:236) and
// none of its tests exercises an RNG-dependent path, so this contract — not rand's own
// mappings — is what "identical draw stream" means for parity (parity unpinned
// by the reference for every RNG-dependent result).
//
// Philox4x32-10 follows Salmon et al., "Parallel Random Numbers: As Easy as 1, 2, 3" (SC'11);
// the known-answer vectors of the Random123 distribution are checked in tests/test_philox.py.
#pragma once
#include <cstdint>

namespace oracle
{

struct U4
{
    uint32_t w[4];
};

inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo)
{
    const uint64_t p = static_cast<uint64_t>(a) * static_cast<uint64_t>(b);
    hi = static_cast<uint32_t>(p >> 32);
    lo = static_cast<uint32_t>(p & 0xffffffffu);
}

inline U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
    for (int round = 0; round < 10; ++round)
    {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo(0xD2511F53u, c0, hi0, lo0);
        mulhilo(0xCD9E8D57u, c2, hi1, lo1);
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n1 = lo1;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        const uint32_t n3 = lo0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return U4{{c0, c1, c2, c3}};
}

// stream tags (one per draw site), ASCII of the four-letter names
enum : uint32_t
{
    T_COIN = 0x434F494Eu,
    T_FSPW = 0x46535057u,
    T_FIGN = 0x4649474Eu,
    T_FSPR = 0x46535052u,
    T_FREG = 0x46524547u,
    T_PSPW = 0x50535057u,
    T_PMOV = 0x504D4F56u,
    T_PINF = 0x50494E46u,
    T_PFAT = 0x50464154u,
    T_SSPW = 0x53535057u,
    T_SMOV = 0x534D4F56u,
    T_STRN = 0x5354524Eu,
    T_SFAT = 0x53464154u,
    T_TBUY = 0x54425559u,
    T_TSEL = 0x5453454Cu,
    T_TPRC = 0x54505243u,
    T_ASPW = 0x41535057u,
    T_AMOV = 0x414D4F56u,
};

struct Stream
{
    uint32_t k0, k1;
    Stream(uint64_t seed, uint32_t replica, uint32_t tag)
        : k0(static_cast<uint32_t>(seed & 0xffffffffu) ^ replica), k1(static_cast<uint32_t>(seed >> 32) ^ tag)
    {
    }
    U4 block(uint32_t c0, uint32_t c1, uint32_t step, uint32_t c3) const
    {
        return philox4x32_10(c0, c1, step, c3, k0, k1);
    }
};

// bernoulli(p) := u < floor(p * 2^32)
inline uint64_t threshold(double p)
{
    if (!(p > 0.0)) return 0;
    if (p >= 1.0) return 4294967296ull;
    return static_cast<uint64_t>(p * 4294967296.0);
}
inline bool bern(uint32_t u, uint64_t thr) { return static_cast<uint64_t>(u) < thr; }
// uniform integer in [0, n)
inline uint32_t below(uint32_t u, uint32_t n) { return static_cast<uint32_t>((static_cast<uint64_t>(u) * n) >> 32); }

} // namespace oracle
