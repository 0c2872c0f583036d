// ORACLE — test infrastructure, not product code (see philox.hpp).
//
// Deterministic f64 transforms of the draw contract ("normal draw").  The reference samples
// rand_distr::Normal (examples/traders.rs:148-153); that crate is not in this environment and no reference
// test pins its stream, so the contract defines its own transform: Box-Muller on 53-bit uniforms, with
// ln / cos evaluated by fixed polynomials made only of IEEE-754 +, -, *, / and sqrt, each rounded once
// (this file is compiled with -ffp-contract=off).  Host and device therefore agree bit for bit, and the
// threshold decisions downstream of a price (delisting at price < EPSILON, traders.rs:157; affordability
// amount > cash, :184) cannot flip on a 1-ulp libm difference.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace oracle
{

// natural logarithm of a positive normal double
inline double det_log(double x)
{
    uint64_t bits;
    std::memcpy(&bits, &x, 8);
    int e = static_cast<int>((bits >> 52) & 0x7ff) - 1023;
    bits = (bits & 0x000fffffffffffffull) | 0x3ff0000000000000ull;
    double m;
    std::memcpy(&m, &bits, 8); // [1, 2)
    if (m > 1.4142135623730951)
    {
        m = m * 0.5;
        e += 1;
    }
    const double f = m - 1.0;
    const double s = f / (2.0 + f);
    const double s2 = s * s;
    // atanh(s) / s = sum s^(2k) / (2k + 1), k = 0..10
    double q = 1.0 / 21.0;
    q = q * s2 + 1.0 / 19.0;
    q = q * s2 + 1.0 / 17.0;
    q = q * s2 + 1.0 / 15.0;
    q = q * s2 + 1.0 / 13.0;
    q = q * s2 + 1.0 / 11.0;
    q = q * s2 + 1.0 / 9.0;
    q = q * s2 + 1.0 / 7.0;
    q = q * s2 + 1.0 / 5.0;
    q = q * s2 + 1.0 / 3.0;
    q = q * s2 + 1.0;
    return static_cast<double>(e) * 0.6931471805599453 + (2.0 * s) * q;
}

// cos(2 pi u) for u in [0, 1)
inline double det_cos2pi(double u)
{
    const double t = u * 4.0;
    const int quadrant = static_cast<int>(t);
    const double r = t - static_cast<double>(quadrant); // [0, 1), exact
    const bool swap = r > 0.5;
    const double rr = swap ? 1.0 - r : r;
    const double th = rr * 1.5707963267948966; // [0, pi/4]
    const double t2 = th * th;
    // sin(th) / th = sum (-1)^k th^(2k) / (2k + 1)!, k = 0..8
    double sp = 1.0 / 355687428096000.0;
    sp = sp * t2 - 1.0 / 1307674368000.0;
    sp = sp * t2 + 1.0 / 6227020800.0;
    sp = sp * t2 - 1.0 / 39916800.0;
    sp = sp * t2 + 1.0 / 362880.0;
    sp = sp * t2 - 1.0 / 5040.0;
    sp = sp * t2 + 1.0 / 120.0;
    sp = sp * t2 - 1.0 / 6.0;
    sp = sp * t2 + 1.0;
    const double sn = th * sp;
    // cos(th) = sum (-1)^k th^(2k) / (2k)!, k = 0..9
    double cp = -1.0 / 6402373705728000.0;
    cp = cp * t2 + 1.0 / 20922789888000.0;
    cp = cp * t2 - 1.0 / 87178291200.0;
    cp = cp * t2 + 1.0 / 479001600.0;
    cp = cp * t2 - 1.0 / 3628800.0;
    cp = cp * t2 + 1.0 / 40320.0;
    cp = cp * t2 - 1.0 / 720.0;
    cp = cp * t2 + 1.0 / 24.0;
    cp = cp * t2 - 1.0 / 2.0;
    cp = cp * t2 + 1.0;
    const double c_full = swap ? sn : cp; // cos(pi/2 * r)
    const double s_full = swap ? cp : sn; // sin(pi/2 * r)
    switch (quadrant & 3)
    {
    case 0: return c_full;
    case 1: return -s_full;
    case 2: return -c_full;
    default: return s_full;
    }
}

// standard normal from the four words of one Philox block
inline double det_normal(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3)
{
    const uint64_t a = (static_cast<uint64_t>(w0 >> 6) << 27) | (w1 >> 5);
    const uint64_t b = (static_cast<uint64_t>(w2 >> 6) << 27) | (w3 >> 5);
    const double u1 = static_cast<double>(a + 1) * 1.1102230246251565e-16; // (0, 1], 2^-53
    const double u2 = static_cast<double>(b) * 1.1102230246251565e-16;     // [0, 1)
    const double r = std::sqrt(-2.0 * det_log(u1));
    return r * det_cos2pi(u2);
}

} // namespace oracle
