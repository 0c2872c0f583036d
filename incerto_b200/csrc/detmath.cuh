// Deterministic f64 transforms of the draw contract ("normal draw").
//
// examples/traders.rs:148-153 samples rand_distr::Normal; the contract replaces it by Box-Muller on
// 53-bit uniforms whose ln / cos are fixed polynomials built only from IEEE-754 add, mul, div and sqrt
// with round-to-nearest (the _rn intrinsics keep nvcc from contracting a*b+c into an fma), so the device
// reproduces the host-side statement of the same formulas bit for bit.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace incerto
{

__device__ __forceinline__ double dm_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dm_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dm_sub(double a, double b) { return __dsub_rn(a, b); }
// q * x + c with two roundings
__device__ __forceinline__ double dm_horner(double q, double x, double c) { return __dadd_rn(__dmul_rn(q, x), c); }

__device__ __forceinline__ double det_log(double x)
{
    unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(x));
    int e = static_cast<int>((bits >> 52) & 0x7ffull) - 1023;
    bits = (bits & 0x000fffffffffffffull) | 0x3ff0000000000000ull;
    double m = __longlong_as_double(static_cast<long long>(bits));
    if (m > 1.4142135623730951)
    {
        m = dm_mul(m, 0.5);
        e += 1;
    }
    const double f = dm_sub(m, 1.0);
    const double s = __ddiv_rn(f, dm_add(2.0, f));
    const double s2 = dm_mul(s, s);
    double q = 1.0 / 21.0;
    q = dm_horner(q, s2, 1.0 / 19.0);
    q = dm_horner(q, s2, 1.0 / 17.0);
    q = dm_horner(q, s2, 1.0 / 15.0);
    q = dm_horner(q, s2, 1.0 / 13.0);
    q = dm_horner(q, s2, 1.0 / 11.0);
    q = dm_horner(q, s2, 1.0 / 9.0);
    q = dm_horner(q, s2, 1.0 / 7.0);
    q = dm_horner(q, s2, 1.0 / 5.0);
    q = dm_horner(q, s2, 1.0 / 3.0);
    q = dm_horner(q, s2, 1.0);
    return dm_add(dm_mul(static_cast<double>(e), 0.6931471805599453), dm_mul(dm_mul(2.0, s), q));
}

__device__ __forceinline__ double det_cos2pi(double u)
{
    const double t = dm_mul(u, 4.0);
    const int quadrant = static_cast<int>(t);
    const double r = dm_sub(t, static_cast<double>(quadrant));
    const bool swap = r > 0.5;
    const double rr = swap ? dm_sub(1.0, r) : r;
    const double th = dm_mul(rr, 1.5707963267948966);
    const double t2 = dm_mul(th, th);
    double sp = 1.0 / 355687428096000.0;
    sp = dm_horner(sp, t2, -1.0 / 1307674368000.0);
    sp = dm_horner(sp, t2, 1.0 / 6227020800.0);
    sp = dm_horner(sp, t2, -1.0 / 39916800.0);
    sp = dm_horner(sp, t2, 1.0 / 362880.0);
    sp = dm_horner(sp, t2, -1.0 / 5040.0);
    sp = dm_horner(sp, t2, 1.0 / 120.0);
    sp = dm_horner(sp, t2, -1.0 / 6.0);
    sp = dm_horner(sp, t2, 1.0);
    const double sn = dm_mul(th, sp);
    double cp = -1.0 / 6402373705728000.0;
    cp = dm_horner(cp, t2, 1.0 / 20922789888000.0);
    cp = dm_horner(cp, t2, -1.0 / 87178291200.0);
    cp = dm_horner(cp, t2, 1.0 / 479001600.0);
    cp = dm_horner(cp, t2, -1.0 / 3628800.0);
    cp = dm_horner(cp, t2, 1.0 / 40320.0);
    cp = dm_horner(cp, t2, -1.0 / 720.0);
    cp = dm_horner(cp, t2, 1.0 / 24.0);
    cp = dm_horner(cp, t2, -1.0 / 2.0);
    cp = dm_horner(cp, t2, 1.0);
    const double c_full = swap ? sn : cp;
    const double s_full = swap ? cp : sn;
    switch (quadrant & 3)
    {
    case 0: return c_full;
    case 1: return -s_full;
    case 2: return -c_full;
    default: return s_full;
    }
}

__device__ __forceinline__ double det_normal(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3)
{
    const unsigned long long a = (static_cast<unsigned long long>(w0 >> 6) << 27) | (w1 >> 5);
    const unsigned long long b = (static_cast<unsigned long long>(w2 >> 6) << 27) | (w3 >> 5);
    const double u1 = dm_mul(static_cast<double>(a + 1), 1.1102230246251565e-16);
    const double u2 = dm_mul(static_cast<double>(b), 1.1102230246251565e-16);
    const double r = __dsqrt_rn(dm_mul(-2.0, det_log(u1)));
    return dm_mul(r, det_cos2pi(u2));
}

} // namespace incerto
