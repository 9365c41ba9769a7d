// crcl_common.cuh -- shared definitions for the sm_100a kernels of caracal_b200.
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>
#include "../../include/caracal_gpu.h"

// gfortran REAL*4 literal rounding: FL(1.5e-3) == (double)1.5e-3f
#ifdef CRCL_LITERALS_EXACT
#define FL(x) ((double)(x))
#else
#define FL(x) ((double)(x##f))
#endif

#ifdef __CUDACC__
#define CRCL_HD __host__ __device__
#else
#define CRCL_HD
#define __forceinline__ inline
#define __noinline__
#endif

namespace crcl {

constexpr double PI_QMDFF = 3.1415926535897932384626433832795029;  // qmdff.f90:44
constexpr double PI_UMBR = 3.1415926535897932384;                  // umbrella.f90:98

CRCL_HD __forceinline__ double sqr(double x) { return x * x; }

// ---- branch-free FP64 elementary functions for the fused trajectory kernels (device only) ----------------------------------
// The toolkit's exp / acos / sqrt / division each carry a range check, a convergence barrier (BSSY/BSYNC) and a call to an
// out-of-line slow path (denormals, huge arguments).  In recross_kernel<PesCBE4<K6>,16> that was 724 barrier pairs and 477
// CALL sites: every one ends a scheduling region, so with 3.5 warps per scheduler the ~41 function evaluations of a CBE
// image ran as ~41 short dependent chains instead of being interleaved.  These versions are straight-line code (one MUFU
// seed + Newton steps, or a polynomial), accurate to <= 2 ulp on the arguments the surfaces produce (normal, finite
// numbers; positive for sqrt / rsqrt), and propagate NaN.  Coefficients: Chebyshev-node fits in 60-digit arithmetic
// (csrc/gen_math_coeffs.py), max relative error of the polynomials 1.6e-17 (exp) and 4.8e-18 (asin kernel).
// -DCRCL_LIBM_MATH restores the toolkit functions (A/B builds, profiles/build_variant.sh).
// -DCRCL_FM_ON_HOST compiles the same code for the CPU with a 20-bit model of the two MUFU seeds (tests/host_harness).
#if (defined(__CUDA_ARCH__) || defined(CRCL_FM_ON_HOST)) && !defined(CRCL_LIBM_MATH)
#define CRCL_FM_ACTIVE 1
namespace fm {
#ifdef __CUDA_ARCH__
__device__ __forceinline__ double seed_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
__device__ __forceinline__ double seed_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
__device__ __forceinline__ int lo_int(double t) { return __double2loint(t); }
__device__ __forceinline__ int hi_int(double t) { return __double2hiint(t); }
__device__ __forceinline__ double hilo(int hi, int lo) { return __hiloint2double(hi, lo); }
__device__ __forceinline__ double pow2(int k) { return __hiloint2double((k + 1023) << 20, 0); }
#else
// host model of MUFU.RCP64H / MUFU.RSQ64H: the upper word only (20 mantissa bits), the lower word zero
inline double hi_word_only(double v)
{
    unsigned long long u;
    memcpy(&u, &v, 8);
    u &= 0xFFFFFFFF00000000ull;
    memcpy(&v, &u, 8);
    return v;
}
inline double seed_rcp(double x) { return hi_word_only(1.0 / hi_word_only(x)); }
inline double seed_rsqrt(double x) { return hi_word_only(1.0 / ::sqrt(hi_word_only(x))); }
inline int lo_int(double t)
{
    unsigned long long u;
    memcpy(&u, &t, 8);
    return (int)(unsigned)(u & 0xFFFFFFFFull);
}
inline int hi_int(double t)
{
    unsigned long long u;
    memcpy(&u, &t, 8);
    return (int)(unsigned)(u >> 32);
}
inline double hilo(int hi, int lo)
{
    const unsigned long long u = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
    double v;
    memcpy(&v, &u, 8);
    return v;
}
inline double pow2(int k)
{
    const unsigned long long u = (unsigned long long)(unsigned)((k + 1023) << 20) << 32;
    double v;
    memcpy(&v, &u, 8);
    return v;
}
#endif
// 1/x: MUFU.RCP64H (>= 20 bits) + one cubic and one quadratic Newton step
CRCL_HD __forceinline__ double rcp(double x)
{
    double y = seed_rcp(x);
    double e = fma(-x, y, 1.0);
    e = fma(e, e, e);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}
CRCL_HD __forceinline__ double div(double a, double b)
{
    const double y = rcp(b), q = a * y;
    return fma(fma(-b, q, a), y, q);
}
// 1/sqrt(x), x > 0 normal: MUFU.RSQ64H + one cubic step y (1 + e/2 + 3 e^2/8), e = 1 - x y^2, + one quadratic step
CRCL_HD __forceinline__ double rsqrt(double x)
{
    double y = seed_rsqrt(x);
    double h = 0.5 * y, t = x * y;
    double e = fma(-t, h, 0.5);                  // e/2
    y = fma(y, fma(1.5 * e, e, e), y);
    h = 0.5 * y, t = x * y;
    e = fma(-t, h, 0.5);
    return fma(y, e, y);
}
// sqrt(x), x > 0 normal: x * rsqrt(x) with one residual correction
template <bool ZERO_OK = false>
CRCL_HD __forceinline__ double sqrt(double x)
{
    double y = seed_rsqrt(ZERO_OK ? fmax(x, 1.0e-300) : x);   // ZERO_OK: sqrt(0) = 0 instead of 0 * inf
    double h = 0.5 * y, t = x * y;
    double e = fma(-t, h, 0.5);
    y = fma(y, fma(1.5 * e, e, e), y);
    t = x * y;
    return fma(fma(-t, t, x), 0.5 * y, t);
}
// e^x: x = k ln2 + r, |r| <= ln2/2, degree-11 polynomial, scaled by 2^k; 0 below 2^-1022 (no denormals), inf above 2^1023
CRCL_HD __forceinline__ double exp(double xin)
{
    // 1024 <= |x| <= inf: beyond 2^31 ln2 the integer part no longer fits the low word of the magic-number sum, beyond
    // 2^52 it is not even an integer -- evaluate at +-1024 instead, which lands on the same end of the range (0 or inf)
    // (BKMP2's H2 singlet curve calls exp(-2e12) at R = 30 a0).  NaN is outside the interval and takes the normal path.
    const int hx = hi_int(xin);
    const bool far = (unsigned)((hx & 0x7fffffff) - 0x40900000) <= (unsigned)(0x7ff00000 - 0x40900000);
    const double x = far ? (hx < 0 ? -1024.0 : 1024.0) : xin;
    const double t = fma(x, 1.4426950408889634, 6755399441055744.0);
    int k = lo_int(t);
    const double kf = t - 6755399441055744.0;
    double r = fma(kf, -6.93147180369123816490e-01, x);
    r = fma(kf, -1.90821492927058770002e-10, r);
    const double r2 = r * r;
    // e^r = 1 + (r + r^2 Q(r)), Q of degree 9 in two interleaved chains; the last two operations keep the error below 1 ulp
    double qe = 2.7620088445409746e-07, qo = 2.510038549551032e-08;                       // r^8, r^9 of Q
    qe = fma(qe, r2, 2.4801521295954376e-05), qo = fma(qo, r2, 2.7557268459997064e-06);   // r^6, r^7
    qe = fma(qe, r2, 0.0013888888917213717), qo = fma(qo, r2, 0.00019841269863053618);    // r^4, r^5
    qe = fma(qe, r2, 0.04166666666662413), qo = fma(qo, r2, 0.008333333333330062);        // r^2, r^3
    qe = fma(qe, r2, 0.5000000000000001), qo = fma(qo, r2, 0.16666666666666669);          // r^0, r^1
    const double p = 1.0 + fma(fma(qo, r, qe), r2, r);
    // 2^k from the biased exponent clamped to [0, 2047]: 0 gives +0.0 (results below 2^-1021 are flushed to zero, as the
    // library's exp(-746) is 0 -- code that tests a term against exactly zero takes the same branch), 2047 gives +inf
    int e = k + 1023;
    e = e < 2 ? 0 : (e > 2047 ? 2047 : e);   // e = 1 with p < 1 would still be a denormal
    return p * pow2(e - 1023);
}
// log(x), x > 0 normal: x = m 2^e with m in [sqrt(1/2), sqrt(2)), log m = 2 atanh(s), s = (m - 1)/(m + 1), |s| <= 0.1716:
// atanh(s) = s + s z P(z), z = s^2, P of degree 7 (relative error of the series 5e-19); e ln2 added in two parts
CRCL_HD __forceinline__ double log(double x)
{
    int hi = hi_int(x) + (0x3ff00000 - 0x3fe6a09e);
    const int e = (hi >> 20) - 0x3ff;
    hi = (hi & 0x000fffff) + 0x3fe6a09e;
    const double m = hilo(hi, lo_int(x));
    const double f = m - 1.0, d = 2.0 + f;
    const double y = rcp(d);
    double sv = f * y;
    sv = fma(fma(-d, sv, f), y, sv);
    const double z = sv * sv;
    double P = 0.06544017425463777;
    P = fma(P, z, 0.06634319184633412);
    P = fma(P, z, 0.07693122368810833);
    P = fma(P, z, 0.09090897773760569);
    P = fma(P, z, 0.11111111196290523);
    P = fma(P, z, 0.14285714285399892);
    P = fma(P, z, 0.20000000000000442);
    P = fma(P, z, 0.3333333333333333);
    const double t = sv * z * P;
    const double ef = (double)e;
    // 2 sv + (2 t + e ln2_lo) + e ln2_hi
    const double r = fma(ef, 1.90821492927058770002e-10, t + t);
    return fma(ef, 6.93147180369123816490e-01, (sv + sv) + r);
}
// x^y, x > 0: exp(y log x); relative error ~ |y log x| ulp (the surfaces use it with |y log x| < 10)
CRCL_HD __forceinline__ double pow(double x, double y) { return fm::exp(y * fm::log(x)); }
// acos(x), |x| <= 1: asin(s) = s + s z P(z), z = s^2 <= 1/4; |x| > 1/2 through s = sqrt((1 - |x|)/2)
CRCL_HD __forceinline__ double acos(double x)
{
    const double a = fabs(x);
    const bool big = a > 0.5;
    const double zb = fma(-0.5, a, 0.5);
    const double z = big ? zb : a * a;
    const double s = big ? fm::sqrt<true>(zb) : a;   // zb = 0 at x = +-1
    const double z2 = z * z;
    double pe = 0.02886097438799853, po = -0.014999855783811349;       // z^12, z^11
    pe = fma(pe, z2, 0.017493569476248205), po = fma(po, z2, 0.005424230839804001);   // z^10, z^9
    pe = fma(pe, z2, 0.010330370295404406), po = fma(po, z2, 0.011478047472510145);   // z^8, z^7
    pe = fma(pe, z2, 0.013971325328728771), po = fma(po, z2, 0.017352385393480947);   // z^6, z^5
    pe = fma(pe, z2, 0.022372173243832207), po = fma(po, z2, 0.0303819441312372);     // z^4, z^3
    pe = fma(pe, z2, 0.04464285714644605), po = fma(po, z2, 0.07499999999998389);     // z^2, z^1
    pe = fma(pe, z2, 0.16666666666666669);                                             // z^0
    const double P = fma(po, z, pe);
    const double as = fma(s * z, P, s);            // asin(s)
    // |x| <= 1/2: pi/2 -+ asin(|x|);  |x| > 1/2: 2 asin(s) or pi - 2 asin(s)
    const double hi = 1.5707963267948966, lo = 6.123233995736766e-17;   // pi/2 = hi + lo
    const double small = (x < 0.0) ? (hi + (as + lo)) : (hi - (as - lo));
    const double large = (x < 0.0) ? fma(-2.0, as, 2.0 * hi) + 2.0 * lo : 2.0 * as;
    return big ? large : small;
}
}  // namespace fm
#define CRCL_EXP(x) ::crcl::fm::exp(x)
#define CRCL_ACOS(x) ::crcl::fm::acos(x)
#define CRCL_LOG(x) ::crcl::fm::log(x)
#define CRCL_POW(x, y) ::crcl::fm::pow(x, y)
#define CRCL_SQRT0(x) ::crcl::fm::sqrt<true>(x)
#define CRCL_SQRT(x) ::crcl::fm::sqrt<>(x)
#define CRCL_RSQRT(x) ::crcl::fm::rsqrt(x)
#define CRCL_RCP(x) ::crcl::fm::rcp(x)
#define CRCL_DIV(a, b) ::crcl::fm::div(a, b)
#else
#define CRCL_EXP(x) exp(x)
#define CRCL_ACOS(x) acos(x)
#define CRCL_LOG(x) log(x)
#define CRCL_POW(x, y) pow(x, y)
#define CRCL_SQRT0(x) sqrt(x)
#define CRCL_SQRT(x) sqrt(x)
#ifdef __CUDA_ARCH__
#define CRCL_RSQRT(x) rsqrt(x)
#else
#define CRCL_RSQRT(x) (1.0 / sqrt(x))
#endif
#define CRCL_RCP(x) (1.0 / (x))
#define CRCL_DIV(a, b) ((a) / (b))
#endif
#define CRCL_OMT_INLINE __forceinline__

// s = sqrt(x) and is = 1/sqrt(x) together (a distance and its inverse): one MUFU seed serves both
CRCL_HD __forceinline__ void sqrt_rsqrt(double x, double& s, double& is)
{
#ifdef CRCL_FM_ACTIVE
    const double y = fm::rsqrt(x), t = x * y;
    s = fma(fma(-t, t, x), 0.5 * y, t);
    is = y;
#else
    s = sqrt(x);
    is = 1.0 / s;
#endif
}

// 1 - tanh(x) and -sech^2(x) from a single exp.  The reference evaluates tanh() and
// cosh()**2 separately; these forms agree with them to a few ulp and stay accurate for
// large |x| where 1-tanh^2 would cancel.
CRCL_HD CRCL_OMT_INLINE void one_minus_tanh(double x, double& omt, double& msech2)
{
    const double e = CRCL_EXP(-2.0 * fabs(x));       // in (0,1]
    const double inv = CRCL_RCP(1.0 + e);
    const double small = 2.0 * e * inv;         // 1 - tanh(|x|)
    omt = (x >= 0.0) ? small : 2.0 - small;     // 1 - tanh(x)
    msech2 = -4.0 * e * inv * inv;              // -1/cosh(x)^2
}

#ifdef __CUDACC__
// 3x3 inverse by Gauss-Jordan with full pivoting, as invert.f90:38-123; returns 1 if singular.
// Same operations in the same order as the reference's loops, but every array index is a compile-time constant (the
// data-dependent pivot row / column select one of 3 x 3 fully unrolled bodies), so the matrix and the pivot bookkeeping
// stay in registers.  The straightforward transcription kept ipivot / indxr / indxc and the matrix in local memory and
// was a visible part of every step that removes the net rotation (verlet.f90:1300-1306: dynamic.x, umbrella phases).
namespace inv3 {
template <int R1, int R2>
__device__ __forceinline__ void swap_rows(double (&a)[3][3])
{
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const double t = a[R1][j];
        a[R1][j] = a[R2][j];
        a[R2][j] = t;
    }
}
template <int C1, int C2>
__device__ __forceinline__ void swap_cols(double (&a)[3][3])
{
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double t = a[k][C1];
        a[k][C1] = a[k][C2];
        a[k][C2] = t;
    }
}
// invert.f90:84-103 for a compile-time pivot column: row swap (irow is data), normalise the pivot row, eliminate
template <int ICOL>
__device__ __forceinline__ int eliminate(double (&a)[3][3], int irow)
{
    constexpr int O1 = (ICOL + 1) % 3, O2 = (ICOL + 2) % 3;
    if (irow == O1)
        swap_rows<O1, ICOL>(a);
    else if (irow == O2)
        swap_rows<O2, ICOL>(a);
    if (a[ICOL][ICOL] == 0.0) return 1;
    const double pivot = a[ICOL][ICOL];
    a[ICOL][ICOL] = 1.0;
#pragma unroll
    for (int j = 0; j < 3; j++) a[ICOL][j] = CRCL_DIV(a[ICOL][j], pivot);
#pragma unroll
    for (int j = 0; j < 3; j++)
        if (j != ICOL) {
            const double t = a[j][ICOL];
            a[j][ICOL] = 0.0;
#pragma unroll
            for (int k = 0; k < 3; k++) a[j][k] -= a[ICOL][k] * t;
        }
    return 0;
}
__device__ __forceinline__ void unswap(double (&a)[3][3], int r, int c)
{
    if (r == c) return;
    const int lo = r < c ? r : c, hi = r < c ? c : r;
    if (lo == 0 && hi == 1)
        swap_cols<0, 1>(a);
    else if (lo == 0 && hi == 2)
        swap_cols<0, 2>(a);
    else
        swap_cols<1, 2>(a);
}
}  // namespace inv3
// Inverse of a SYMMETRIC 3x3 matrix by cofactors: straight-line code (21 multiply-adds and one reciprocal) where the
// Gauss-Jordan form above is a tree of data-dependent pivot branches -- ~10 % of the stall samples and a good part of the
// instruction footprint of every step that removes the net rotation (ncu, verlet_kernel<PesH3,16>,
// profiles/r2m_verlet_h3_nb16.txt).  Differs from invert.f90's pivoted elimination by rounding only (relative error of
// both ~ cond(A) eps; a ring polymer's inertia tensor has cond <= 1e4); singular (returns 1) iff the determinant is exactly
// zero, which is what the elimination reports for the cases that occur -- a collinear arrangement on an axis.
__device__ __forceinline__ int invert3_sym(double (&a)[3][3])
{
    const double p = a[0][0], q = a[0][1], r = a[0][2], u = a[1][1], v = a[1][2], w = a[2][2];
    const double A = fma(u, w, -v * v), B = fma(r, v, -q * w), C = fma(q, v, -r * u);
    const double D = fma(p, w, -r * r), E = fma(q, r, -p * v), F = fma(p, u, -q * q);
    const double det = fma(p, A, fma(q, B, r * C));
    if (det == 0.0) return 1;
    const double id = CRCL_RCP(det);
    a[0][0] = A * id, a[0][1] = B * id, a[0][2] = C * id;
    a[1][0] = B * id, a[1][1] = D * id, a[1][2] = E * id;
    a[2][0] = C * id, a[2][1] = E * id, a[2][2] = F * id;
    return 0;
}
__device__ __forceinline__ int invert3(double (&a)[3][3])
{
    int ip0 = 0, ip1 = 0, ip2 = 0;                 // ipivot(1:3)
    int xr[3], xc[3];                              // indxr, indxc: only ever indexed by the unrolled i
    int irow = 0, icol = 0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double big = 0.0;
        const int ip[3] = {ip0, ip1, ip2};
#pragma unroll
        for (int j = 0; j < 3; j++)
            if (ip[j] != 1) {
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    if (ip[k] == 0) {
                        if (fabs(a[j][k]) >= big) {
                            big = fabs(a[j][k]);
                            irow = j;
                            icol = k;
                        }
                    } else if (ip[k] > 1)
                        return 1;
                }
            }
        int sing;
        if (icol == 0) {
            ip0++;
            sing = inv3::eliminate<0>(a, irow);
        } else if (icol == 1) {
            ip1++;
            sing = inv3::eliminate<1>(a, irow);
        } else {
            ip2++;
            sing = inv3::eliminate<2>(a, irow);
        }
        if (sing) return 1;
        xr[i] = irow;
        xc[i] = icol;
    }
#pragma unroll
    for (int i = 2; i >= 0; i--) inv3::unswap(a, xr[i], xc[i]);
    return 0;
}

#endif

}  // namespace crcl
