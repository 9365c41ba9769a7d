// crcl_common.cuh -- shared definitions for the sm_100a kernels of caracal_b200.
#pragma once
#include <cstdint>
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

// 1 - tanh(x) and -sech^2(x) from a single exp.  The reference evaluates tanh() and
// cosh()**2 separately; these forms agree with them to a few ulp and stay accurate for
// large |x| where 1-tanh^2 would cancel.
CRCL_HD __forceinline__ void one_minus_tanh(double x, double& omt, double& msech2)
{
    const double e = exp(-2.0 * fabs(x));       // in (0,1]
    const double inv = 1.0 / (1.0 + e);
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
    for (int j = 0; j < 3; j++) a[ICOL][j] /= pivot;
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
