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
// 3x3 inverse by Gauss-Jordan with full pivoting, as invert.f90:38-123; returns 1 if singular
__device__ __forceinline__ int invert3(double a[3][3])
{
    int ipivot[3] = {0, 0, 0}, indxr[3], indxc[3];
    int irow = 0, icol = 0;
    for (int i = 0; i < 3; i++) {
        double big = 0.0;
        for (int j = 0; j < 3; j++)
            if (ipivot[j] != 1)
                for (int k = 0; k < 3; k++) {
                    if (ipivot[k] == 0) {
                        if (fabs(a[j][k]) >= big) {
                            big = fabs(a[j][k]);
                            irow = j;
                            icol = k;
                        }
                    } else if (ipivot[k] > 1)
                        return 1;
                }
        ipivot[icol]++;
        if (irow != icol)
            for (int j = 0; j < 3; j++) {
                const double t = a[irow][j];
                a[irow][j] = a[icol][j];
                a[icol][j] = t;
            }
        indxr[i] = irow;
        indxc[i] = icol;
        if (a[icol][icol] == 0.0) return 1;
        const double pivot = a[icol][icol];
        a[icol][icol] = 1.0;
        for (int j = 0; j < 3; j++) a[icol][j] /= pivot;
        for (int j = 0; j < 3; j++)
            if (j != icol) {
                const double t = a[j][icol];
                a[j][icol] = 0.0;
                for (int k = 0; k < 3; k++) a[j][k] -= a[icol][k] * t;
            }
    }
    for (int i = 2; i >= 0; i--)
        if (indxr[i] != indxc[i])
            for (int k = 0; k < 3; k++) {
                const double t = a[k][indxr[i]];
                a[k][indxr[i]] = a[k][indxc[i]];
                a[k][indxc[i]] = t;
            }
    return 0;
}

#endif

}  // namespace crcl
