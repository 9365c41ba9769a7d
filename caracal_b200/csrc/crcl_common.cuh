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

}  // namespace crcl
