// rng.cuh -- counter-based normal deviates for the Andersen thermostat.
//
// Replaces andersen.f90:84-126 (Marsaglia polar on gfortran's random_number, xoshiro256**),
// whose stream exists only inside one gfortran build.  Philox4x32-10 (Salmon, Moraes, Dror,
// Shaw, SC'11) keyed by the run seed, counter = (pair, bead, event, trajectory):
//   trajectory : global index of the ring polymer (child pair index for recrossing)
//   event      : how many full momentum resamplings this trajectory has consumed
//   bead       : bead index
//   pair       : m/2 for momentum component m = atom*3 + xyz of that bead
// One Philox block gives two 53-bit uniforms in (0,1] -> one Box-Muller pair; component m
// takes element m&1.  Results depend only on (seed, trajectory, event, bead, m): independent
// of launch geometry and of the number of GPUs.
#pragma once
#include "crcl_common.cuh"

namespace crcl {

CRCL_HD __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

CRCL_HD __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                           uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t h0 = mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0;
        c1 = l1;
        c2 = n2;
        c3 = l0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

CRCL_HD __forceinline__ void normal_pair(uint64_t seed, uint32_t traj, uint32_t event, uint32_t bead,
                                         uint32_t pair, double& z0, double& z1)
{
    uint32_t o[4];
    philox4x32_10(pair, bead, event, traj, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    const double u1 =
        ((double)((((uint64_t)o[0]) << 21) | (o[1] >> 11)) + 1.0) * (1.0 / 9007199254740992.0);
    const double u2 =
        ((double)((((uint64_t)o[2]) << 21) | (o[3] >> 11)) + 1.0) * (1.0 / 9007199254740992.0);
    const double r = sqrt(-2.0 * log(u1));
    const double a = 6.283185307179586476925286766559 * u2;
    double s, c;
#ifdef __CUDA_ARCH__
    sincos(a, &s, &c);
#else
    s = sin(a);
    c = cos(a);
#endif
    z0 = r * c;
    z1 = r * s;
}

}  // namespace crcl
