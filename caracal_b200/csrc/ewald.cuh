// ewald.cuh -- device state and launch interface of the SPME reciprocal-space kernels
// (ewald_kernels.cu).  Replaces ewald_recip.f90 (+ bsplgen.f90, ewald_adjust.f90, setchunk.f90)
// with the set-up quantities of set_periodic.f90:114-231 received as tables.
#pragma once
#include <cuda_runtime.h>
#include "../../include/caracal_gpu.h"

namespace crcl {

struct FftPlan {
    int nfac;
    int r[16];
};

struct EwaldDev {
    int nfft, bsorder;
    double box[3], a_ewald;
    double* bsmod;            // [3][nfft]
    // grow-only work space
    FftPlan plan;             // factors of nfft, in the order the Stockham stages take them
    int fft_L;                // lines per CTA of ew_fft_lines_kernel
    double2* tw;              // [2][nfft]: e^{-2 pi i t/nfft} and conjugates
    double2* grid;            // [nimg][nfft][nfft][nfft], x fastest
    size_t grid_cap;
    double* theta;            // [nimg*n][3][5][2]: B-spline values and first derivatives
    int* igrid;               // [nimg*n][3]
    size_t atom_cap;
    double* dq;               // [n] charges
    size_t q_cap;
};

int ewald_upload(const crcl_ewald_params* P, EwaldDev** out, const char** err);
void ewald_free(EwaldDev* E);
// xyz [nimg][n][3], q [n] (device); energy [nimg], grad [nimg][n][3] (device, overwritten)
int ewald_recip(EwaldDev* E, int n, int nimg, const double* d_xyz, const double* d_q, double* d_energy, double* d_grad,
                cudaStream_t s, long long* launches, const char** err);

}  // namespace crcl
