// dgevb.cuh -- device parameters and launch interface of the DG-EVB mixing kernel.
#pragma once
#include <cuda_runtime.h>
#include "../../include/caracal_gpu.h"

namespace crcl {

struct DgevbDev {
    int mode, npoints, nat6;
    int* coord_def;      // [nat6][5]: type, atoms 0-based
    double* point_int;   // [npoints][nat6]
    double* alph;        // [npoints]
    double* b_vec;       // [mat_size]
    double g_thres;
};

int dgevb_upload(const crcl_dgevb_params* E, int natoms, DgevbDev** out, const char** err);
void dgevb_free(DgevbDev* P);
cudaError_t dgevb_mix(const DgevbDev* P, int natoms, const double* xyz, int nimg, const double* V1, const double* G1,
                      const double* V2, const double* G2, double* V, double* G, cudaStream_t s);

}  // namespace crcl
