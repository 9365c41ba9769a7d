// water.cuh -- device tables and launch interface of the flexible SPC water box (water_kernels.cu).
// Replaces egrad_water.f90:36-333 (pes WATER_SPC, dispatched at gradient.f90:212-213) with the parameters
// of water_init.f90:75-107 and the periodic settings of set_periodic.f90:66-104.
#pragma once
#include <cuda_runtime.h>
#include "../../include/caracal_gpu.h"

namespace crcl {

struct WaterDev {
    int n, periodic, zahn;
    double box[3], coul_cut, zahn_a, zahn_par, pars[11];
    double* q;   // [n]
    int* is_O;   // [n]
};

int water_upload(const crcl_water_params* P, WaterDev** out, const char** err);
void water_free(WaterDev* D);
// nimg structures, AoS [img][atom][xyz]; d_V[nimg], d_g like d_xyz (both overwritten)
cudaError_t water_egrad(const WaterDev* D, const double* d_xyz, int nimg, double* d_V, double* d_g, cudaStream_t s,
                        long long* launches);

}  // namespace crcl
