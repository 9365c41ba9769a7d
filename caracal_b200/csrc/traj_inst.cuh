// traj_inst.cuh -- launch dispatch (kernel kind x bead count) for one PES; each
// traj_<pes>_<kind>.cu instantiates exactly one (PES, kind) so the heavy kernels compile in
// parallel translation units.
#pragma once
#include "traj_kernel.cuh"

namespace crcl {

enum { K_VERLET = 0, K_MDINIT = 1, K_RECROSS = 2 };

// returns cudaSuccess / a CUDA error; *nosup = 1 if the bead count has no instantiation
typedef cudaError_t (*traj_launch_fn)(int nbeads, const TrajArgs& A, int bias_mode, double nose_q,
                                      cudaStream_t s, int* nosup);

template <class PES, int KIND, int NB>
static cudaError_t launch_one(const TrajArgs& A, int bias_mode, double nose_q, cudaStream_t s)
{
    using L = SmemLayout<PES::NATOMS, NB, PES::LANES, coop_scratch<PES>::value>;
    constexpr int gpb = Group<NB, PES::LANES>::GPB, tpb = Group<NB, PES::LANES>::TPB;
    const int grid = (A.ntraj + gpb - 1) / gpb;
    const size_t smem = L::bytes();
    if (grid <= 0) return cudaSuccess;
    cudaError_t e = cudaSuccess;
    if constexpr (KIND == K_VERLET) {
        if (smem > 48 * 1024)
            e = cudaFuncSetAttribute(verlet_kernel<PES, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        verlet_kernel<PES, NB><<<grid, tpb, smem, s>>>(A);
    } else if constexpr (KIND == K_MDINIT) {
        if (smem > 48 * 1024)
            e = cudaFuncSetAttribute(mdinit_kernel<PES, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        mdinit_kernel<PES, NB><<<grid, tpb, smem, s>>>(A, bias_mode, nose_q);
    } else {
        if (smem > 48 * 1024)
            e = cudaFuncSetAttribute(recross_kernel<PES, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        recross_kernel<PES, NB><<<grid, tpb, smem, s>>>(A);
    }
    return cudaGetLastError();
}

template <class PES, int KIND>
static cudaError_t launch_traj_pes(int nbeads, const TrajArgs& A, int bias_mode, double nose_q,
                                   cudaStream_t s, int* nosup)
{
    *nosup = 0;
    switch (nbeads) {
    case 1: return launch_one<PES, KIND, 1>(A, bias_mode, nose_q, s);
    case 2: return launch_one<PES, KIND, 2>(A, bias_mode, nose_q, s);
    case 4: return launch_one<PES, KIND, 4>(A, bias_mode, nose_q, s);
    case 8: return launch_one<PES, KIND, 8>(A, bias_mode, nose_q, s);
    case 16: return launch_one<PES, KIND, 16>(A, bias_mode, nose_q, s);
    case 32: return launch_one<PES, KIND, 32>(A, bias_mode, nose_q, s);
    case 64: return launch_one<PES, KIND, 64>(A, bias_mode, nose_q, s);
    case 128: return launch_one<PES, KIND, 128>(A, bias_mode, nose_q, s);
    }
    *nosup = 1;
    return cudaSuccess;
}

#define CRCL_DECLARE_TRAJ(name)                                                                   \
    cudaError_t name(int nbeads, const TrajArgs& A, int bias_mode, double nose_q, cudaStream_t s, \
                     int* nosup)
CRCL_DECLARE_TRAJ(launch_h3_verlet);
CRCL_DECLARE_TRAJ(launch_h3_mdinit);
CRCL_DECLARE_TRAJ(launch_h3_recross);
CRCL_DECLARE_TRAJ(launch_oh3_verlet);
CRCL_DECLARE_TRAJ(launch_oh3_mdinit);
CRCL_DECLARE_TRAJ(launch_oh3_recross);
CRCL_DECLARE_TRAJ(launch_ch4h_verlet);
CRCL_DECLARE_TRAJ(launch_ch4h_mdinit);
CRCL_DECLARE_TRAJ(launch_ch4h_recross);
CRCL_DECLARE_TRAJ(launch_brh2_verlet);
CRCL_DECLARE_TRAJ(launch_brh2_mdinit);
CRCL_DECLARE_TRAJ(launch_brh2_recross);
CRCL_DECLARE_TRAJ(launch_o3_verlet);
CRCL_DECLARE_TRAJ(launch_o3_mdinit);
CRCL_DECLARE_TRAJ(launch_o3_recross);
CRCL_DECLARE_TRAJ(launch_ch4oh_verlet);
CRCL_DECLARE_TRAJ(launch_ch4oh_mdinit);
CRCL_DECLARE_TRAJ(launch_ch4oh_recross);
CRCL_DECLARE_TRAJ(launch_geh4oh_verlet);
CRCL_DECLARE_TRAJ(launch_geh4oh_mdinit);
CRCL_DECLARE_TRAJ(launch_geh4oh_recross);
CRCL_DECLARE_TRAJ(launch_ch4cn_verlet);
CRCL_DECLARE_TRAJ(launch_ch4cn_mdinit);
CRCL_DECLARE_TRAJ(launch_ch4cn_recross);
CRCL_DECLARE_TRAJ(launch_clnh3_verlet);
CRCL_DECLARE_TRAJ(launch_clnh3_mdinit);
CRCL_DECLARE_TRAJ(launch_clnh3_recross);
CRCL_DECLARE_TRAJ(launch_nh3oh_verlet);
CRCL_DECLARE_TRAJ(launch_nh3oh_mdinit);
CRCL_DECLARE_TRAJ(launch_nh3oh_recross);
CRCL_DECLARE_TRAJ(launch_h2co_verlet);
CRCL_DECLARE_TRAJ(launch_h2co_mdinit);
CRCL_DECLARE_TRAJ(launch_h2co_recross);

}  // namespace crcl
