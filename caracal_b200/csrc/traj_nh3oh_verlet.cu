// traj_nh3oh_verlet.cu -- instantiates the verlet trajectory kernels for the "nh3oh" surface.
// FMA form of the free ring-polymer step: the step of this surface is 19 energy evaluations, the transform does not show,
// and nvcc 12.9 crashes on the tensor-core form in this unit
#define CRCL_DMMA_TRANSFORM 0
#include "pes_nh3x.cuh"
#include "traj_inst.cuh"
namespace crcl {
CRCL_DECLARE_TRAJ(launch_nh3oh_verlet) { return launch_traj_pes<PesNH3OH4, K_VERLET>(nbeads, A, bias_mode, nose_q, s, nosup); }
}  // namespace crcl
