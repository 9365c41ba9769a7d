// traj_o3_verlet.cu -- instantiates the verlet trajectory kernels for the "o3" surface.
#include "pes_o3.cuh"
#include "traj_inst.cuh"
namespace crcl {
CRCL_DECLARE_TRAJ(launch_o3_verlet) { return launch_traj_pes<PesO3, K_VERLET>(nbeads, A, bias_mode, nose_q, s, nosup); }
}  // namespace crcl
