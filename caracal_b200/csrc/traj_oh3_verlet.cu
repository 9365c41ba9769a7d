// traj_oh3_verlet.cu -- instantiates the verlet trajectory kernels for the "oh3" surface.
#include "pes_oh3.cuh"
#include "traj_inst.cuh"
namespace crcl {
CRCL_DECLARE_TRAJ(launch_oh3_verlet) { return launch_traj_pes<PesOH3, K_VERLET>(nbeads, A, bias_mode, nose_q, s, nosup); }
}  // namespace crcl
