// traj_h3_verlet.cu -- instantiates the verlet trajectory kernels for the "h3" surface.
#include "pes_h3.cuh"
#include "traj_inst.cuh"
namespace crcl {
CRCL_DECLARE_TRAJ(launch_h3_verlet) { return launch_traj_pes<PesH3, K_VERLET>(nbeads, A, bias_mode, nose_q, s, nosup); }
}  // namespace crcl
