// traj_brh2_verlet.cu -- instantiates the verlet trajectory kernels for the "brh2" surface.
#include "pes_brh2.cuh"
#include "traj_inst.cuh"
namespace crcl {
CRCL_DECLARE_TRAJ(launch_brh2_verlet) { return launch_traj_pes<PesBrH2, K_VERLET>(nbeads, A, bias_mode, nose_q, s, nosup); }
}  // namespace crcl
