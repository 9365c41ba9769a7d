// traj_clnh3_verlet.cu -- instantiates the verlet trajectory kernels for the "clnh3" surface.
#include "pes_nh3x.cuh"
#include "traj_inst.cuh"
namespace crcl {
CRCL_DECLARE_TRAJ(launch_clnh3_verlet) { return launch_traj_pes<PesClNH3, K_VERLET>(nbeads, A, bias_mode, nose_q, s, nosup); }
}  // namespace crcl
