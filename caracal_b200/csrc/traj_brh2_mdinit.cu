// traj_brh2_mdinit.cu -- instantiates the mdinit trajectory kernels for the "brh2" surface.
#include "pes_brh2.cuh"
#include "traj_inst.cuh"
namespace crcl {
CRCL_DECLARE_TRAJ(launch_brh2_mdinit) { return launch_traj_pes<PesBrH2, K_MDINIT>(nbeads, A, bias_mode, nose_q, s, nosup); }
}  // namespace crcl
