// traj_ch4h_mdinit.cu -- instantiates the mdinit trajectory kernels for the "ch4h" surface.
#include "pes_ch4h.cuh"
#include "traj_inst.cuh"
namespace crcl {
CRCL_DECLARE_TRAJ(launch_ch4h_mdinit) { return launch_traj_pes<PesCH4H4, K_MDINIT>(nbeads, A, bias_mode, nose_q, s, nosup); }
}  // namespace crcl
