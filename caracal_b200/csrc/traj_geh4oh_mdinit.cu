// traj_geh4oh_mdinit.cu -- instantiates the mdinit trajectory kernels for the "geh4oh" surface.
#include "pes_ch4oh.cuh"
#include "traj_inst.cuh"
namespace crcl {
CRCL_DECLARE_TRAJ(launch_geh4oh_mdinit) { return launch_traj_pes<PesGeH4OH4, K_MDINIT>(nbeads, A, bias_mode, nose_q, s, nosup); }
}  // namespace crcl
