// traj_geh4oh_recross.cu -- instantiates the recross trajectory kernels for the "geh4oh" surface.
#include "pes_ch4oh.cuh"
#include "traj_inst.cuh"
namespace crcl {
CRCL_DECLARE_TRAJ(launch_geh4oh_recross) { return launch_traj_pes<PesGeH4OH4, K_RECROSS>(nbeads, A, bias_mode, nose_q, s, nosup); }
}  // namespace crcl
