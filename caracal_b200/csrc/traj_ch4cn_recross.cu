// traj_ch4cn_recross.cu -- instantiates the recross trajectory kernels for the "ch4cn" surface.
#include "pes_ch4oh.cuh"
#include "traj_inst.cuh"
namespace crcl {
CRCL_DECLARE_TRAJ(launch_ch4cn_recross) { return launch_traj_pes<PesCH4CN4, K_RECROSS>(nbeads, A, bias_mode, nose_q, s, nosup); }
}  // namespace crcl
