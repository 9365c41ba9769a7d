// traj_h3_recross.cu -- instantiates the recross trajectory kernels for the "h3" surface.
#include "pes_h3.cuh"
#include "traj_inst.cuh"
namespace crcl {
CRCL_DECLARE_TRAJ(launch_h3_recross) { return launch_traj_pes<PesH3, K_RECROSS>(nbeads, A, bias_mode, nose_q, s, nosup); }
}  // namespace crcl
