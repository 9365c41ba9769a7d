// traj_oh3_recross.cu -- instantiates the recross trajectory kernels for the "oh3" surface.
#include "pes_oh3.cuh"
#include "traj_inst.cuh"
namespace crcl {
CRCL_DECLARE_TRAJ(launch_oh3_recross) { return launch_traj_pes<PesOH3, K_RECROSS>(nbeads, A, bias_mode, nose_q, s, nosup); }
}  // namespace crcl
