// traj_brh2_recross.cu -- instantiates the recross trajectory kernels for the "brh2" surface.
#include "pes_brh2.cuh"
#include "traj_inst.cuh"
namespace crcl {
CRCL_DECLARE_TRAJ(launch_brh2_recross) { return launch_traj_pes<PesBrH2, K_RECROSS>(nbeads, A, bias_mode, nose_q, s, nosup); }
}  // namespace crcl
