// traj_nh3oh_recross.cu -- instantiates the recross trajectory kernels for the "nh3oh" surface.
#include "pes_nh3x.cuh"
#include "traj_inst.cuh"
namespace crcl {
CRCL_DECLARE_TRAJ(launch_nh3oh_recross) { return launch_traj_pes<PesNH3OH4, K_RECROSS>(nbeads, A, bias_mode, nose_q, s, nosup); }
}  // namespace crcl
