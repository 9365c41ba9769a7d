// traj_clnh3_recross.cu -- instantiates the recross trajectory kernels for the "clnh3" surface.
#include "pes_nh3x.cuh"
#include "traj_inst.cuh"
namespace crcl {
CRCL_DECLARE_TRAJ(launch_clnh3_recross) { return launch_traj_pes<PesClNH3, K_RECROSS>(nbeads, A, bias_mode, nose_q, s, nosup); }
}  // namespace crcl
