// traj_h2co_mdinit.cu -- instantiates the mdinit trajectory kernels for the "h2co" surface.
// FMA form of the free ring-polymer step: a step of this surface is 25 evaluations of a 1561-term polynomial, the transform
// does not show (and the units of the other difference-quotient surface crash nvcc 12.9 with the tensor-core form)
#define CRCL_DMMA_TRANSFORM 0
#include "pes_h2co.cuh"
#include "traj_inst.cuh"
namespace crcl {
CRCL_DECLARE_TRAJ(launch_h2co_mdinit) { return launch_traj_pes<PesH2CO4, K_MDINIT>(nbeads, A, bias_mode, nose_q, s, nosup); }
}  // namespace crcl
