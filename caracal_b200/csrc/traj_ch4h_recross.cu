// traj_ch4h_recross.cu -- instantiates the recross trajectory kernels for the "ch4h" surface.
// Four 16-bead child trajectories (256 threads) to a CTA instead of two: measured 11.60 against 11.74 ms per 1000 steps of
// the 1024-child batch (profiles/r2n_bench_ctpb256.json) -- eight warps that re-align once per step share their fetched
// instruction lines.  Set for this unit only: CRCL_CTPB also re-packs every other multi-warp kernel of a unit.
#ifndef CRCL_CTPB
#define CRCL_CTPB 256
#endif
#include "pes_ch4h.cuh"
#include "traj_inst.cuh"
namespace crcl {
CRCL_DECLARE_TRAJ(launch_ch4h_recross) { return launch_traj_pes<PesCH4H4, K_RECROSS>(nbeads, A, bias_mode, nose_q, s, nosup); }
}  // namespace crcl
