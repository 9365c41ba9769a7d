// spread_host.cu -- the ownership map of the spread trajectory forms (PesSpread, csrc/traj_inst.cuh) callable on the CPU:
// which component (atom * 3 + xyz) lane x of a bead owns in slot k, -1 for none.  Test infrastructure.
#include "../../include/caracal_gpu.h"
#include "../../caracal_b200/csrc/pes_h3.cuh"
#include "../../caracal_b200/csrc/pes_oh3.cuh"
#include "../../caracal_b200/csrc/traj_inst.cuh"

template <class P>
static int owned_of(int L, int lane, int k, int* nown)
{
    using namespace crcl;
    switch (L) {
    case 16: *nown = PesSpread<P, 16>::NOWN; return PesSpread<P, 16>::owned(lane, k);
    case 8: *nown = PesSpread<P, 8>::NOWN; return PesSpread<P, 8>::owned(lane, k);
    case 4: *nown = PesSpread<P, 4>::NOWN; return PesSpread<P, 4>::owned(lane, k);
    case 2: *nown = PesSpread<P, 2>::NOWN; return PesSpread<P, 2>::owned(lane, k);
    }
    return -2;
}

// the spread evaluation of one lane: energy share and the gradient of the lane's owned slots (on the CPU the lanes of a
// surface with a shared London term each evaluate all three pair curves: the shuffles exist on the device only)
template <class P, int L>
static int eval_one(const double* q, int lane, double* V, double* gown)
{
    return crcl::PesSpread<P, L>::eval_coop([&](int c) { return q[c]; }, lane, 0u, *V, gown);
}
template <class P>
static int eval_of(int L, const double* q, int lane, double* V, double* gown)
{
    switch (L) {
    case 16: return eval_one<P, 16>(q, lane, V, gown);
    case 8: return eval_one<P, 8>(q, lane, V, gown);
    case 4: return eval_one<P, 4>(q, lane, V, gown);
    case 2: return eval_one<P, 2>(q, lane, V, gown);
    }
    return -2;
}
extern "C" int hh_spread_eval(int pes, int L, const double* q, int lane, double* V, double* gown)
{
    if (pes == CRCL_PES_H3) return eval_of<crcl::PesH3>(L, q, lane, V, gown);
    if (pes == CRCL_PES_OH3) return eval_of<crcl::PesOH3>(L, q, lane, V, gown);
    return -2;
}
extern "C" int hh_plain_eval(int pes, const double* q, double* V, double* g)
{
    if (pes == CRCL_PES_H3) return crcl::PesH3::eval(q, *V, g);
    if (pes == CRCL_PES_OH3) return crcl::PesOH3::eval(q, *V, g);
    return -2;
}

extern "C" int hh_spread_owned(int pes, int L, int lane, int k, int* nown)
{
    if (pes == CRCL_PES_H3) return owned_of<crcl::PesH3>(L, lane, k, nown);
    if (pes == CRCL_PES_OH3) return owned_of<crcl::PesOH3>(L, lane, k, nown);
    return -2;
}
