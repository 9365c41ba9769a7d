// traj_inst.cuh -- launch dispatch (kernel kind x bead count) for one PES; each
// traj_<pes>_<kind>.cu instantiates exactly one (PES, kind) so the heavy kernels compile in
// parallel translation units.
#pragma once
#include "traj_kernel.cuh"

namespace crcl {

enum { K_VERLET = 0, K_MDINIT = 1, K_RECROSS = 2 };

// returns cudaSuccess / a CUDA error; *nosup = 1 if the bead count has no instantiation
typedef cudaError_t (*traj_launch_fn)(int nbeads, const TrajArgs& A, int bias_mode, double nose_q,
                                      cudaStream_t s, int* nosup);

template <class PES, int KIND, int NB>
static cudaError_t launch_one(const TrajArgs& A, int bias_mode, double nose_q, cudaStream_t s)
{
    using L = SmemLayout<PES::NATOMS, NB, PES::LANES, coop_scratch<PES>::value>;
    constexpr int gpb = Group<NB, PES::LANES>::GPB, tpb = Group<NB, PES::LANES>::TPB;
    const int grid = (A.ntraj + gpb - 1) / gpb;
    const size_t smem = L::bytes();
    if (grid <= 0) return cudaSuccess;
    cudaError_t e = cudaSuccess;
    if constexpr (KIND == K_VERLET) {
        if (smem > 48 * 1024)
            e = cudaFuncSetAttribute(verlet_kernel<PES, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        verlet_kernel<PES, NB><<<grid, tpb, smem, s>>>(A);
    } else if constexpr (KIND == K_MDINIT) {
        if (smem > 48 * 1024)
            e = cudaFuncSetAttribute(mdinit_kernel<PES, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        mdinit_kernel<PES, NB><<<grid, tpb, smem, s>>>(A, bias_mode, nose_q);
    } else {
        if (smem > 48 * 1024)
            e = cudaFuncSetAttribute(recross_kernel<PES, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        recross_kernel<PES, NB><<<grid, tpb, smem, s>>>(A);
    }
    return cudaGetLastError();
}

// ---- spread forms: SMALL batches of trajectories that have fewer threads than components in the packed form ------------
// Batches of more than spread_max trajectories x beads (default CRCL_SPREAD_MAX_BEADS, crcl_set_spread_max_beads) keep the
// packed form: the spread forms are for batches that leave most SMs with at most one warp, where nothing but the latency
// of a step counts -- measured (r2an): the 1100 8-bead trajectories of the H + H2 umbrella phase ran 0.45 -> 0.97 s when
// they were spread two lanes per bead.
#ifndef CRCL_SPREAD_MAX_BEADS
#define CRCL_SPREAD_MAX_BEADS 256
#endif
// surfaces with a lane-split evaluation for the spread form (P::eval_split<L>, P::SPLIT_OK)
template <class P, class = void>
struct has_split {
    static constexpr bool value = false;
};
template <class P>
struct has_split<P, decltype((void)P::SPLIT_OK)> {
    static constexpr bool value = true;
};
template <class P, class = void>
struct has_spread_q {
    static constexpr bool value = false;
};
template <class P>
struct has_spread_q<P, decltype((void)P::SPREAD_OK)> {
    static constexpr bool value = true;
};
// A one-lane surface spread over L threads per bead.  The start-structure chain of a rate calculation
// (calc_rate.f90:651-1148) is 111 one-bead trajectories one after the other, i.e. a single thread of the GPU at 4 400
// instructions per step, 42 % of them the serial form of xi and its Hessian products, 30 % the per-component loops of the
// step (profiles/r2ai_chain_h3_source.txt).  With L = 16 >= 3 NATOMS lanes every lane owns ONE component: the cooperative
// xi (calc_xi_coop) applies, the loops over the owned components have one trip, Andersen draws its pairs in parallel; the
// surface itself is evaluated by every lane on the same structure (the same instructions the one thread issued, or fewer
// where the surface shares work between the lanes: P::eval_split), lane 0 reports the energy.
// The same with 8, 4 or 2 lanes per bead for trajectories of 2, 4 or 8 beads (the constrained recrossing parent of the
// H + H2 example is ONE 8-bead trajectory of 150 000 steps): lane x of a bead owns the components x NOWN ...
// x NOWN + NOWN - 1, NOWN = ceil(3 NATOMS / L).
template <class P, int L>
struct PesSpread {
    static_assert(P::LANES == 1 && L >= 2 && L <= 32, "spreads a one-lane surface");
    static constexpr int NATOMS = P::NATOMS;
    static constexpr int ID = P::ID;
    static constexpr int LANES = L;
    static constexpr int NOWN = (3 * P::NATOMS + L - 1) / L;
    CRCL_HD static __forceinline__ int owned(int lane, int k)
    {
        const int c = lane * NOWN + k;
        return (k < NOWN && c < 3 * NATOMS) ? c : -1;
    }
    template <class QF>
    CRCL_HD static __forceinline__ int eval_coop(QF qf, int lane, unsigned mask, double& V, double* gown)
    {
        constexpr int NC = 3 * NATOMS;
        double x[NC], g[NC];
#pragma unroll
        for (int c = 0; c < NC; c++) x[c] = qf(c);
        double e;
        int w;
        if constexpr (has_split<P>::value && L >= 4)
            w = P::template eval_split<L>(x, lane, mask, e, g);
        else
            w = P::eval(x, e, g);
        V = (lane == 0) ? e : 0.0;
#pragma unroll
        for (int k = 0; k < NOWN; k++) {
            double own = 0.0;
#pragma unroll
            for (int c = k; c < NC; c++) own = (c == lane * NOWN + k) ? g[c] : own;
            gown[k] = own;
        }
        return w;
    }
};

// The same for a surface that is already split over four lanes per bead (PesCBE4 and its family): R quads of one warp
// evaluate it on the same structure -- quad r, surface lane x owns the ONE component P::owned(x, r) -- so that a
// one-bead trajectory has 4 R >= 3 NATOMS threads and the cooperative xi applies (with the four threads of a bead alone
// every thread ran the serial form over all 18 or 21 components).  Each quad keeps its own block of the surface's
// lane-exchange scratch; quad 0 reports the energy.
template <class P, int R>
struct PesSpreadQ {
    static_assert(P::LANES == 4 && P::NOWN <= R && 4 * R <= 32, "one quad per owned slot, within a warp");
    static constexpr int NATOMS = P::NATOMS;
    static constexpr int ID = P::ID;
    static constexpr int LANES = 4 * R;
    static constexpr int NOWN = 1;
    static constexpr int COOP_SCRATCH = coop_scratch<P>::value;   // per thread: [quad][surface lane][scratch]
    __device__ static __forceinline__ int owned(int lane, int k)
    {
        return (k == 0 && (lane >> 2) < P::NOWN) ? P::owned(lane & 3, lane >> 2) : -1;
    }
    template <class QF>
    __device__ static __forceinline__ int eval_coop(QF qf, int lane, unsigned mask, double& V, double* gown,
                                                   double* scr = nullptr)
    {
        double e, g[P::NOWN];
        int w;
        if constexpr (coop_scratch<P>::value > 0)
            w = P::eval_coop(qf, lane & 3, mask, e, g, scr + (lane >> 2) * 4 * coop_scratch<P>::value);
        else
            w = P::eval_coop(qf, lane & 3, mask, e, g);
        V = ((lane >> 2) == 0) ? e : 0.0;
        double own = 0.0;
#pragma unroll
        for (int k = 0; k < P::NOWN; k++) own = (k == (lane >> 2)) ? g[k] : own;
        gown[0] = own;
        return w;
    }
};

template <class PES, int KIND>
static cudaError_t launch_traj_pes(int nbeads, const TrajArgs& A, int bias_mode, double nose_q,
                                   cudaStream_t s, int* nosup)
{
    *nosup = 0;
    switch (nbeads) {
    case 1:
        if constexpr (PES::LANES == 1 && 3 * PES::NATOMS <= 16) {
            if (A.ntraj <= A.spread_max) return launch_one<PesSpread<PES, 16>, KIND, 1>(A, bias_mode, nose_q, s);
        } else if constexpr (PES::LANES == 4 && 3 * PES::NATOMS <= 32 && has_spread_q<PES>::value) {
            if (A.ntraj <= A.spread_max) return launch_one<PesSpreadQ<PES, 8>, KIND, 1>(A, bias_mode, nose_q, s);
        }
        return launch_one<PES, KIND, 1>(A, bias_mode, nose_q, s);
    case 2:
        if constexpr (PES::LANES == 1 && 3 * PES::NATOMS <= 16) {
            if (2 * A.ntraj <= A.spread_max) return launch_one<PesSpread<PES, 8>, KIND, 2>(A, bias_mode, nose_q, s);
        }
        return launch_one<PES, KIND, 2>(A, bias_mode, nose_q, s);
    case 4:
        if constexpr (PES::LANES == 1 && 3 * PES::NATOMS <= 16) {
            if (4 * A.ntraj <= A.spread_max) return launch_one<PesSpread<PES, 4>, KIND, 4>(A, bias_mode, nose_q, s);
        }
        return launch_one<PES, KIND, 4>(A, bias_mode, nose_q, s);
    case 8:
        if constexpr (PES::LANES == 1 && 3 * PES::NATOMS <= 16 && 3 * PES::NATOMS > 8) {
            if (8 * A.ntraj <= A.spread_max) return launch_one<PesSpread<PES, 2>, KIND, 8>(A, bias_mode, nose_q, s);
        }
        return launch_one<PES, KIND, 8>(A, bias_mode, nose_q, s);
    case 16: return launch_one<PES, KIND, 16>(A, bias_mode, nose_q, s);
    case 32: return launch_one<PES, KIND, 32>(A, bias_mode, nose_q, s);
    case 64: return launch_one<PES, KIND, 64>(A, bias_mode, nose_q, s);
    case 128: return launch_one<PES, KIND, 128>(A, bias_mode, nose_q, s);
    }
    *nosup = 1;
    return cudaSuccess;
}

#define CRCL_DECLARE_TRAJ(name)                                                                   \
    cudaError_t name(int nbeads, const TrajArgs& A, int bias_mode, double nose_q, cudaStream_t s, \
                     int* nosup)
CRCL_DECLARE_TRAJ(launch_h3_verlet);
CRCL_DECLARE_TRAJ(launch_h3_mdinit);
CRCL_DECLARE_TRAJ(launch_h3_recross);
CRCL_DECLARE_TRAJ(launch_oh3_verlet);
CRCL_DECLARE_TRAJ(launch_oh3_mdinit);
CRCL_DECLARE_TRAJ(launch_oh3_recross);
CRCL_DECLARE_TRAJ(launch_ch4h_verlet);
CRCL_DECLARE_TRAJ(launch_ch4h_mdinit);
CRCL_DECLARE_TRAJ(launch_ch4h_recross);
CRCL_DECLARE_TRAJ(launch_brh2_verlet);
CRCL_DECLARE_TRAJ(launch_brh2_mdinit);
CRCL_DECLARE_TRAJ(launch_brh2_recross);
CRCL_DECLARE_TRAJ(launch_o3_verlet);
CRCL_DECLARE_TRAJ(launch_o3_mdinit);
CRCL_DECLARE_TRAJ(launch_o3_recross);
CRCL_DECLARE_TRAJ(launch_ch4oh_verlet);
CRCL_DECLARE_TRAJ(launch_ch4oh_mdinit);
CRCL_DECLARE_TRAJ(launch_ch4oh_recross);
CRCL_DECLARE_TRAJ(launch_geh4oh_verlet);
CRCL_DECLARE_TRAJ(launch_geh4oh_mdinit);
CRCL_DECLARE_TRAJ(launch_geh4oh_recross);
CRCL_DECLARE_TRAJ(launch_ch4cn_verlet);
CRCL_DECLARE_TRAJ(launch_ch4cn_mdinit);
CRCL_DECLARE_TRAJ(launch_ch4cn_recross);
CRCL_DECLARE_TRAJ(launch_clnh3_verlet);
CRCL_DECLARE_TRAJ(launch_clnh3_mdinit);
CRCL_DECLARE_TRAJ(launch_clnh3_recross);
CRCL_DECLARE_TRAJ(launch_nh3oh_verlet);
CRCL_DECLARE_TRAJ(launch_nh3oh_mdinit);
CRCL_DECLARE_TRAJ(launch_nh3oh_recross);
CRCL_DECLARE_TRAJ(launch_h2co_verlet);
CRCL_DECLARE_TRAJ(launch_h2co_mdinit);
CRCL_DECLARE_TRAJ(launch_h2co_recross);

}  // namespace crcl
