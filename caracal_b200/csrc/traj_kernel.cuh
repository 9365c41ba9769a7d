// traj_kernel.cuh -- fused ring-polymer trajectory kernels for small analytic systems.
//
// One launch advances a whole batch of independent ring polymers by many steps; positions,
// forces and all per-trajectory scalars stay in registers / shared memory for the whole
// launch, so HBM traffic per step is zero (state is read once and written once).
//
// Replaces, per step, the reference's verlet.f90:65-1308 (operation order: SURVEY.md 3.5)
// with everything it calls on the graded paths: rfft/irfft (rfft.f90, irfft.f90),
// get_centroid.f90, gradient.f90 -> egrad_*, umbrella.f90 + calc_xi.f90, constrain_q.f90,
// constrain_p.f90, nhc.f90, andersen.f90, transrot.f90 + invert.f90; and, for the
// recrossing work unit, the child-pair body recross.f90:515-628 / recross_serial.f90:172-229.
//
// Mapping: one thread per (trajectory, bead).  The NB beads of a trajectory form a
// "group": a sub-warp segment for NB <= 32 (32/NB trajectories per warp, synchronised with
// masked __syncwarp / shuffles, so trajectories in the same warp never wait on each other),
// the whole CTA for NB > 32.
//
// Free ring-polymer step (verlet.f90:377-463).  rfft and irfft are the same map
// T = Re(DFT)/sqrt(N), so the reference applies T.D.T with D the per-mode
// 2x2 propagator; T.T = (I+J)/2 with J the bead reversal a -> N-a, hence
//     T.D.T = Circ(f) . (I+J)/2 ,   f(j) = (1/N) sum_k d_k cos(2 pi k j / N),
// i.e. "symmetrise over a <-> N-a, then apply the exact circulant propagator".  The three
// kernels f_c (cos wt), f_a (-w sin wt) and f_b (sin wt / w) depend only on (N, beta, dt)
// and are built once on the host (api.cu); masses enter as p' = Fc p + m Fa q,
// q' = Fb p / m + Fc q.  CRCL_TRANSFORM_EXACT skips the symmetrisation.
#pragma once
#include "crcl_common.cuh"
#include "rng.cuh"
#include "xi.cuh"

namespace crcl {

constexpr int TRAJ_MAXNAT = XI_MAXAT;

struct TrajArgs {
    int ntraj, nsteps, istep0, constrain, thermostat, andersen_step, symmetrize, nbeads;
    int spread_max;   // host side only: largest batch (trajectories x beads) that runs in the spread forms (traj_inst.cuh)
    double beta, dt, kelvin;
    double mass[TRAJ_MAXNAT];
    int at_move[TRAJ_MAXNAT];
    Mech mech;
    // per-trajectory inputs; if the pointer is null the scalar is used for every trajectory
    const double* xi_ideal;
    const double* k_force;
    double xi_ideal_s, k_force_s;
    // state, reference layout [traj][bead][atom][xyz]
    double* q;
    double* p;
    double* g;
    double* dxi;      // [traj][atom][xyz] (in/out)
    double* epot;     // [traj]
    double* xi_real;  // [traj]
    int* status;      // [traj]
    double* nhc;      // [traj][8]: vnh[4], qnh[4]
    const uint32_t* traj_id;
    uint32_t traj_id0;
    uint32_t* event;  // [traj]
    uint64_t seed;
    const double* fker;  // [3][NB]
    // umbrella sampling accumulators (may be null): sum xi, sum xi^2 over the steps of this launch
    double* xi_sum;
    double* xi_sum2;
    // recrossing work unit
    const double* q_parents;  // [nparent][bead][atom][xyz]
    int nparent, pair0;
    unsigned char* theta;  // [nsteps][ntraj]: xi_real > 0 after step l
    double* weight;        // [ntraj]: v_s/f_s of the child
    double* denom_part;    // [ntraj]: v_s/f_s if v_s > 0 else 0
    // rpmd_check.f90:100-116 after every step with constrain 0 / 1 / 3 (crcl_set_rpmd_check)
    int chk_on;
    double chk_emax;       // (ts_energy + energy_tol) * nbeads
    double chk_xi_tol;
};

// The T = NB*LANES threads of one trajectory: a sub-warp segment when T <= 32 (32/T trajectories
// per warp, masked shuffles / __syncwarp so trajectories sharing a warp never wait on each
// other), otherwise one whole CTA.
// Sub-warp trajectories are packed CRCL_WTPB threads to a CTA and the CTA's warps are re-aligned with one barrier per
// step: the step body of the unrolled surfaces is ~90 KB of SASS, and single-warp CTAs drifting through it
// independently were bound by instruction-cache misses (ncu, H + H2 x 16 beads: no_instruction 7.4 of 13 stall
// cycles per issued instruction, profiles/r1m_verlet_h3_nb16_minb12.txt); warps in step share the fetched lines.
#ifndef CRCL_WTPB
#define CRCL_WTPB 128
#endif
// Multi-warp trajectories (T > 32) are packed up to CRCL_CTPB threads to a CTA for the same reason; each trajectory
// then synchronises on its own named barrier (bar.sync 1 + gib, T), because SHAKE's iteration count and the failure
// paths are per trajectory, and barrier 0 re-aligns the whole CTA once per step.
#ifndef CRCL_CTPB
#define CRCL_CTPB 128
#endif
template <int NB, int LANES>
struct Group {
    static constexpr int T = NB * LANES;
    static constexpr bool WARP = (T <= 32);
    static constexpr int CPB = (CRCL_CTPB / T > 1) ? (CRCL_CTPB / T < 15 ? CRCL_CTPB / T : 15) : 1;   // 15 named barriers
    static constexpr int TPB = WARP ? CRCL_WTPB : T * CPB;  // threads per block
    static constexpr int RED_N = 9;                          // values per warp in the reduction scratch (sum_n)
    static constexpr int GPB = WARP ? CRCL_WTPB / T : CPB;  // trajectories per block
    // re-align the warps of a CTA that holds several trajectories (one barrier per step)
    static __device__ __forceinline__ void align_warps()
    {
        if (WARP ? (CRCL_WTPB > 32) : (GPB > 1)) __syncthreads();
    }
    int tig, bead, lane, gib;                      // thread in group, bead, lane of the bead, group in block
    unsigned mask;
    double* red;  // CTA-wide scratch (T > 32): [T/32]

    __device__ __forceinline__ Group(double* red_)
    {
        tig = threadIdx.x % T;
        bead = tig / LANES;
        lane = tig % LANES;
        gib = threadIdx.x / T;
        if (WARP) {
            const unsigned m = (T == 32) ? 0xffffffffu : ((1u << T) - 1u);
            mask = m << ((threadIdx.x & 31) / T * T);
        } else {
            mask = 0xffffffffu;
        }
        red = red_;
    }
    __device__ __forceinline__ void sync() const
    {
        if (WARP)
            __syncwarp(mask);
        else if (GPB == 1)
            __syncthreads();
        else
            asm volatile("bar.sync %0, %1;" ::"r"(gib + 1), "r"(T) : "memory");
    }
    // all-reduce sum over the threads of the trajectory; every thread gets the same bits
    __device__ __forceinline__ double sum(double v) const
    {
        if (WARP) {
#pragma unroll
            for (int o = T / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
            return v;
        } else {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            sync();
            if ((threadIdx.x & 31) == 0) red[(threadIdx.x >> 5) * RED_N] = v;   // red is indexed by the warp's number in the CTA
            sync();
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < T / 32; w++) t += red[(gib * (T / 32) + w) * RED_N];
            return t;
        }
    }
    // the same all-reduce for N values at once: every value goes through the same shuffle tree and the same order over the
    // warps as in sum() (identical bits), but a multi-warp trajectory pays ONE pair of barriers instead of N
    template <int N>
    __device__ __forceinline__ void sum_n(double (&v)[N]) const
    {
        static_assert(N <= RED_N, "enlarge the reduction scratch");
        if (WARP) {
#pragma unroll
            for (int i = 0; i < N; i++)
#pragma unroll
                for (int o = T / 2; o > 0; o >>= 1) v[i] += __shfl_xor_sync(mask, v[i], o);
        } else {
#pragma unroll
            for (int i = 0; i < N; i++)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
            sync();
            if ((threadIdx.x & 31) == 0) {
#pragma unroll
                for (int i = 0; i < N; i++) red[(threadIdx.x >> 5) * RED_N + i] = v[i];
            }
            sync();
#pragma unroll
            for (int i = 0; i < N; i++) {
                double t = 0.0;
#pragma unroll
                for (int w = 0; w < T / 32; w++) t += red[(gib * (T / 32) + w) * RED_N + i];
                v[i] = t;
            }
        }
    }
    __device__ __forceinline__ int any(int pred) const
    {
        if (WARP)
            return __any_sync(mask, pred);
        else if (GPB == 1)
            return __syncthreads_or(pred);
        else {
            const int w = __any_sync(0xffffffffu, pred);
            sync();
            if ((threadIdx.x & 31) == 0) red[(threadIdx.x >> 5) * RED_N] = w ? 1.0 : 0.0;
            sync();
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < T / 32; k++) t += red[(gib * (T / 32) + k) * RED_N];
            return t != 0.0;
        }
    }
};

// Free ring-polymer step on the FP64 tensor cores (mma.sync.m8n8k4.f64) for the lane-split surfaces whose trajectories are
// whole warps with dense tables (NB = 8, 16, 32 at four lanes per bead; Traj::free_rp_dmma).  -DCRCL_DMMA_TRANSFORM=0
// restores the FMA loop (A/B builds: profiles/build_variant.py).
#ifndef CRCL_DMMA_TRANSFORM
#define CRCL_DMMA_TRANSFORM 1
#endif

// doubles of lane-exchange scratch per thread a lane-split surface asks for (PES::COOP_SCRATCH), 0 if it has none
template <class P, class = void>
struct coop_scratch {
    static constexpr int value = 0;
};
template <class P>
struct coop_scratch<P, decltype((void)P::COOP_SCRATCH)> {
    static constexpr int value = P::COOP_SCRATCH;
};

// shared-memory footprint (doubles) of one trajectory and of the block-wide part; SCR: coop_scratch of the surface
template <int NAT, int NB, int LANES, int SCR = 0>
struct SmemLayout {
    static constexpr int NC = 3 * NAT;
    // bead stride of the {p,q}[c][b] staging in double2 units: +1 when several lanes of a bead
    // address different components of the same bead slot (keeps them in different banks)
    static constexpr int NBP = (LANES > 1 && NB > 1) ? NB + 1 : NB;
    // staging of the bead-symmetrised {p,q} sums s_b = x_b + x_{N-b}, b = 0..N/2 (reference transform, NB <= 32)
    static constexpr int NH = NB / 2 + 1;
    static constexpr int SYM_STAGE = (NB > 1) ? 2 * NC * NH : 0;
    static constexpr int XI_SCR = (3 * NC + 2) & ~1;        // calc_xi_coop scratch (ds0, ds1, v, xi), even for double2 alignment
    static constexpr int COOP = SCR * NB * LANES;           // lane-exchange scratch of the surface, [bead][lane][SCR]
    static constexpr int PER_GROUP = 2 * NC * NBP + 4 * NC + SYM_STAGE + XI_SCR + COOP;  // {p,q}[c][b], cen, dxi, add, ham, sym, xi scratch, lane exchange
    // free-RP kernels: for NB <= 32 the three N x N tables H[b][a] (symmetrisation folded in, see
    // load_fker), otherwise the three circulant kernels f[N]
    static constexpr bool HTAB = (NB > 1 && NB <= 32);
    static constexpr int FKER = HTAB ? 3 * NB * NB : 3 * NB;
    // after the tables and the reduction scratch (RED_N values for each of up to 32 warps): {mass, 1/mass}[lane][NOWN] (double2), see Traj::mt
    static constexpr int MT_OFF = (FKER + 32 * Group<NB, LANES>::RED_N + 1) & ~1;
    static constexpr int MT_MAX = 2 * 3 * XI_MAXAT + 16;         // >= 2 LANES NOWN for every surface (static_assert in load_fker)
    // {mass, 1/mass}[component] (double2) for the tensor-core form of the free ring-polymer step, whose accumulator
    // fragments hold components the thread does not own
    static constexpr bool DMMA = CRCL_DMMA_TRANSFORM && HTAB && LANES > 1 && (NB % 8 == 0) && ((NB * LANES) % 32 == 0);
    static constexpr int MC_OFF = MT_OFF + MT_MAX;
    static constexpr int BLOCK = MC_OFF + (DMMA ? 2 * NC : 0);   // even: double2 alignment of what follows
    static constexpr size_t bytes()
    {
        return sizeof(double) * (BLOCK + Group<NB, LANES>::GPB * PER_GROUP);
    }
};

// Per-thread view of a trajectory.  Every thread owns up to NOWN components c = atom*3+xyz of
// its bead (all of them when LANES == 1); momenta and positions of the whole trajectory live in
// shared memory as {p,q}[c][bead], the forces of the owned components in registers.
template <class PES, int NB>
struct Traj {
    static constexpr int NAT = PES::NATOMS;
    static constexpr int NC = 3 * NAT;
    static constexpr int L = PES::LANES;
    static constexpr int NO = PES::NOWN;
    static constexpr int NBP = SmemLayout<NAT, NB, L>::NBP;
    using Grp = Group<NB, L>;
    const TrajArgs& A;
    const Grp& G;
    double g[NO];   // forces of the owned components
    const double2* mt;  // shared {mass, 1/mass} of the owned components' atoms, [lane][NO] (load_fker): re-read where it is
                        // needed instead of 4 NO registers that would stay live across the surface
    int oc[NO];     // owned component numbers (-1: none)
    int ob[NO];     // their row offset in pq (component 0's row for unowned slots)
    unsigned mv;    // bit k: owned component k exists and its atom is movable
    double2* pq;    // shared {p,q}[c][b]
    double* cen;    // shared centroid [NC]
    double* dxi;    // shared dxi [NC]
    double* add;    // shared k (xi - xi0) dxi, the umbrella force added to every bead [NC]
    double* ham;    // shared hams force (umbrella.f90:144-174) [NC]
    double2* sq;    // shared bead-symmetrised sums {p,q}[c][b], b = 0..N/2 (free_rp, reference transform)
    double* xis;    // shared scratch of calc_xi_coop [3 NC]
    double* coop;   // shared lane-exchange scratch of this thread's bead (surfaces with COOP_SCRATCH), else unused
    bool want_epot; // false: forces() skips the all-reduce of the bead energies (recrossing children)
    const double* fk;
    const double2* mc;  // shared {mass, 1/mass}[component] (SmemLayout::DMMA only)
    double mass_sum;    // sum of the atomic masses in atom order (transrot.f90:87-92)
    double xi_ideal, k_force, xi_real, epot;
    double vnh[4], qnh[4];
    int nfree, status;
    uint32_t tid, event;

    __device__ __forceinline__ Traj(const TrajArgs& a, const Grp& grp, double* smem) : A(a), G(grp)
    {
        using Lay = SmemLayout<NAT, NB, L, coop_scratch<PES>::value>;
        fk = smem;
        mc = reinterpret_cast<const double2*>(smem + Lay::MC_OFF);
        mt = reinterpret_cast<const double2*>(smem + Lay::MT_OFF) + grp.lane * NO;
        double* base = smem + Lay::BLOCK + grp.gib * Lay::PER_GROUP;
        pq = reinterpret_cast<double2*>(base);
        cen = base + 2 * NC * NBP;
        dxi = cen + NC;
        add = dxi + NC;
        ham = add + NC;
        sq = reinterpret_cast<double2*>(ham + NC);
        xis = ham + NC + Lay::SYM_STAGE;
        coop = xis + Lay::XI_SCR + grp.bead * (L * coop_scratch<PES>::value);
        want_epot = true;
        mass_sum = 0.0;
#pragma unroll
        for (int j = 0; j < NAT; j++) mass_sum += A.mass[j];
        status = 0;
        xi_real = 0.0;
        epot = 0.0;
        nfree = 0;
#pragma unroll
        for (int j = 0; j < NAT; j++)
            if (A.at_move[j]) nfree += 3;
        mv = 0u;
#pragma unroll
        for (int k = 0; k < NO; k++) {
            oc[k] = PES::owned(grp.lane, k);
            ob[k] = (oc[k] >= 0 ? oc[k] : 0) * NBP;
            if (oc[k] >= 0 && A.at_move[oc[k] / 3]) mv |= 1u << k;
            g[k] = 0.0;
        }
    }
    __device__ __forceinline__ double& P(int c) { return pq[c * NBP + G.bead].x; }
    __device__ __forceinline__ double& Q(int c) { return pq[c * NBP + G.bead].y; }
    __device__ __forceinline__ double& Pk(int k) { return pq[ob[k] + G.bead].x; }
    __device__ __forceinline__ double& Qk(int k) { return pq[ob[k] + G.bead].y; }
    // threads that hold a valid xi_real after a child step (umbrella mode 2), and the one that reports it
    __device__ __forceinline__ bool xi_thread() const { return Grp::WARP || G.tig >= Grp::T - 32; }
    __device__ __forceinline__ bool xi_writer() const { return Grp::WARP ? G.tig == 0 : G.tig == Grp::T - 32; }
    __device__ __forceinline__ bool own(int k) const { return oc[k] >= 0; }
    __device__ __forceinline__ bool mov(int k) const { return (mv >> k) & 1u; }

    // p <- p - dt/2 * g (verlet.f90:216-218, 1060-1062) followed by the fixed-atom mask (:225-231)
    __device__ __forceinline__ void half_kick()
    {
        const double h = 0.5 * A.dt;
#pragma unroll
        for (int k = 0; k < NO; k++)
            if (own(k)) Pk(k) = mov(k) ? Pk(k) - h * g[k] : 0.0;
    }
    // the second half kick of one step and the first of the next in one pass over the momenta (the same two operations in
    // the same order, without the store and reload between them)
    __device__ __forceinline__ void double_half_kick()
    {
        const double h = 0.5 * A.dt;
#pragma unroll
        for (int k = 0; k < NO; k++)
            if (own(k)) Pk(k) = mov(k) ? (Pk(k) - h * g[k]) - h * g[k] : 0.0;
    }
    __device__ __forceinline__ void mask_p()
    {
#pragma unroll
        for (int k = 0; k < NO; k++)
            if (own(k) && !mov(k)) Pk(k) = 0.0;
    }
    // centroid of q (get_centroid.f90:67-82) into shared cen[]; beads summed in order
    __device__ __forceinline__ void centroid()
    {
        G.sync();
        for (int c = G.tig; c < NC; c += Grp::T) {
            double s = 0.0;
            for (int b = 0; b < NB; b++) s += pq[c * NBP + b].y;
            cen[c] = s / NB;
        }
        G.sync();
    }
    // The same step on the FP64 tensor cores: out[a][c] = sum_b H[a][b] x[b][c] as m8n8k4 tiles -- 8 beads x 8 components
    // per accumulator pair, A fragments = the three dense tables (stored in fragment order by load_fker: one conflict-free
    // 8-byte load per fragment), B fragments = {p,q} of 4 source beads x 8 components (one 16-byte load yields the P and the
    // Q fragment).  Warp w of the trajectory takes row tile w mod MT and every (T/32/MT)-th column tile; its A fragments
    // are loaded once per step.  No staging of bead-symmetrised sums: the symmetrisation is in the table.
    // Measured against the FMA loop on the transform alone (profiles/r2u_bench_transform.json): 2.2x at 16 beads.
    __device__ __forceinline__ void free_rp_dmma()
    {
        constexpr int WPT = Grp::T / 32, MT = NB / 8, KS = NB / 4, NT = (NC + 7) / 8;
        constexpr int WPM = (WPT >= MT) ? WPT / MT : 1;            // warps that share a row tile (split its column tiles)
        static_assert(WPT % MT == 0 || MT % WPT == 0, "row tiles and warps of a trajectory must divide each other");
        const int w = G.tig >> 5, lane = G.tig & 31, gr = lane >> 2, tg = lane & 3;
        auto mma = [](double& c0, double& c1, double a, double b) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
        };
        G.sync();
        constexpr int RT = (MT >= WPT) ? MT / WPT : 1;             // row tiles per warp
        constexpr int CT = (NT + WPM - 1) / WPM;                   // column tiles per warp
        double2 res[RT][CT][2];
#pragma unroll
        for (int ir = 0; ir < RT; ir++) {
            const int mt_ = (MT >= WPT) ? w + WPT * ir : w % MT;
            double fc[KS], fa[KS], fb[KS];
            const double* hf = fk + mt_ * (KS * 32) + lane;
#pragma unroll
            for (int ks = 0; ks < KS; ks++) {
                fc[ks] = hf[ks * 32];
                fa[ks] = hf[MT * KS * 32 + ks * 32];
                fb[ks] = hf[2 * MT * KS * 32 + ks * 32];
            }
#pragma unroll
            for (int ic = 0; ic < CT; ic++) {
                const int nt = (WPT >= MT) ? (w / MT) + WPM * ic : ic;
                const int c = nt * 8 + gr;
                const double2* xb = pq + ((c < NC) ? c : NC - 1) * NBP + tg;   // padded columns: a copy of the last one
                double cp[2] = {0.0, 0.0}, aq[2] = {0.0, 0.0}, bp[2] = {0.0, 0.0}, cq[2] = {0.0, 0.0};
#pragma unroll
                for (int ks = 0; ks < KS; ks++) {
                    const double2 u = xb[ks * 4];
                    mma(cp[0], cp[1], fc[ks], u.x);
                    mma(aq[0], aq[1], fa[ks], u.y);
                    mma(bp[0], bp[1], fb[ks], u.x);
                    mma(cq[0], cq[1], fc[ks], u.y);
                }
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int j = nt * 8 + 2 * tg + e;
                    const double2 m = mc[(j < NC) ? j : NC - 1];
                    res[ir][ic][e] = make_double2(fma(m.x, aq[e], cp[e]), fma(m.y, bp[e], cq[e]));
                }
            }
        }
        G.sync();   // every B fragment has been read
#pragma unroll
        for (int ir = 0; ir < RT; ir++) {
            const int mt_ = (MT >= WPT) ? w + WPT * ir : w % MT;
#pragma unroll
            for (int ic = 0; ic < CT; ic++) {
                const int nt = (WPT >= MT) ? (w / MT) + WPM * ic : ic;
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int j = nt * 8 + 2 * tg + e;
                    if (j < NC && nt < NT) pq[j * NBP + mt_ * 8 + gr] = res[ir][ic][e];
                }
            }
        }
        // the caller's next phase starts with G.sync() (centroid), which publishes these stores
    }
    // free ring-polymer propagation (verlet.f90:353-357 for one bead, :377-463 otherwise):
    // p' = Fc p + m Fa q, q' = Fb p / m + Fc q with the circulant kernels of the header comment
    __device__ __forceinline__ void free_rp()
    {
        if (NB == 1) {
#pragma unroll
            for (int k = 0; k < NO; k++)
                if (own(k)) Qk(k) = Qk(k) + CRCL_DIV(Pk(k) * A.dt, mt[k].x);   // branch-free quotient: the nine overlap
            return;
        }
        if constexpr (SmemLayout<NAT, NB, L>::DMMA) {
            free_rp_dmma();
            return;
        }
        G.sync();
        if (SmemLayout<NAT, NB, L>::HTAB) {
            // out_a = sum_b H[b][a] x_b with H = Circ(f) (I+J)/2 (or Circ(f)) tabulated by load_fker
            double cp[NO], aq[NO], bp[NO], cq[NO];
#pragma unroll
            for (int k = 0; k < NO; k++) cp[k] = aq[k] = bp[k] = cq[k] = 0.0;
            const double* hk = fk + G.bead;
            if (A.symmetrize) {
                // H[b][a] = H[N-b][a]: out_a = sum_{b=0}^{N/2} H[b][a] s_b with s_b = x_b + x_{N-b}
                // (s_0 = x_0, s_{N/2} = x_{N/2}) staged once per trajectory -- N/2+1 terms instead of N
                constexpr int NH = SmemLayout<NAT, NB, L>::NH;
                for (int i = G.tig; i < NC * NH; i += Grp::T) {
                    const int c = i / NH, b = i - c * NH;
                    double2 u = pq[c * NBP + b];
                    if (b > 0 && 2 * b < NB) {
                        const double2 w = pq[c * NBP + NB - b];
                        u.x += w.x;
                        u.y += w.y;
                    }
                    sq[i] = u;
                }
                G.sync();
#pragma unroll 3
                for (int b = 0; b < NH; b++) {
                    const double fc = hk[b * NB], fa = hk[NB * NB + b * NB], fb = hk[2 * NB * NB + b * NB];
#pragma unroll
                    for (int k = 0; k < NO; k++) {
                        const double2 u = sq[(ob[k] / NBP) * NH + b];
                        cp[k] = fma(fc, u.x, cp[k]);
                        aq[k] = fma(fa, u.y, aq[k]);
                        bp[k] = fma(fb, u.x, bp[k]);
                        cq[k] = fma(fc, u.y, cq[k]);
                    }
                }
                // every read of pq happened before the barrier above: no second barrier before the write
            } else {
#pragma unroll 2
                for (int b = 0; b < NB; b++) {
                    const double fc = hk[b * NB], fa = hk[NB * NB + b * NB], fb = hk[2 * NB * NB + b * NB];
#pragma unroll
                    for (int k = 0; k < NO; k++) {
                        const double2 u = pq[ob[k] + b];
                        cp[k] = fma(fc, u.x, cp[k]);
                        aq[k] = fma(fa, u.y, aq[k]);
                        bp[k] = fma(fb, u.x, bp[k]);
                        cq[k] = fma(fc, u.y, cq[k]);
                    }
                }
                G.sync();
            }
#pragma unroll
            for (int k = 0; k < NO; k++)
                if (own(k)) pq[ob[k] + G.bead] = make_double2(fma(mt[k].x, aq[k], cp[k]), fma(mt[k].y, bp[k], cq[k]));
            return;
        }
        double cp[NO], aq[NO], bp[NO], cq[NO];
#pragma unroll
        for (int k = 0; k < NO; k++) cp[k] = aq[k] = bp[k] = cq[k] = 0.0;
        if (A.symmetrize) {
            // NB > 32: no room for the dense tables; same half-sum form with the coefficients
            // (f(a-b) + f(a+b))/2 built on the fly from the circulant kernels
            constexpr int NH = SmemLayout<NAT, NB, L>::NH;
            for (int i = G.tig; i < NC * NH; i += Grp::T) {
                const int c = i / NH, b = i - c * NH;
                double2 u = pq[c * NBP + b];
                if (b > 0 && 2 * b < NB) {
                    const double2 w = pq[c * NBP + NB - b];
                    u.x += w.x;
                    u.y += w.y;
                }
                sq[i] = u;
            }
            G.sync();
#pragma unroll 2
            for (int b = 0; b < NH; b++) {
                const int i1 = (G.bead - b) & (NB - 1), i2 = (G.bead + b) & (NB - 1);
                const double fc = 0.5 * (fk[i1] + fk[i2]), fa = 0.5 * (fk[NB + i1] + fk[NB + i2]),
                             fb = 0.5 * (fk[2 * NB + i1] + fk[2 * NB + i2]);
#pragma unroll
                for (int k = 0; k < NO; k++) {
                    const double2 u = sq[(ob[k] / NBP) * NH + b];
                    cp[k] = fma(fc, u.x, cp[k]);
                    aq[k] = fma(fa, u.y, aq[k]);
                    bp[k] = fma(fb, u.x, bp[k]);
                    cq[k] = fma(fc, u.y, cq[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < NO; k++)
                if (own(k)) pq[ob[k] + G.bead] = make_double2(fma(mt[k].x, aq[k], cp[k]), fma(mt[k].y, bp[k], cq[k]));
            return;
        }
#pragma unroll 2
        for (int b = 0; b < NB; b++) {
            const int idx = (G.bead - b) & (NB - 1);
            const double fc = fk[idx], fa = fk[NB + idx], fb = fk[2 * NB + idx];
#pragma unroll
            for (int k = 0; k < NO; k++) {
                const double2 u = pq[ob[k] + b];
                cp[k] = fma(fc, u.x, cp[k]);
                aq[k] = fma(fa, u.y, aq[k]);
                bp[k] = fma(fb, u.x, bp[k]);
                cq[k] = fma(fc, u.y, cq[k]);
            }
        }
        G.sync();
#pragma unroll
        for (int k = 0; k < NO; k++)
            if (own(k)) pq[ob[k] + G.bead] = make_double2(fma(mt[k].x, aq[k], cp[k]), fma(mt[k].y, bp[k], cq[k]));
    }
    // gradient.f90 -> egrad_<pes> for this bead; returns epot = sum over beads (verlet.f90:772-777)
    __device__ __forceinline__ double forces()
    {
        G.sync();
        double e;
        const double2* base = pq + G.bead;
        int w;
        if constexpr (coop_scratch<PES>::value > 0)
            w = PES::eval_coop([&](int c) { return base[c * NBP].y; }, G.lane, G.mask, e, g, coop);
        else
            w = PES::eval_coop([&](int c) { return base[c * NBP].y; }, G.lane, G.mask, e, g);
        if (w) status |= CRCL_TRAJ_PESWARN;
        return want_epot ? G.sum(e) : 0.0;
    }
    // umbrella.f90:66-175 on the shared centroid.  mode 0: bias + hams force added to the forces,
    // xi in umbrella form; mode 1: xi in recrossing form, nothing added.
    __device__ __forceinline__ void umbrella(int mode)
    {
        if (mode == 2) {
            // child trajectory: only the value of xi is ever used (recross.f90:597-602); read the
            // shared centroid in place
            // shared centroid in place.  One thread of the trajectory consumes xi_real (theta), so for
            // CTA-wide groups only the LAST warp evaluates it -- the first one has just done the
            // centroid sums, which keeps the warps of a CTA in step at the next barrier.
            if (xi_thread()) xi_real = xi_value<NAT>(A.mech, cen, xi_ideal, 2);
            return;
        }
        // xi, dxi (and the hams force) cooperatively: thread t < NC gathers component t (xi.cuh, calc_xi_coop);
        // results go straight to the shared dxi / ham
        auto gsync = [&]() { G.sync(); };
        if (mode == 1) {
            xi_real = calc_xi_coop<NAT, Grp::T>(A.mech, A.mass, cen, xi_ideal, 2, false, A.beta, G.tig, gsync, xis, dxi, ham);
        } else {
            xi_real = calc_xi_coop<NAT, Grp::T>(A.mech, A.mass, cen, xi_ideal, 1, true, A.beta, G.tig, gsync, xis, dxi, ham);
            const double kd = k_force * (xi_real - xi_ideal);
            G.sync();
#pragma unroll
            for (int k = 0; k < NO; k++)
                if (own(k)) g[k] = (g[k] + kd * dxi[oc[k]]) + ham[oc[k]];
        }
        G.sync();
    }
    // constrain_q.f90:30-112 (SHAKE on the centroid with the previous step's dxi).  The trial structure lives in
    // the shared `add` slot, the gradient of xi on it in `ham` (both free in the constrained mode); calc_xi_coop
    // evaluates it across the threads of the trajectory; every thread forms the same dsigma in the reference's order.
    __device__ __forceinline__ int constrain_q()
    {
        auto gsync = [&]() { G.sync(); };
        const double dt = A.dt;
        double mult = 0.0, coeff = 0.0;
        int ok = 0;
        for (int iter = 1; iter <= 200; iter++) {
            coeff = mult * dt * dt / NB;
            G.sync();
            for (int c = G.tig; c < NC; c += Grp::T) add[c] = cen[c] + coeff * dxi[c] / A.mass[c / 3];
            G.sync();
            const double xin =
                calc_xi_coop<NAT, Grp::T>(A.mech, A.mass, add, xi_ideal, 2, false, A.beta, G.tig, gsync, xis, ham, ham);
            G.sync();
            double dsigma = 0.0;
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
                for (int j = 0; j < NAT; j++)
                    dsigma += ham[3 * j + k] * dt * dt * dxi[3 * j + k] / (A.mass[j] * NB);
            const double dx = xin / dsigma;
            mult -= dx;
            // 1.0E-8 / 1.0E-10 are REAL*4 literals in constrain_q.f90:93
            if (fabs(dx) < FL(1.0E-8) || fabs(xin) < FL(1.0E-10)) {
                ok = 1;
                break;
            }
        }
        if (!ok) return 1;
#pragma unroll
        for (int k = 0; k < NO; k++)
            if (own(k)) {
                Qk(k) = Qk(k) + coeff / mt[k].x * dxi[oc[k]];
                Pk(k) = Pk(k) + mult * dt / NB * dxi[oc[k]];
            }
        return 0;
    }
    // constrain_p.f90:30-75 (RATTLE)
    __device__ __forceinline__ void constrain_p()
    {
        double c1 = 0.0, c2 = 0.0;
#pragma unroll
        for (int k = 0; k < NO; k++)
            if (own(k)) c1 += dxi[oc[k]] * Pk(k) / mt[k].x;
#pragma unroll
        for (int j = 0; j < NAT; j++)
#pragma unroll
            for (int k = 0; k < 3; k++) c2 += dxi[3 * j + k] * dxi[3 * j + k] / A.mass[j];
        c1 = G.sum(c1);
        const double lam = -c1 / c2 / NB;
#pragma unroll
        for (int k = 0; k < NO; k++)
            if (own(k)) Pk(k) = Pk(k) + lam * dxi[oc[k]];
    }
    // andersen.f90:36-74: full resample p = N(0,1) sqrt(m/beta_n); component m of the bead takes
    // element m&1 of Box-Muller pair m>>1 (rng.cuh)
    __device__ __forceinline__ void andersen()
    {
        const double beta_n = A.beta / NB;
#pragma unroll
        for (int k = 0; k < NO; k++)
            if (oc[k] >= 0) {
                const int m = oc[k];
                double z0, z1;
                normal_pair(A.seed, tid, event, (uint32_t)G.bead, (uint32_t)(m >> 1), z0, z1);
                Pk(k) = ((m & 1) ? z1 : z0) * sqrt(mt[k].x / beta_n);
            }
        event++;
    }
    // nhc.f90:34-170
    __device__ __forceinline__ void nhc()
    {
        constexpr float ektf = 1.380649E-23f / 4.3597447E-18f;  // REAL*4 division, nhc.f90:51
        const double ekt = (double)ektf * A.kelvin;
        const double dtc = A.dt / 5.0;
        double w[3];
        w[0] = 1.0 / (2.0 - cbrt(2.0));
        w[1] = 1.0 - 2.0 * w[0];
        w[2] = w[0];
        double ek = 0.0;
#pragma unroll
        for (int k = 0; k < NO; k++)
            if (mov(k)) ek += Pk(k) * Pk(k) / (2.0 * mt[k].x) / NB / NB;
        double eksum = G.sum(ek);
        double scale = 1.0, gn;
        const double nf = (double)nfree;
        for (int i = 0; i < 5; i++)
            for (int j = 0; j < 3; j++) {
                const double dts = w[j] * dtc, dt2 = 0.5 * dts, dt4 = 0.25 * dts, dt8 = 0.125 * dts;
                double ex;
                gn = (qnh[2] * vnh[2] * vnh[2] - ekt) / qnh[3];
                vnh[3] = vnh[3] + gn * dt4;
                gn = (qnh[1] * vnh[1] * vnh[1] - ekt) / qnh[2];
                ex = exp(-vnh[3] * dt8);
                vnh[2] = ex * (vnh[2] * ex + gn * dt4);
                gn = (qnh[0] * vnh[0] * vnh[0] - ekt) / qnh[1];
                ex = exp(-vnh[2] * dt8);
                vnh[1] = ex * (vnh[1] * ex + gn * dt4);
                gn = (2.0 * eksum - nf * ekt) / qnh[0];
                ex = exp(-vnh[1] * dt8);
                vnh[0] = ex * (vnh[0] * ex + gn * dt4);
                ex = exp(-vnh[0] * dt2);
                scale = scale * ex;
                eksum = eksum * ex * ex;
                gn = (2.0 * eksum - nf * ekt) / qnh[0];
                ex = exp(-vnh[1] * dt8);
                vnh[0] = ex * (vnh[0] * ex + gn * dt4);
                gn = (qnh[0] * vnh[0] * vnh[0] - ekt) / qnh[1];
                ex = exp(-vnh[2] * dt8);
                vnh[1] = ex * (vnh[1] * ex + gn * dt4);
                gn = (qnh[1] * vnh[1] * vnh[1] - ekt) / qnh[2];
                ex = exp(-vnh[3] * dt8);
                vnh[2] = ex * (vnh[2] * ex + gn * dt4);
                gn = (qnh[2] * vnh[2] * vnh[2] - ekt) / qnh[3];
                vnh[3] = vnh[3] + gn * dt4;
            }
#pragma unroll
        for (int k = 0; k < NO; k++)
            if (own(k)) Pk(k) = mov(k) ? scale * Pk(k) : 0.0;
    }
    // NHC masses and zeroed chain (mdinit.f90:126-146)
    __device__ __forceinline__ void nhc_init(double nose_q)
    {
        const double ekt = 0.316679e-5 * A.kelvin;
        const double qterm = ekt * nose_q * nose_q;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            qnh[j] = qterm;
            vnh[j] = 0.0;
        }
        qnh[0] = (double)nfree * qnh[0];
    }
    // transrot.f90:36-236, including the totmass*nbeads double count.  A thread
    // contributes, per owned component (atom j, direction d) with velocity v:
    //   linear momentum m v e_d, angular momentum m (q x e_d) v; the thread owning (j,0) also adds
    //   atom j's m q and inertia terms.
    __device__ __forceinline__ int transrot()
    {
        G.sync();
        double s[15];
#pragma unroll
        for (int i = 0; i < 15; i++) s[i] = 0.0;
        // masses from the shared {mass, 1/mass} table of the owned components and their sum taken once per kernel: the
        // global loads of the mass array (6 + 3 NO per step, dependent) were long-scoreboard stalls of the biased step
        const double msum = mass_sum;
        double vk[NO];   // velocity of the owned components, p / m as the reference divides it (kept for the last loop)
#pragma unroll
        for (int k = 0; k < NO; k++) {
            vk[k] = 0.0;
            if (oc[k] >= 0) {
                const int c = oc[k], j = c / 3, d = c - 3 * j;
                const double w = mt[k].x;
                const double v = CRCL_DIV(P(c), w);
                vk[k] = v;
                const double qx = Q(3 * j), qy = Q(3 * j + 1), qz = Q(3 * j + 2);
                // s[d] += v w without a run-time index (d is one when a lane owns a single component, PesSpread: s would
                // move to local memory); adding 0.0 changes no bits
                const double vw = v * w;
                s[0] += (d == 0) ? vw : 0.0;
                s[1] += (d == 1) ? vw : 0.0;
                s[2] += (d == 2) ? vw : 0.0;
                // (q x e_d) v w
                if (d == 0) {
                    s[7] += qz * v * w;
                    s[8] -= qy * v * w;
                    s[3] += qx * w;
                    s[4] += qy * w;
                    s[5] += qz * w;
                } else if (d == 1) {
                    s[6] -= qz * v * w;
                    s[8] += qx * v * w;
                } else {
                    s[6] += qy * v * w;
                    s[7] -= qx * v * w;
                }
            }
        }
        {
            double s9[9];
#pragma unroll
            for (int i = 0; i < 9; i++) s9[i] = s[i];
            G.sum_n(s9);
#pragma unroll
            for (int i = 0; i < 9; i++) s[i] = s9[i];
        }
        const double totmass = (msum * NB) * NB;
        double vtot[3], ctr[3], mang[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            vtot[d] = CRCL_DIV(s[d], totmass);
            ctr[d] = CRCL_DIV(s[3 + d], totmass);
        }
        mang[0] = s[6] - (ctr[1] * vtot[2] - ctr[2] * vtot[1]) * totmass;
        mang[1] = s[7] - (ctr[2] * vtot[0] - ctr[0] * vtot[2]) * totmass;
        mang[2] = s[8] - (ctr[0] * vtot[1] - ctr[1] * vtot[0]) * totmass;
        double xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0;
#pragma unroll
        for (int k = 0; k < NO; k++)
            if (oc[k] >= 0 && (oc[k] % 3) == 0) {
                const int j = oc[k] / 3;
                const double w = mt[k].x;
                const double xd = Q(3 * j) - ctr[0], yd = Q(3 * j + 1) - ctr[1], zd = Q(3 * j + 2) - ctr[2];
                xx += xd * xd * w;
                xy += xd * yd * w;
                xz += xd * zd * w;
                yy += yd * yd * w;
                yz += yd * zd * w;
                zz += zd * zd * w;
            }
        {
            double s6[6] = {xx, xy, xz, yy, yz, zz};
            G.sum_n(s6);
            xx = s6[0];
            xy = s6[1];
            xz = s6[2];
            yy = s6[3];
            yz = s6[4];
            zz = s6[5];
        }
        double t[3][3] = {{yy + zz, -xy, -xz}, {-xy, xx + zz, -yz}, {-xz, -yz, xx + yy}};
        if (NAT <= 2) {
            t[0][0] += 0.000001;
            t[1][1] += 0.000001;
            t[2][2] += 0.000001;
        }
        // every thread inverts its own copy of the tensor: the sums above are bit-identical on all threads of the
        // trajectory and the inverse lives in registers, so this costs no barrier and nobody waits for a publishing thread
#ifdef CRCL_TRANSROT_GAUSS_JORDAN   // invert.f90's pivoted elimination operation by operation (A/B builds; the split path uses it)
        if (invert3(t)) return 1;
#else
        if (invert3_sym(t)) return 1;
#endif
        double vang[3];
#pragma unroll
        for (int i = 0; i < 3; i++) vang[i] = t[i][0] * mang[0] + t[i][1] * mang[1] + t[i][2] * mang[2];
#pragma unroll
        for (int k = 0; k < NO; k++)
            if (oc[k] >= 0) {
                const int c = oc[k], j = c / 3, d = c - 3 * j;
                const double w = mt[k].x;
                const double xd = Q(3 * j) - ctr[0], yd = Q(3 * j + 1) - ctr[1], zd = Q(3 * j + 2) - ctr[2];
                double v = vk[k] - ((d == 0) ? vtot[0] : ((d == 1) ? vtot[1] : vtot[2]));
                if (d == 0)
                    v = v - vang[1] * zd + vang[2] * yd;
                else if (d == 1)
                    v = v - vang[2] * xd + vang[0] * zd;
                else
                    v = v - vang[0] * yd + vang[1] * xd;
                P(c) = mov(k) ? v * w : 0.0;
            }
        return 0;
    }

    // NaN / Inf scan of the positions (verlet.f90:1256-1275)
    __device__ __forceinline__ void nan_scan()
    {
        int nan = 0;
#pragma unroll
        for (int k = 0; k < NO; k++)
            if (oc[k] >= 0) {
                const double v = Q(oc[k]);
                nan |= (v != v) || (v > 1.79769313486231570815e308);
            }
        if (G.any(nan)) status |= CRCL_TRAJ_NAN;
    }

    // one verlet step (SURVEY.md 3.5 numbering)
    // check_nan: run the NaN / Inf scan of verlet.f90:1256-1275 in this step.  A NaN coordinate never
    // recovers, so the callers scan every 16th and the last step of a launch: same status, one
    // CTA-wide vote less per step.
    // CM: the constrain mode as a compile-time constant (the recrossing children always run mode 2: their kernel then
    // carries none of the thermostat / bias / SHAKE / rotation-removal code), or -99 to read it from the arguments.
    template <int CM = -99>
    __device__ __forceinline__ void step(int istep, bool check_nan = true)
    {
        const int c = (CM == -99) ? A.constrain : CM, th = A.thermostat;
        if (c != 2 && th == 2) nhc();                          // 1
        half_kick();                                           // 2,3
        free_rp();                                             // 4
        centroid();                                            // 6
        mask_p();                                              // 7
        int bad = 0;
        if (c == 1) bad = constrain_q();                       // 9
        epot = forces();                                       // 10
        if (bad) {
            epot += 100000.0;
            status |= CRCL_TRAJ_SHAKE_FAIL;
        }
        if (c == 0 || c == 3)                                  // 12
            umbrella(0);
        else if (c == 1)
            umbrella(1);
        else if (c == 2)
            umbrella(2);
        half_kick();                                           // 13
        if (c == 1) constrain_p();                             // 14
        if (c != 2 && th == 2) nhc();                          // 15
        if (c != 2 && th == 1 && A.andersen_step > 0 && (istep % A.andersen_step) == 0)
            andersen();                                        // 16
        if (check_nan) nan_scan();                             // 18
        if (c <= 0)                                            // 19
            if (transrot()) status |= CRCL_TRAJ_SINGULAR;
        // the drivers' rpmd_check right after verlet (rpmd_check.f90:88-116; calc_rate.f90:945,1575,1633,
        // recross.f90:275 -- the latter passes xi_ideal twice, so the xi test only exists in the umbrella modes)
        if (A.chk_on && c >= 0 && c != 2) {
            if (epot != epot || epot > 1.79769313486231570815e308) status |= CRCL_TRAJ_NAN;
            if (epot > A.chk_emax) status |= CRCL_TRAJ_ENERGY;
            if (c != 1 && fabs(xi_real - xi_ideal) > A.chk_xi_tol) status |= CRCL_TRAJ_XI_RANGE;
        }
    }

    // state <-> HBM in the reference layout [traj][bead][atom][xyz]
    __device__ __forceinline__ void load_qp(const double* qsrc, size_t qoff, const double* psrc, size_t poff)
    {
#pragma unroll
        for (int k = 0; k < NO; k++)
            if (oc[k] >= 0)
                pq[ob[k] + G.bead] = make_double2(psrc ? psrc[poff + oc[k]] : 0.0, qsrc[qoff + oc[k]]);
    }
};

// minimum resident CTAs per SM requested from ptxas (register cap = 64K / (MINB * threads)).
// The lane-split CH4+H kernels need 7 CTAs of 64 threads per SM to hold the BASELINE batch
// (1024 trajectories on 148 SMs) in a single wave; see profiles/ for the measured trade-off.
#ifndef CRCL_MINB_L4
#define CRCL_MINB_L4 7
#endif
// one-lane surfaces in single-warp CTAs (H3 / OH3 with <= 32 beads): 12 CTAs per SM (168 registers) instead
// of the 236 the compiler takes unbounded: +19 % on a 65536-replica H + H2 batch, -3 % at 1024 replicas
// (profiles/r1k_*; the kernel is bound by instruction-cache misses at 2 warps per scheduler)
#ifndef CRCL_MINB_L1
#define CRCL_MINB_L1 12
#endif
// one-lane surfaces on multi-warp trajectories (OH + H2 x 64 beads: 12 components per thread, four accumulators
// each in the free ring-polymer step): resident 128-thread CTAs requested.  4 (128 registers instead of the 220 the
// compiler takes unbounded, 184 bytes of spills) puts the 512 CTAs of 1024 child trajectories in one wave instead of
// 1.7: 12.9 -> 11.7 ms per 500 steps (3: 13.0 ms)
#ifndef CRCL_MINB_C1
#define CRCL_MINB_C1 4
#endif
template <class PES, int NB>
struct LaunchCfg {
    static constexpr int TPB = Group<NB, PES::LANES>::TPB;
    // one-lane surfaces (H3, OH3): CRCL_MINB_L1 resident CTAs requested when a CTA is a single warp
    static constexpr bool WARP = Group<NB, PES::LANES>::WARP;
    // lane-split surfaces on multi-warp trajectories: the CRCL_MINB_L4 x 64 threads of the single-wave tuning, whatever
    // the packing (7 CTAs of 64 threads or 4 of 128: a 128-register cap either way)
    static constexpr int MINB = WARP ? ((PES::LANES > 1) ? 2 : (CRCL_MINB_L1 * 32 / CRCL_WTPB > 0 ? CRCL_MINB_L1 * 32 / CRCL_WTPB : 1))
                                     : ((PES::LANES > 1 && (TPB <= 128 || Group<NB, PES::LANES>::GPB > 1)) ? (CRCL_MINB_L4 * 64 / TPB > 0 ? (CRCL_MINB_L4 * 64 + TPB - 1) / TPB : 1)
                                                                       : ((PES::LANES == 1 && TPB <= 128) ? CRCL_MINB_C1 : 1));
};

// Free ring-polymer kernels into shared memory.  NB <= 32: the dense tables
//   H_x[b][a] = f_x((a-b) mod N)                          (CRCL_TRANSFORM_EXACT)
//   H_x[b][a] = (f_x((a-b) mod N) + f_x((a+b) mod N)) / 2  (reference: Circ(f).(I+J)/2)
// for x = c, a, b, so that the symmetrisation costs nothing per step; otherwise f_x[N] as they come.
template <class PES, int NB>
__device__ __forceinline__ void load_fker(const TrajArgs& A, double* smem)
{
    constexpr int NAT = PES::NATOMS, LANES = PES::LANES;
    static_assert(2 * LANES * PES::NOWN <= SmemLayout<NAT, NB, LANES>::MT_MAX, "enlarge MT_MAX");
    for (int i = threadIdx.x; i < LANES * PES::NOWN; i += blockDim.x) {
        const int oc = PES::owned(i / PES::NOWN, i % PES::NOWN);
        const double m = (oc >= 0) ? A.mass[oc / 3] : 1.0;
        reinterpret_cast<double2*>(smem + SmemLayout<NAT, NB, LANES>::MT_OFF)[i] = make_double2(m, 1.0 / m);
    }
    if (SmemLayout<NAT, NB, LANES>::DMMA) {
        for (int i = threadIdx.x; i < 3 * PES::NATOMS; i += blockDim.x) {
            const double m = A.mass[i / 3];
            reinterpret_cast<double2*>(smem + SmemLayout<NAT, NB, LANES>::MC_OFF)[i] = make_double2(m, 1.0 / m);
        }
    }
    if (SmemLayout<NAT, NB, LANES>::HTAB) {
        for (int i = threadIdx.x; i < 3 * NB * NB; i += blockDim.x) {
            const int x = i / (NB * NB), r = i - x * NB * NB, b = r / NB, a = r - b * NB;
            const double f1 = A.fker[x * NB + ((a - b) & (NB - 1))];
            const double v = A.symmetrize ? 0.5 * (f1 + A.fker[x * NB + ((a + b) & (NB - 1))]) : f1;
            if (SmemLayout<NAT, NB, LANES>::DMMA) {
                // A-fragment order of mma.m8n8k4 (row a = lane / 4, column b = lane % 4 of an 8 x 4 tile):
                // [x][row tile a / 8][k step b / 4][lane]
                constexpr int MT = NB / 8, KS = NB / 4;
                smem[((x * MT + a / 8) * KS + b / 4) * 32 + (a % 8) * 4 + (b % 4)] = v;
            } else {
                smem[i] = v;
            }
        }
    } else {
        for (int i = threadIdx.x; i < 3 * NB; i += blockDim.x) smem[i] = A.fker[i];
    }
    __syncthreads();
}

// ---- generic batched verlet: state in HBM in, nsteps steps, state out --------------------
template <class PES, int NB>
__global__ void __launch_bounds__(LaunchCfg<PES, NB>::TPB, LaunchCfg<PES, NB>::MINB)
verlet_kernel(const __grid_constant__ TrajArgs A)
{
    extern __shared__ __align__(16) double smem[];
    constexpr int NC = 3 * PES::NATOMS, NO = PES::NOWN;
    using Grp = Group<NB, PES::LANES>;
    load_fker<PES, NB>(A, smem);
    Grp G(smem + SmemLayout<PES::NATOMS, NB, PES::LANES>::FKER);
    const int traj = blockIdx.x * Grp::GPB + G.gib;
    if (traj >= A.ntraj) return;
    Traj<PES, NB> T(A, G, smem);
    const size_t off = ((size_t)traj * NB + G.bead) * NC;
    T.load_qp(A.q, off, A.p, off);
#pragma unroll
    for (int k = 0; k < NO; k++)
        if (T.oc[k] >= 0) T.g[k] = A.g[off + T.oc[k]];
    T.xi_ideal = A.xi_ideal ? A.xi_ideal[traj] : A.xi_ideal_s;
    T.k_force = A.k_force ? A.k_force[traj] : A.k_force_s;
    T.tid = A.traj_id ? A.traj_id[traj] : A.traj_id0 + (uint32_t)traj;
    T.event = A.event ? A.event[traj] : 0u;
    if (A.dxi)
        for (int c = G.tig; c < NC; c += Grp::T) T.dxi[c] = A.dxi[(size_t)traj * NC + c];
    if (A.nhc) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            T.vnh[i] = A.nhc[(size_t)traj * 8 + i];
            T.qnh[i] = A.nhc[(size_t)traj * 8 + 4 + i];
        }
    }
    T.status = A.status ? A.status[traj] : 0;
    G.sync();
    double sx = 0.0, sx2 = 0.0;
    for (int s = 1; s <= A.nsteps; s++) {
        Grp::align_warps();
        // a failed trajectory is frozen (the reference aborts or restarts it)
        if (!(T.status & CRCL_TRAJ_FATAL)) {
            // epot leaves the kernel from the last step only (and feeds rpmd_check when that is on): the other steps skip
            // the trajectory-wide sum of the bead energies.  A failed SHAKE still raises the status on its own step.
            T.want_epot = (s == A.nsteps) || A.chk_on;
            T.step(A.istep0 + s, (s & 15) == 0 || s == A.nsteps);
            sx += T.xi_real;
            sx2 += T.xi_real * T.xi_real;
        }
    }
    // child steps only evaluate the value of xi; leave dxi as verlet.f90:1049-1050 would
    if (A.constrain == 2 && A.nsteps > 0) T.umbrella(1);
    G.sync();
#pragma unroll
    for (int k = 0; k < NO; k++)
        if (T.oc[k] >= 0) {
            const int c = T.oc[k];
            A.q[off + c] = T.Q(c);
            A.p[off + c] = T.P(c);
            A.g[off + c] = T.g[k];
        }
    if (A.dxi)
        for (int c = G.tig; c < NC; c += Grp::T) A.dxi[(size_t)traj * NC + c] = T.dxi[c];
    if (G.tig == 0) {
        if (A.epot) A.epot[traj] = T.epot;
        if (A.xi_real) A.xi_real[traj] = T.xi_real;
        if (A.status) A.status[traj] = T.status;
        if (A.event) A.event[traj] = T.event;
        if (A.xi_sum) A.xi_sum[traj] += sx;
        if (A.xi_sum2) A.xi_sum2[traj] += sx2;
        if (A.nhc) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                A.nhc[(size_t)traj * 8 + i] = T.vnh[i];
                A.nhc[(size_t)traj * 8 + 4 + i] = T.qnh[i];
            }
        }
    }
}

// ---- mdinit (mdinit.f90:40-172) -------------------------------------------------------------
// bias_mode: 0 no umbrella call, 1 xi only (umbrella mode 1), 2 bias applied (umbrella mode 0).
template <class PES, int NB>
__global__ void __launch_bounds__(LaunchCfg<PES, NB>::TPB, LaunchCfg<PES, NB>::MINB)
mdinit_kernel(const __grid_constant__ TrajArgs A, const int bias_mode, const double nose_q)
{
    extern __shared__ __align__(16) double smem[];
    constexpr int NC = 3 * PES::NATOMS, NO = PES::NOWN;
    using Grp = Group<NB, PES::LANES>;
    load_fker<PES, NB>(A, smem);
    Grp G(smem + SmemLayout<PES::NATOMS, NB, PES::LANES>::FKER);
    const int traj = blockIdx.x * Grp::GPB + G.gib;
    if (traj >= A.ntraj) return;
    Traj<PES, NB> T(A, G, smem);
    const size_t off = ((size_t)traj * NB + G.bead) * NC;
    T.load_qp(A.q, off, A.p, off);
    T.xi_ideal = A.xi_ideal ? A.xi_ideal[traj] : A.xi_ideal_s;
    T.k_force = A.k_force ? A.k_force[traj] : A.k_force_s;
    T.tid = A.traj_id ? A.traj_id[traj] : A.traj_id0 + (uint32_t)traj;
    T.event = A.event ? A.event[traj] : 0u;
    T.epot = T.forces();
    T.centroid();
    if (bias_mode == 1)
        T.umbrella(1);
    else if (bias_mode == 2)
        T.umbrella(0);
    if (A.thermostat == 0 || A.thermostat == 1) T.andersen();
    if (A.thermostat == 2) {
        T.andersen();
        T.nhc_init(nose_q);
    }
#pragma unroll
    for (int k = 0; k < NO; k++)
        if (T.oc[k] >= 0) {
            A.g[off + T.oc[k]] = T.g[k];
            A.p[off + T.oc[k]] = T.P(T.oc[k]);
        }
    if (A.dxi && bias_mode != 0)
        for (int c = G.tig; c < NC; c += Grp::T) A.dxi[(size_t)traj * NC + c] = T.dxi[c];
    if (G.tig == 0) {
        if (A.event) A.event[traj] = T.event;
        if (A.nhc && A.thermostat == 2) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                A.nhc[(size_t)traj * 8 + i] = T.vnh[i];
                A.nhc[(size_t)traj * 8 + 4 + i] = T.qnh[i];
            }
        }
    }
}

// ---- recrossing child trajectories (recross.f90:515-628 / recross_serial.f90:172-229) ------
// trajectory t = 2*g + k is child k (0: +p, 1: -p) of pair pair0+g.  Both children of a pair
// draw the same momenta (RNG stream keyed by the pair index).  Writes weight = v_s/f_s,
// denom_part and, per step, theta = [xi_real > 0]; kappa sums are formed by reduce_kappa.
// CRCL_RECROSS_MAXNREG (tuning builds): an explicit register cap for the lane-split child kernel instead of the
// minimum-blocks form, whose cap ptxas rounds down to a power-of-two-ish 128 (64-thread CTAs x 7 per SM would allow 144)
#ifdef CRCL_RECROSS_MAXNREG
#define CRCL_RECROSS_BOUNDS(PES, NB) __maxnreg__(CRCL_RECROSS_MAXNREG)
#else
#define CRCL_RECROSS_BOUNDS(PES, NB) __launch_bounds__(LaunchCfg<PES, NB>::TPB, LaunchCfg<PES, NB>::MINB)
#endif
template <class PES, int NB>
__global__ void CRCL_RECROSS_BOUNDS(PES, NB)
recross_kernel(const __grid_constant__ TrajArgs A)
{
    extern __shared__ __align__(16) double smem[];
    constexpr int NC = 3 * PES::NATOMS, NO = PES::NOWN;
    using Grp = Group<NB, PES::LANES>;
    load_fker<PES, NB>(A, smem);
    Grp G(smem + SmemLayout<PES::NATOMS, NB, PES::LANES>::FKER);
    const int traj = blockIdx.x * Grp::GPB + G.gib;
    if (traj >= A.ntraj) return;
    Traj<PES, NB> T(A, G, smem);
    T.want_epot = false;   // the child body never reads epot (recross_serial.f90:192-199)
    const int pair = A.pair0 + (traj >> 1);
    const double sign = (traj & 1) ? -1.0 : 1.0;
    const size_t poff = ((size_t)(pair % A.nparent) * NB + G.bead) * NC;
    T.load_qp(A.q_parents, poff, nullptr, 0);
    T.xi_ideal = A.xi_ideal_s;
    T.k_force = 0.0;
    T.tid = (uint32_t)pair;
    T.event = 0u;
    T.andersen();
#pragma unroll
    for (int k = 0; k < NO; k++)
        if (T.oc[k] >= 0) T.P(T.oc[k]) = sign * T.P(T.oc[k]);
    T.centroid();
    T.umbrella(1);  // calc_xi(mode 2) on the centroid -> dxi (recross_serial.f90:186-187)
    T.epot = T.forces();
    double vs = 0.0, fs = 0.0;
#pragma unroll
    for (int k = 0; k < NO; k++)
        if (T.own(k)) vs += T.dxi[T.oc[k]] * T.Pk(k) / T.mt[k].x;
#pragma unroll
    for (int c = 0; c < NC; c++) fs += T.dxi[c] * T.dxi[c] / A.mass[c / 3];
    vs = G.sum(vs) / NB;
    fs = sqrt(fs / (2.0 * PI_UMBR * A.beta));
    const double w = vs / fs;
    if (G.tig == 0) {
        A.weight[traj] = w;
        A.denom_part[traj] = (vs > 0) ? w : 0.0;
    }
    // Tried and rejected on measurement (profiles/r2j_*): a child step arranged around two barriers instead of six (CTA
    // re-alignment doubling as the staging barrier, centroid of q' from the staged sums through the column means of the
    // tables, surface behind a warp-level fence): 13.8 against 12.6 ms per 1000 steps -- warps that meet at fewer barriers
    // drift apart in the ~50 KB step body and stop sharing fetched instruction lines (no_instruction 0.59 -> 1.00 cycles
    // per issue), and the second copy of the transform raised the spills.
    // step() in mode 2 (A.constrain = 2, api.cu) written out, so that the second half kick of a step and the first of the
    // next are one pass: kick, [free ring polymer, centroid, mask, forces, xi, kick + kick] ..., kick
    if (A.nsteps > 0) T.half_kick();                           // 2,3 of the first step
    // loop state in registers (steps left, theta cursor): the kernel arguments live in the constant bank, and ncu showed the
    // per-step reloads of nsteps / ntraj / theta behind `l < A.nsteps` among the top long-scoreboard lines
    unsigned char* th = A.theta + traj;
    const int ntraj = A.ntraj;
    for (int l = 1, rem = A.nsteps; rem > 0; l++, rem--, th += ntraj) {
        Grp::align_warps();
        if (!(T.status & CRCL_TRAJ_NAN)) {
            // (tried and rejected on measurement, profiles/r2o_*: the centroid of q' from the staged sums -- the centroid
            // mode of a free ring polymer moves as a free particle -- on the warp that evaluates xi, without centroid()'s
            // two barriers: 11.80 against 11.76 ms per 1000 steps)
            T.free_rp();                                       // 4
            T.centroid();                                      // 6
            T.mask_p();                                        // 7
            T.epot = T.forces();                               // 10
            T.umbrella(2);                                     // 12
            if (rem > 1)
                T.double_half_kick();                          // 13 and 2,3 of the next step
            else
                T.half_kick();                                 // 13
            if ((l & 15) == 0 || rem == 1) T.nan_scan();       // 18
        }
        if (T.xi_writer()) *th = (T.xi_real > 0) ? 1 : 0;
    }
    if (G.tig == 0 && A.status) A.status[traj] = T.status;
}

}  // namespace crcl
