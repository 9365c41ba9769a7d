// transform_bench.cuh -- the free ring-polymer step (verlet.f90:377-463) as a stand-alone kernel in two forms, for the
// question BASELINE.json's north_star asks: do the FP64 tensor cores (DMMA, mma.sync.m8n8k4.f64 -- tcgen05 has no FP64
// kind) beat the FMA path on the bead transform?  Both forms keep a CTA's trajectories resident in shared memory, as the
// fused trajectory kernels do, and apply
//     P' = Fc P + m Fa Q,   Q' = Fb P / m + Fc Q        (Fc, Fa, Fb: NB x NB circulant kernels, P, Q: NB x NC per trajectory)
// `reps` times, so that the timing is the transform from shared memory and not the HBM load around it.
//   MODE 0 (DFMA): the loop of Traj::free_rp -- a thread owns a bead and up to NO components, walks the NB source beads,
//                  one 16-byte shared load of {p,q} and three 8-byte loads of the coefficients per 4 NO fused multiply-adds;
//   MODE 1 (DMMA): output tiles of 8 beads x 8 components per warp; A fragments = the coefficient matrices (8 x 4), B
//                  fragments = {p,q} of 4 source beads x 8 components (one 16-byte load yields the P and the Q fragment),
//                  four accumulators (Fc P, Fa Q, Fb P, Fc Q); two column tiles share the A fragments.
// The columns of the TPC trajectories of a CTA are laid side by side (2 x 12 = 24 = 3 tiles for OH + H2, no padding;
// 4 x 18 = 72 = 9 tiles for CH4 + H).  crcl_bench_transform runs both on the same input and compares the results.
#pragma once
#include <cuda_runtime.h>

namespace crcl {

template <int NB, int NC, int TPC, int MODE>
struct TransformBench {
    static constexpr int LN = (NB >= 64) ? 1 : 4;                  // lanes per bead of the DFMA form (as PesOH3 / PesCBE4)
    static constexpr int NO = (NC + LN - 1) / LN;                  // components per thread
    static constexpr int TPT = NB * LN;                            // threads per trajectory
    // DFMA: one thread per (trajectory, bead, lane); DMMA: warps share the (row tile, column-tile group) work items evenly
    // -- 8 x 3 items of two column tiles on 8 warps for 64 beads, 2 x 9 items of one column tile on 6 warps for 16 beads
    static constexpr int NG = (NB >= 64) ? 2 : 1;                  // column tiles per work item (they share the A fragments)
    static constexpr int THREADS = (MODE == 1) ? ((NB >= 64) ? 256 : 192) : TPC * TPT;
    static constexpr int NBP = (MODE == 1) ? NB + 4 : NB + 1;      // bead stride of {p,q}[c][b]: conflict-free for each form
    static constexpr int LDH = (MODE == 1) ? NB + 4 : NB;          // row stride of the coefficient matrices
    static constexpr int NCOL = TPC * NC;
    static constexpr int NT = (NCOL + 7) / 8;                      // column tiles
    static constexpr int HEAD = (3 * NB * LDH + 2 * NC + 1) & ~1;  // doubles before the {p,q} buffers (even: 16-byte aligned)
    static constexpr size_t smem_bytes() { return sizeof(double) * HEAD + 2 * sizeof(double2) * TPC * NC * NBP; }
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// f: [3][NB] circulant kernels (build_fker), mass: [NC], pq: [ntraj][NC][NB] {p,q}; ntraj a multiple of TPC
template <int NB, int NC, int TPC, int MODE>
__global__ void __launch_bounds__(TransformBench<NB, NC, TPC, MODE>::THREADS)
    transform_bench_kernel(const double* __restrict__ f, const double* __restrict__ mass, double2* __restrict__ pq, int reps)
{
    using B = TransformBench<NB, NC, TPC, MODE>;
    extern __shared__ double tb_sh[];
    double* H = tb_sh;                                             // MODE 0: [k][b][a] (threads of consecutive beads a read
    double* ms = H + 3 * NB * B::LDH;                              //         consecutive words); MODE 1: [k][a][b + pad]
    double* ims = ms + NC;                                         // masses and their inverses
    double2* X0 = reinterpret_cast<double2*>(tb_sh + B::HEAD);
    double2* X1 = X0 + TPC * NC * B::NBP;
    const int tid = threadIdx.x;
    for (int i = tid; i < 3 * NB * NB; i += B::THREADS) {
        const int k = i / (NB * NB), r = (i / NB) % NB, c = i % NB;
        // H_k[a][b] = f_k[(a - b) mod NB]; MODE 0 stores the transpose ([b][a]), which for a circulant of a symmetric
        // kernel is the same matrix
        H[k * NB * B::LDH + r * B::LDH + c] = f[k * NB + ((r - c + NB) % NB)];
    }
    for (int i = tid; i < NC; i += B::THREADS) {
        ms[i] = mass[i];
        ims[i] = 1.0 / mass[i];
    }
    double2* g = pq + (size_t)blockIdx.x * TPC * NC * NB;
    for (int i = tid; i < TPC * NC * NB; i += B::THREADS) X0[(i / NB) * B::NBP + (i % NB)] = g[i];
    __syncthreads();
    double2 *in = X0, *out = X1;
    if (MODE == 0) {
        const int t = tid / B::TPT, tig = tid % B::TPT, a = tig / B::LN, l = tig % B::LN;
        const double2* xin0 = in + t * NC * B::NBP;
        for (int r = 0; r < reps; r++) {
            const double2* xin = (r & 1) ? X1 + t * NC * B::NBP : xin0;
            double2* xout = ((r & 1) ? X0 : X1) + t * NC * B::NBP;
            double cp[B::NO], aq[B::NO], bp[B::NO], cq[B::NO];
#pragma unroll
            for (int k = 0; k < B::NO; k++) cp[k] = aq[k] = bp[k] = cq[k] = 0.0;
            const double* hk = H + a;
#pragma unroll 2
            for (int b = 0; b < NB; b++) {
                const double fc = hk[b * NB], fa = hk[NB * NB + b * NB], fb = hk[2 * NB * NB + b * NB];
#pragma unroll
                for (int k = 0; k < B::NO; k++) {
                    const int c = (l + B::LN * k < NC) ? l + B::LN * k : 0;
                    const double2 u = xin[c * B::NBP + b];
                    cp[k] = fma(fc, u.x, cp[k]);
                    aq[k] = fma(fa, u.y, aq[k]);
                    bp[k] = fma(fb, u.x, bp[k]);
                    cq[k] = fma(fc, u.y, cq[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < B::NO; k++) {
                const int c = l + B::LN * k;
                if (c < NC) {
                    xout[c * B::NBP + a] = make_double2(fma(ms[c], aq[k], cp[k]), fma(ims[c], bp[k], cq[k]));
                }
            }
            __syncthreads();
        }
    } else {
        const int warp = tid >> 5, lane = tid & 31, nwarp = B::THREADS / 32;
        const int gr = lane >> 2, tg = lane & 3;                   // fragment coordinates of this lane
        constexpr int MT = NB / 8, NG = B::NG, NGR = (B::NT + NG - 1) / NG;   // row tiles, groups of column tiles
        for (int r = 0; r < reps; r++) {
            const double2* xin = (r & 1) ? X1 : X0;
            double2* xout = (r & 1) ? X0 : X1;
            for (int w = warp; w < MT * NGR; w += nwarp) {
                const int m0 = (w % MT) * 8, n0 = (w / MT) * (8 * NG);
                // B fragment sources: column n0 + 8 i + gr of tile i (clamped: a padded column computes on a copy of the
                // last one and is not stored)
                const double2* xb[NG];
#pragma unroll
                for (int i = 0; i < NG; i++) {
                    const int j = n0 + 8 * i + gr;
                    xb[i] = xin + ((j < B::NCOL) ? j : B::NCOL - 1) * B::NBP + tg;
                }
                const double* ha = H + (m0 + gr) * B::LDH + tg;
                double cp[NG][2], aq[NG][2], bp[NG][2], cq[NG][2];
#pragma unroll
                for (int i = 0; i < NG; i++) cp[i][0] = cp[i][1] = aq[i][0] = aq[i][1] = bp[i][0] = bp[i][1] = cq[i][0] = cq[i][1] = 0.0;
#pragma unroll 4
                for (int kb = 0; kb < NB; kb += 4) {
                    const double fc = ha[kb], fa = ha[NB * B::LDH + kb], fb = ha[2 * NB * B::LDH + kb];
#pragma unroll
                    for (int i = 0; i < NG; i++) {
                        const double2 u = xb[i][kb];
                        dmma884(cp[i][0], cp[i][1], fc, u.x);
                        dmma884(aq[i][0], aq[i][1], fa, u.y);
                        dmma884(bp[i][0], bp[i][1], fb, u.x);
                        dmma884(cq[i][0], cq[i][1], fc, u.y);
                    }
                }
                // accumulator layout: row m0 + gr, columns n0 + 8 i + 2 tg + {0, 1}
#pragma unroll
                for (int i = 0; i < NG; i++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int j = n0 + 8 * i + 2 * tg + e;
                        if (j < B::NCOL) {
                            const int c = j % NC;
                            xout[j * B::NBP + m0 + gr] = make_double2(fma(ms[c], aq[i][e], cp[i][e]), fma(ims[c], bp[i][e], cq[i][e]));
                        }
                    }
            }
            __syncthreads();
        }
    }
    const double2* res = (reps & 1) ? X1 : X0;
    for (int i = tid; i < TPC * NC * NB; i += B::THREADS) g[i] = res[(i / NB) * B::NBP + (i % NB)];
}

}  // namespace crcl
