// split_kernels.cuh -- HBM-resident ("split") RPMD step for systems that do not fit the fused
// in-register trajectory kernels: any number of atoms, any bead count, PES evaluated by a separate
// device kernel or by the host callback (custom_grad / external_grad seam).
//
// One verlet step (verlet.f90:65-1308) becomes
//   sp_kick_freerp   2,3,4,6,7   p -= dt/2 g ; mask ; free ring polymer ; centroid     HBM-bound
//   <PES>            10          egrad kernel on all images, or D2H -> host callback -> H2D
//   sp_kick          13          p -= dt/2 g ; mask ; NaN/Inf scan (18)                 HBM-bound
//   sp_andersen      16          Philox resampling (every andersen_step steps)
//   sp_xi_value      12          xi on the centroid (child trajectories, constrain = 2)
//   sp_transrot_*    19          two-pass reductions + apply (constrain <= 0)
// State stays in the reference layout [traj][bead][atom][xyz]; threads map to consecutive
// components (atom,xyz) so every global access of a warp is one contiguous segment.
//
// sp_kick_freerp is the kernel the bandwidth roofline is quoted on: algorithmic traffic per
// (trajectory, bead, atom) = read q,p,g + write q,p = 120 B.  Each thread owns one
// component of one trajectory, holds its NB beads of p and q in registers and applies
// Circ(f).(I+J)/2 (see traj_kernel.cuh) -- NB^2 * 4 FMA per component against 40*NB bytes.
#pragma once
#include "crcl_common.cuh"
#include "rng.cuh"

namespace crcl {

struct SplitArgs {
    int ntraj, natoms, nbeads, symmetrize;
    double dt, beta;
    const double* mass;     // [natoms] device
    const int* at_move;     // [natoms] device
    const double* fker;     // [3][nbeads] device
    double* q;
    double* p;
    const double* g;
    double* cen;            // [traj][natoms][3]
    int* status;            // [traj]
    uint64_t seed;
    const uint32_t* traj_id;
    uint32_t traj_id0;
    uint32_t* event;
    int periodic;           // pbc_mod: wrap of verlet.f90:591-641 after the position update
    double box[3];          // boxlen_x, boxlen_y, boxlen_z
};

// Periodic wrap of one (atom, xyz) column of beads, plain-box branch of verlet.f90:591-641: the beads are visited in
// order; while the visited bead lies below 0 (above L) ALL beads of the column are shifted by +L (-L); `tries` counts
// both directions together and more than 100 is `fatal` in the reference (returns true; the caller sets
// CRCL_TRAJ_PBC_FAIL).  get(b) / set(b, v) address bead b of the column.
template <class Get, class Set>
__device__ __forceinline__ bool sp_wrap_column(int nb, double boxlen, Get get, Set set)
{
    bool fail = false;
    for (int i = 0; i < nb; i++) {
        int tries = 0;
        while (get(i) < 0 && tries <= 100) {
            for (int b = 0; b < nb; b++) set(b, get(b) + boxlen);
            tries++;
        }
        while (get(i) > boxlen && tries <= 100) {
            for (int b = 0; b < nb; b++) set(b, get(b) - boxlen);
            tries++;
        }
        fail |= tries > 100;
    }
    return fail;
}

// ---- kick + free ring polymer + centroid, beads in registers (NB <= 16 compile-time) ----------
template <int NB>
__global__ void __launch_bounds__(128) sp_kick_freerp_reg(const SplitArgs A)
{
    const int nc = 3 * A.natoms;
    // flat index over (trajectory, component): CTAs stay full for small systems, and the threads of a
    // warp still read contiguous runs (a whole trajectory's components are adjacent)
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)A.ntraj * nc) return;
    const size_t t = gid / nc;
    const int c = (int)(gid - t * nc);
    const int atom = c / 3;
    const double m = A.mass[atom];
    const bool mv = A.at_move[atom] != 0;
    const double h = 0.5 * A.dt;
    const size_t base = t * NB * nc + c;
    double p[NB], q[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const size_t i = base + (size_t)b * nc;
        const double pv = A.p[i] - h * A.g[i];
        p[b] = mv ? pv : 0.0;
        q[b] = A.q[i];
    }
    if (NB == 1) {
        q[0] = q[0] + p[0] * A.dt / m;
    } else {
        if (A.symmetrize) {
#pragma unroll
            for (int b = 1; b <= (NB - 1) / 2; b++) {
                const double ps = 0.5 * (p[b] + p[NB - b]), qs = 0.5 * (q[b] + q[NB - b]);
                p[b] = p[NB - b] = ps;
                q[b] = q[NB - b] = qs;
            }
        }
        double fc[NB], fa[NB], fb[NB];
        const double im = 1.0 / m;
#pragma unroll
        for (int j = 0; j < NB; j++) {
            fc[j] = A.fker[j];
            fa[j] = m * A.fker[NB + j];
            fb[j] = im * A.fker[2 * NB + j];
        }
        double pn[NB], qn[NB];
#pragma unroll
        for (int a = 0; a < NB; a++) {
            double sp = 0.0, sq = 0.0;
#pragma unroll
            for (int b = 0; b < NB; b++) {
                const int j = (a - b + NB) % NB;
                sp = fma(fc[j], p[b], fma(fa[j], q[b], sp));
                sq = fma(fb[j], p[b], fma(fc[j], q[b], sq));
            }
            pn[a] = sp;
            qn[a] = sq;
        }
#pragma unroll
        for (int b = 0; b < NB; b++) {
            p[b] = pn[b];
            q[b] = qn[b];
        }
    }
    if (A.periodic) {                                                     // 5
        const double boxlen = A.box[c - 3 * atom];
        bool out = false;
#pragma unroll
        for (int b = 0; b < NB; b++) out |= (q[b] < 0) || (q[b] > boxlen);
        if (out) {   // rare: the column goes through local memory only on the step an atom crosses a face
            double qw[NB];
#pragma unroll
            for (int b = 0; b < NB; b++) qw[b] = q[b];
            if (sp_wrap_column(NB, boxlen, [&](int b) { return qw[b]; }, [&](int b, double v) { qw[b] = v; }))
                atomicOr(&A.status[t], CRCL_TRAJ_PBC_FAIL);
#pragma unroll
            for (int b = 0; b < NB; b++) q[b] = qw[b];
        }
    }
    double cs = 0.0;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const size_t i = base + (size_t)b * nc;
        A.p[i] = mv ? p[b] : 0.0;   // verlet.f90:655-661 masks again after the position update
        A.q[i] = q[b];
        cs += q[b];
    }
    A.cen[t * nc + c] = cs / NB;
}

// ---- same, any bead count: the thread's column of beads lives in shared memory -------------------
__global__ void sp_kick_freerp_smem(const SplitArgs A)
{
    extern __shared__ __align__(16) double sm[];
    const int NB = A.nbeads, BD = blockDim.x;
    double* fk = sm;                     // [3][NB]
    double* ps = sm + 3 * NB;            // [NB][BD]
    double* qs = ps + (size_t)NB * BD;   // [NB][BD]
    for (int i = threadIdx.x; i < 3 * NB; i += BD) fk[i] = A.fker[i];
    __syncthreads();
    const int nc = 3 * A.natoms;
    const size_t gid = (size_t)blockIdx.x * BD + threadIdx.x;   // flat (trajectory, component) index
    if (gid >= (size_t)A.ntraj * nc) return;
    const size_t t = gid / nc;
    const int c = (int)(gid - t * nc);
    const int atom = c / 3, x = threadIdx.x;
    const double m = A.mass[atom], im = 1.0 / m;
    const bool mv = A.at_move[atom] != 0;
    const double h = 0.5 * A.dt;
    const size_t base = t * NB * nc + c;
    for (int b = 0; b < NB; b++) {
        const size_t i = base + (size_t)b * nc;
        const double pv = A.p[i] - h * A.g[i];
        ps[b * BD + x] = mv ? pv : 0.0;
        qs[b * BD + x] = A.q[i];
    }
    double cs = 0.0;
    if (NB == 1) {
        const double qv = qs[x] + ps[x] * A.dt / m;
        A.q[base] = qv;
        A.p[base] = mv ? ps[x] : 0.0;
        cs = qv;
    } else {
        if (A.symmetrize)
            for (int b = 1; b <= (NB - 1) / 2; b++) {
                const double pv = 0.5 * (ps[b * BD + x] + ps[(NB - b) * BD + x]);
                const double qv = 0.5 * (qs[b * BD + x] + qs[(NB - b) * BD + x]);
                ps[b * BD + x] = ps[(NB - b) * BD + x] = pv;
                qs[b * BD + x] = qs[(NB - b) * BD + x] = qv;
            }
        for (int a = 0; a < NB; a++) {
            double cp = 0.0, aq = 0.0, bp = 0.0, cq = 0.0;
            int j = a;
            for (int b = 0; b < NB; b++) {
                const double pv = ps[b * BD + x], qv = qs[b * BD + x];
                const double fc = fk[j], fa = fk[NB + j], fb = fk[2 * NB + j];
                cp = fma(fc, pv, cp);
                aq = fma(fa, qv, aq);
                bp = fma(fb, pv, bp);
                cq = fma(fc, qv, cq);
                j = (j == 0) ? NB - 1 : j - 1;
            }
            const size_t i = base + (size_t)a * nc;
            const double pn = fma(m, aq, cp), qn = fma(im, bp, cq);
            A.p[i] = mv ? pn : 0.0;
            A.q[i] = qn;
            cs += qn;
        }
    }
    if (A.periodic) {                                                     // 5: on the thread's own column in HBM
        const double boxlen = A.box[c - 3 * atom];
        bool out = false;
        for (int b = 0; b < NB; b++) {
            const double qv = A.q[base + (size_t)b * nc];
            out |= (qv < 0) || (qv > boxlen);
        }
        if (out) {
            if (sp_wrap_column(NB, boxlen, [&](int b) { return A.q[base + (size_t)b * nc]; },
                               [&](int b, double v) { A.q[base + (size_t)b * nc] = v; }))
                atomicOr(&A.status[t], CRCL_TRAJ_PBC_FAIL);
            cs = 0.0;
            for (int b = 0; b < NB; b++) cs += A.q[base + (size_t)b * nc];
        }
    }
    A.cen[t * nc + c] = cs / NB;
}

// ---- second half kick + mask + NaN/Inf scan (verlet.f90:1060-1073, 1256-1275) --------------------
__global__ void sp_kick(const SplitArgs A)
{
    const int nc = 3 * A.natoms;
    const size_t per = (size_t)A.nbeads * nc;
    const int tiles = gridDim.x / A.ntraj;
    const int t = blockIdx.x / tiles;
    const size_t i = (size_t)(blockIdx.x - t * tiles) * blockDim.x + threadIdx.x;
    if (i >= per) return;
    const int atom = (int)(i % nc) / 3;
    const size_t k = (size_t)t * per + i;
    const double pv = A.p[k] - 0.5 * A.dt * A.g[k];
    A.p[k] = A.at_move[atom] ? pv : 0.0;
    const double qv = A.q[k];
    if (qv != qv || qv > 1.79769313486231570815e308) atomicOr(&A.status[t], CRCL_TRAJ_NAN);
}

// ---- Andersen resampling (andersen.f90:36-74) with the stream of rng.cuh ------------------------
__global__ void sp_andersen(const SplitArgs A)
{
    const int nc = 3 * A.natoms;
    const size_t per = (size_t)A.nbeads * nc;
    const int tiles = gridDim.x / A.ntraj;
    const int t = blockIdx.x / tiles;
    const size_t i = (size_t)(blockIdx.x - t * tiles) * blockDim.x + threadIdx.x;
    if (i >= per) return;
    const int bead = (int)(i / nc), mcomp = (int)(i % nc);
    const uint32_t tid = A.traj_id ? A.traj_id[t] : A.traj_id0 + (uint32_t)t;
    double z0, z1;
    normal_pair(A.seed, tid, A.event[t], (uint32_t)bead, (uint32_t)(mcomp >> 1), z0, z1);
    const double beta_n = A.beta / A.nbeads;
    A.p[(size_t)t * per + i] = ((mcomp & 1) ? z1 : z0) * sqrt(A.mass[mcomp / 3] / beta_n);
}
__global__ void sp_bump_event(uint32_t* event, int ntraj)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ntraj) event[t]++;
}

// ---- per-trajectory sum of the bead energies (verlet.f90:772-777) ---------------------------------
__global__ void sp_epot(const double* V, int nbeads, int ntraj, double* epot)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntraj) return;
    double s = 0.0;
    for (int b = 0; b < nbeads; b++) s += V[(size_t)t * nbeads + b];
    epot[t] = s;
}

// ---- generic-size reaction coordinate value on the centroid (calc_xi.f90:108-212) ----------------
struct MechDev {
    int form_num, break_num, sum_reacs;
    int bf[8][2], bb[8][2];
    double fref[8], bref[8];
    const int* frag;       // [natoms] device: fragment of each atom or -1
    const double* wfrag;   // [natoms] device: mass[a]/mass_reac[frag[a]]
    double R_inf;
    // umbr_type family as in xi.cuh: 0 BIMOLEC family, 1 unimolecular, 2 ATOM_SHIFT
    int type;
    double freac[8], breac[8];
    int shift_atom, shift_c1, shift_c2;
    double shift_lo, shift_hi, shift2_lo, shift2_hi;
};
__device__ __forceinline__ void sp_shift_s(const MechDev& M, const double* x, double& s0, double& s1)
{
    const double a = x[3 * M.shift_atom + M.shift_c1];
    if (M.shift_c2 < 0) {
        s1 = a - M.shift_hi;
        s0 = a - M.shift_lo;
    } else {
        const double b = x[3 * M.shift_atom + M.shift_c2];
        s1 = ((a - M.shift_hi) + (b - M.shift2_hi)) / 2.0;
        s0 = ((a - M.shift_lo) + (b - M.shift2_lo)) / 2.0;
    }
}
__global__ void sp_xi_value(MechDev M, int natoms, int ntraj, const double* cen, const double* xi_ideal,
                            double xi_ideal_s, int mode, double* xi_out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntraj) return;
    const double* x = cen + (size_t)t * 3 * natoms;
    double s1 = 0.0, s0 = 0.0;
    if (M.type == 2) {
        sp_shift_s(M, x, s0, s1);
    } else {
        double s0u = 0.0;
        for (int i = 0; i < M.break_num; i++) {
            const int a1 = M.bb[i][0], a2 = M.bb[i][1];
            const double dx = x[3 * a1] - x[3 * a2], dy = x[3 * a1 + 1] - x[3 * a2 + 1], dz = x[3 * a1 + 2] - x[3 * a2 + 2];
            const double r = sqrt(dx * dx + dy * dy + dz * dz);
            s1 += (r - M.bref[i]) / (double)M.break_num;
            s0u += (r - M.breac[i]) / (double)M.break_num;
        }
        for (int i = 0; i < M.form_num; i++) {
            const int a1 = M.bf[i][0], a2 = M.bf[i][1];
            const double dx = x[3 * a1] - x[3 * a2], dy = x[3 * a1 + 1] - x[3 * a2 + 1], dz = x[3 * a1 + 2] - x[3 * a2 + 2];
            const double r = sqrt(dx * dx + dy * dy + dz * dz);
            s1 -= (r - M.fref[i]) / (double)M.form_num;
            s0u -= (r - M.freac[i]) / (double)M.form_num;
        }
        if (M.type == 1) {
            s0 = s0u;
        } else {
            double com[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
            for (int a = 0; a < natoms; a++) {
                const int k = M.frag[a];
                if (k >= 0)
                    for (int d = 0; d < 3; d++) com[k][d] += M.wfrag[a] * x[3 * a + d];
            }
            for (int i = 0; i < M.sum_reacs; i++)
                for (int j = i + 1; j < M.sum_reacs; j++) {
                    const double dx = com[j][0] - com[i][0], dy = com[j][1] - com[i][1], dz = com[j][2] - com[i][2];
                    s0 += M.R_inf - sqrt(dx * dx + dy * dy + dz * dz);
                }
            s0 = s0 / (double)((M.sum_reacs * M.sum_reacs - M.sum_reacs) / 2);
        }
    }
    const double xid = xi_ideal ? xi_ideal[t] : xi_ideal_s;
    xi_out[t] = (mode == 1) ? s0 / (s0 - s1) : xid * s1 + (1 - xid) * s0;
}

// ---- transrot (transrot.f90:36-236) in three passes ----------------------------------------------
// pass 1: sums[t][0..8]  = sum m v (3), sum m q (3), sum m q x v (3)
// pass 2: sums[t][9..14] = inertia xx,xy,xz,yy,yz,zz about the (bugged) centre, needs pass 1
// pass 3: apply
__global__ void sp_transrot_sums1(const SplitArgs A, double* sums)
{
    const int na = A.natoms, nb = A.nbeads, rb = gridDim.x / A.ntraj, t = blockIdx.x / rb, bx = blockIdx.x - t * rb;
    const size_t nab = (size_t)na * nb;
    double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t e = (size_t)bx * blockDim.x + threadIdx.x; e < nab; e += (size_t)rb * blockDim.x) {
        const int atom = (int)(e % na);
        const size_t k = ((size_t)t * nab + e) * 3;
        const double w = A.mass[atom];
        const double qx = A.q[k], qy = A.q[k + 1], qz = A.q[k + 2];
        const double vx = A.p[k] / w, vy = A.p[k + 1] / w, vz = A.p[k + 2] / w;
        s[0] += vx * w;
        s[1] += vy * w;
        s[2] += vz * w;
        s[3] += qx * w;
        s[4] += qy * w;
        s[5] += qz * w;
        s[6] += (qy * vz - qz * vy) * w;
        s[7] += (qz * vx - qx * vz) * w;
        s[8] += (qx * vy - qy * vx) * w;
    }
    __shared__ double sh[9][4];
#pragma unroll
    for (int i = 0; i < 9; i++) {
        double v = s[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) sh[i][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        double v = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) v += sh[threadIdx.x][w];
        atomicAdd(&sums[(size_t)t * 16 + threadIdx.x], v);
    }
}
__device__ __forceinline__ void sp_centre(const SplitArgs& A, const double* s, double mt, double& totmass,
                                          double vtot[3], double ctr[3])
{
    totmass = (mt * A.nbeads) * A.nbeads;   // transrot.f90:62-79: multiplied by nbeads twice (F9)
    for (int d = 0; d < 3; d++) {
        vtot[d] = s[d] / totmass;
        ctr[d] = s[3 + d] / totmass;
    }
}
__global__ void sp_transrot_sums2(const SplitArgs A, double mt, double* sums)
{
    const int na = A.natoms, nb = A.nbeads, rb = gridDim.x / A.ntraj, t = blockIdx.x / rb, bx = blockIdx.x - t * rb;
    const size_t nab = (size_t)na * nb;
    double totmass, vtot[3], ctr[3];
    sp_centre(A, sums + (size_t)t * 16, mt, totmass, vtot, ctr);
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (size_t e = (size_t)bx * blockDim.x + threadIdx.x; e < nab; e += (size_t)rb * blockDim.x) {
        const int atom = (int)(e % na);
        const size_t k = ((size_t)t * nab + e) * 3;
        const double w = A.mass[atom];
        const double xd = A.q[k] - ctr[0], yd = A.q[k + 1] - ctr[1], zd = A.q[k + 2] - ctr[2];
        s[0] += xd * xd * w;
        s[1] += xd * yd * w;
        s[2] += xd * zd * w;
        s[3] += yd * yd * w;
        s[4] += yd * zd * w;
        s[5] += zd * zd * w;
    }
    __shared__ double sh[6][4];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double v = s[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) sh[i][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) v += sh[threadIdx.x][w];
        atomicAdd(&sums[(size_t)t * 16 + 9 + threadIdx.x], v);
    }
}

}  // namespace crcl
