// split_xi.cuh -- reaction coordinate, umbrella bias, SHAKE / RATTLE, Nose-Hoover chain and the
// recrossing bookkeeping for the HBM-resident ("split") path: any number of atoms and beads.
//
// Replaces, for systems beyond the in-register kernels of traj_kernel.cuh,
//   calc_xi.f90:108-502 + calc_com.f90   sp_xi_grad (value, gradient, hams force; Hessian-free as xi.cuh)
//   umbrella.f90:66-175                  sp_calc_xi_kernel (mode 1 + hams) + sp_add_bias
//   constrain_q.f90:30-112               sp_shake_solve (Newton on the multiplier) + sp_shake_apply
//   constrain_p.f90:30-75                sp_rattle
//   nhc.f90:34-170, mdinit.f90:126-146   sp_nhc, sp_nhc_init
//   recross_serial.f90:172-229           sp_recross_init, sp_recross_weights, sp_theta
// The reaction-coordinate work is O(natoms) per trajectory and runs one thread per trajectory (its
// work arrays live in a global scratch slab); everything that touches all beads is elementwise over
// (trajectory, bead, component) or one CTA per trajectory with a block reduction.
#pragma once
#include "split_kernels.cuh"
#include "xi.cuh"

namespace crcl {

constexpr int SPX_NWORK = 7;   // work arrays of 3*natoms doubles per trajectory

// calc_xi on one structure x[natoms][3].  mode 1: xi = s0/(s0-s1); mode 2: xi_ideal*s1+(1-xi_ideal)*s0.
// dxi[3n] always; hams[3n] (mode 1 only) if non-null.  w: SPX_NWORK*3n doubles of scratch.
__device__ inline void sp_xi_grad(const MechDev& M, int natoms, const double* __restrict__ mass, const double* x,
                                  double xi_ideal, int mode, double beta, double& xi, double* dxi, double* hams,
                                  double* w)
{
    const int nc = 3 * natoms;
    double *ds0 = w, *ds1 = w + nc, *v = w + 2 * nc, *H1v = w + 3 * nc, *H0v = w + 4 * nc;
    for (int t = 0; t < nc; t++) {
        ds0[t] = 0.0;
        ds1[t] = 0.0;
    }
    double Rf[8][3], Rb[8][3], fi[8], bi[8];
    double s1 = 0.0, s0u = 0.0;
    const double fnum = (double)M.form_num, bnum = (double)M.break_num;
    // ATOM_SHIFT has no bond terms, the unimolecular mechanisms no fragment terms (xi.cuh)
    const int nform = (M.type == 2) ? 0 : M.form_num, nbreak = (M.type == 2) ? 0 : M.break_num;
    const int nreac = (M.type == 0) ? M.sum_reacs : 0;
    for (int i = 0; i < nbreak; i++) {
        const int a1 = M.bb[i][0], a2 = M.bb[i][1];
        for (int d = 0; d < 3; d++) Rb[i][d] = x[3 * a1 + d] - x[3 * a2 + d];
        const double r = sqrt(Rb[i][0] * Rb[i][0] + Rb[i][1] * Rb[i][1] + Rb[i][2] * Rb[i][2]);
        bi[i] = 1.0 / r;
        s1 += (r - M.bref[i]) / bnum;
        s0u += (r - M.breac[i]) / bnum;
        for (int d = 0; d < 3; d++) {
            const double u = Rb[i][d] * bi[i] / bnum;
            ds1[3 * a1 + d] += u;
            ds1[3 * a2 + d] -= u;
        }
    }
    for (int i = 0; i < nform; i++) {
        const int a1 = M.bf[i][0], a2 = M.bf[i][1];
        for (int d = 0; d < 3; d++) Rf[i][d] = x[3 * a1 + d] - x[3 * a2 + d];
        const double r = sqrt(Rf[i][0] * Rf[i][0] + Rf[i][1] * Rf[i][1] + Rf[i][2] * Rf[i][2]);
        fi[i] = 1.0 / r;
        s1 -= (r - M.fref[i]) / fnum;
        s0u -= (r - M.freac[i]) / fnum;
        for (int d = 0; d < 3; d++) {
            const double u = Rf[i][d] * fi[i] / fnum;
            ds1[3 * a1 + d] -= u;
            ds1[3 * a2 + d] += u;
        }
    }
    double com[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int a = 0; a < natoms && nreac > 0; a++) {
        const int k = M.frag[a];
        if (k >= 0)
            for (int d = 0; d < 3; d++) com[k][d] += M.wfrag[a] * x[3 * a + d];
    }
    const double fterms = (double)((M.sum_reacs * M.sum_reacs - M.sum_reacs) / 2);
    double s0 = 0.0, Red[6][3], ri[6];
    int np = 0;
    for (int i = 0; i < nreac; i++)
        for (int j = i + 1; j < nreac; j++, np++) {
            for (int d = 0; d < 3; d++) Red[np][d] = com[j][d] - com[i][d];
            const double r = sqrt(Red[np][0] * Red[np][0] + Red[np][1] * Red[np][1] + Red[np][2] * Red[np][2]);
            ri[np] = 1.0 / r;
            s0 += M.R_inf - r;
            for (int a = 0; a < natoms; a++) {
                const int k = M.frag[a];
                if (k == i || k == j) {
                    const double sg = (k == i) ? 1.0 : -1.0;
                    const double u = sg * ri[np] * M.wfrag[a] / fterms;
                    for (int d = 0; d < 3; d++) ds0[3 * a + d] += Red[np][d] * u;
                }
            }
        }
    s0 = s0 / fterms;
    if (M.type == 1) {          // calc_xi.f90:722-728, :765
        s0 = s0u;
        for (int t = 0; t < nc; t++) ds0[t] = ds1[t];
    } else if (M.type == 2) {   // calc_xi.f90:523-620
        sp_shift_s(M, x, s0, s1);
        const double u = (M.shift_c2 < 0) ? 1.0 : 0.5;
        ds0[3 * M.shift_atom + M.shift_c1] = ds1[3 * M.shift_atom + M.shift_c1] = u;
        if (M.shift_c2 >= 0) ds0[3 * M.shift_atom + M.shift_c2] = ds1[3 * M.shift_atom + M.shift_c2] = u;
    }
    if (mode == 1) {
        const double D = s0 - s1;
        xi = s0 / D;
        const double iD2 = 1.0 / (D * D);
        for (int t = 0; t < nc; t++) dxi[t] = (s0 * ds1[t] - s1 * ds0[t]) * iD2;
    } else {
        xi = xi_ideal * s1 + (1 - xi_ideal) * s0;
        for (int t = 0; t < nc; t++) dxi[t] = xi_ideal * ds1[t] + (1 - xi_ideal) * ds0[t];
    }
    if (!hams) return;
    // hams[b] = coeff2/(coeff1 fs2) * sum_a d2xi(a,b) dxi_a/m_a without forming d2xi (xi.cuh)
    double fs2 = 0.0, d1v = 0.0, d0v = 0.0;
    for (int a = 0; a < natoms; a++)
        for (int d = 0; d < 3; d++) {
            const int t = 3 * a + d;
            v[t] = dxi[t] / mass[a];
            fs2 += dxi[t] * v[t];
            d1v += ds1[t] * v[t];
            d0v += ds0[t] * v[t];
            H1v[t] = 0.0;
            H0v[t] = 0.0;
        }
    for (int i = 0; i < nform; i++) {
        const int a1 = M.bf[i][0], a2 = M.bf[i][1];
        double u[3], o[3];
        for (int d = 0; d < 3; d++) u[d] = v[3 * a1 + d] - v[3 * a2 + d];
        proj(Rf[i], fi[i], u, o);
        for (int d = 0; d < 3; d++) {
            H1v[3 * a1 + d] -= o[d] / fnum;
            H1v[3 * a2 + d] += o[d] / fnum;
        }
    }
    for (int i = 0; i < nbreak; i++) {
        const int a1 = M.bb[i][0], a2 = M.bb[i][1];
        double u[3], o[3];
        for (int d = 0; d < 3; d++) u[d] = v[3 * a1 + d] - v[3 * a2 + d];
        proj(Rb[i], bi[i], u, o);
        for (int d = 0; d < 3; d++) {
            H1v[3 * a1 + d] += o[d] / bnum;
            H1v[3 * a2 + d] -= o[d] / bnum;
        }
    }
    np = 0;
    for (int i = 0; i < nreac; i++)
        for (int j = i + 1; j < nreac; j++, np++) {
            double W[3] = {0, 0, 0}, o[3];
            for (int a = 0; a < natoms; a++) {
                const int k = M.frag[a];
                if (k == i || k == j) {
                    const double u = ((k == i) ? 1.0 : -1.0) * M.wfrag[a];
                    for (int d = 0; d < 3; d++) W[d] += u * v[3 * a + d];
                }
            }
            proj(Red[np], ri[np], W, o);
            for (int a = 0; a < natoms; a++) {
                const int k = M.frag[a];
                if (k == i || k == j) {
                    const double u = -((k == i) ? 1.0 : -1.0) * M.wfrag[a] / fterms;
                    for (int d = 0; d < 3; d++) H0v[3 * a + d] += u * o[d];
                }
            }
        }
    if (M.type == 1)
        for (int t = 0; t < nc; t++) H0v[t] = H1v[t];   // d2s0 = d2s1 (calc_xi.f90:913)
    const double coeff1 = 2.0 * PI_UMBR * beta;
    fs2 = fs2 / coeff1;
    const double pref = (-1.0 / beta) / (coeff1 * fs2);
    const double D = s0 - s1, iD3 = 1.0 / (D * D * D);
    const double cross2 = 2.0 * (s0 * d1v - s1 * d0v);
    for (int t = 0; t < nc; t++) {
        const double hv = ((s0 * H1v[t] + ds0[t] * d1v - ds1[t] * d0v - s1 * H0v[t]) * D - cross2 * (ds0[t] - ds1[t])) * iD3;
        hams[t] = hv * pref;
    }
}

// per-trajectory scalars of the biased / constrained modes
struct SpTraj {
    int ntraj, natoms, nbeads;
    double dt, beta;
    const double* mass;
    const int* at_move;
    const double* xi_ideal;   // [ntraj] or null -> xi_ideal_s
    const double* k_force;    // [ntraj] or null -> k_force_s
    double xi_ideal_s, k_force_s;
    double* xi_real;          // [ntraj]
    double* dxi;              // [ntraj][3n]
    double* hams;             // [ntraj][3n]
    double* work;             // [ntraj][SPX_NWORK][3n]
    int* status;              // [ntraj]
    int* bad;                 // [ntraj]: SHAKE failed in this step
    double* coeff;            // [ntraj][2]: coeff, mult of constrain_q
};

// umbrella.f90:110-132 (modes) on the centroid of every trajectory
__global__ void sp_calc_xi_kernel(MechDev M, SpTraj S, const double* __restrict__ cen, int mode, int want_hams)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S.ntraj) return;
    const int nc = 3 * S.natoms;
    double xi;
    sp_xi_grad(M, S.natoms, S.mass, cen + (size_t)t * nc, S.xi_ideal ? S.xi_ideal[t] : S.xi_ideal_s, mode, S.beta, xi,
               S.dxi + (size_t)t * nc, want_hams ? S.hams + (size_t)t * nc : nullptr,
               S.work + (size_t)t * SPX_NWORK * nc);
    S.xi_real[t] = xi;
}

// umbrella.f90:136-174: g += k (xi - xi0) dxi on every bead, then the hams force, in that order
__global__ void sp_add_bias(SpTraj S, double* __restrict__ g)
{
    const int nc = 3 * S.natoms;
    const size_t per = (size_t)S.nbeads * nc;
    const int tiles = gridDim.x / S.ntraj, t = blockIdx.x / tiles;
    const size_t i = (size_t)(blockIdx.x - t * tiles) * blockDim.x + threadIdx.x;
    if (i >= per) return;
    const int c = (int)(i % nc);
    const double xi0 = S.xi_ideal ? S.xi_ideal[t] : S.xi_ideal_s, kf = S.k_force ? S.k_force[t] : S.k_force_s;
    const double kd = kf * (S.xi_real[t] - xi0);
    const size_t k = (size_t)t * per + i;
    g[k] = (g[k] + kd * S.dxi[(size_t)t * nc + c]) + S.hams[(size_t)t * nc + c];
}

// constrain_q.f90:62-99: Newton iteration on the multiplier with the previous step's dxi
__global__ void sp_shake_solve(MechDev M, SpTraj S, const double* __restrict__ cen)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S.ntraj) return;
    const int na = S.natoms, nc = 3 * na, nb = S.nbeads;
    const double* x = cen + (size_t)t * nc;
    const double* d = S.dxi + (size_t)t * nc;
    double* w = S.work + (size_t)t * SPX_NWORK * nc;
    double *xt = w + 5 * nc, *dn = w + 6 * nc;
    const double dt = S.dt, xid = S.xi_ideal ? S.xi_ideal[t] : S.xi_ideal_s;
    double mult = 0.0, coeff = 0.0;
    int ok = 0;
    for (int iter = 1; iter <= 200; iter++) {
        coeff = mult * dt * dt / nb;
        for (int j = 0; j < na; j++)
            for (int k = 0; k < 3; k++) xt[3 * j + k] = x[3 * j + k] + coeff * d[3 * j + k] / S.mass[j];
        double xin;
        sp_xi_grad(M, na, S.mass, xt, xid, 2, S.beta, xin, dn, nullptr, w);
        double dsigma = 0.0;
        for (int k = 0; k < 3; k++)
            for (int j = 0; j < na; j++) dsigma += dn[3 * j + k] * dt * dt * d[3 * j + k] / (S.mass[j] * nb);
        const double dx = xin / dsigma;
        mult -= dx;
        if (fabs(dx) < (double)1.0E-8f || fabs(xin) < (double)1.0E-10f) {   // REAL*4 literals (constrain_q.f90:93)
            ok = 1;
            break;
        }
    }
    S.coeff[2 * t] = ok ? coeff : 0.0;
    S.coeff[2 * t + 1] = ok ? mult : 0.0;
    S.bad[t] = !ok;
    if (!ok) atomicOr(&S.status[t], CRCL_TRAJ_SHAKE_FAIL);
}
// constrain_q.f90:102-109
__global__ void sp_shake_apply(SpTraj S, double* __restrict__ q, double* __restrict__ p)
{
    const int nc = 3 * S.natoms;
    const size_t per = (size_t)S.nbeads * nc;
    const int tiles = gridDim.x / S.ntraj, t = blockIdx.x / tiles;
    const size_t i = (size_t)(blockIdx.x - t * tiles) * blockDim.x + threadIdx.x;
    if (i >= per || S.bad[t]) return;
    const int c = (int)(i % nc);
    const double dx = S.dxi[(size_t)t * nc + c], coeff = S.coeff[2 * t], mult = S.coeff[2 * t + 1];
    const size_t k = (size_t)t * per + i;
    q[k] = q[k] + coeff / S.mass[c / 3] * dx;
    p[k] = p[k] + mult * S.dt / S.nbeads * dx;
}

__device__ __forceinline__ double sp_block_sum(double v, double* sh)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += sh[w];
    return s;
}

// constrain_p.f90:30-75: one CTA per trajectory
__global__ void __launch_bounds__(256) sp_rattle(SpTraj S, double* __restrict__ p)
{
    __shared__ double sh[8];
    const int t = blockIdx.x, nc = 3 * S.natoms;
    const size_t per = (size_t)S.nbeads * nc;
    const double* d = S.dxi + (size_t)t * nc;
    double* pt = p + (size_t)t * per;
    double c1 = 0.0, c2 = 0.0;
    for (size_t i = threadIdx.x; i < per; i += blockDim.x) {
        const int c = (int)(i % nc);
        c1 += d[c] * pt[i] / S.mass[c / 3];
    }
    for (int c = threadIdx.x; c < nc; c += blockDim.x) c2 += d[c] * d[c] / S.mass[c / 3];
    c1 = sp_block_sum(c1, sh);
    c2 = sp_block_sum(c2, sh);
    const double lam = -c1 / c2 / S.nbeads;
    for (size_t i = threadIdx.x; i < per; i += blockDim.x) pt[i] = pt[i] + lam * d[i % nc];
}

// nhc.f90:34-170: one CTA per trajectory; nhc[t][8] = vnh[4], qnh[4]
__global__ void __launch_bounds__(256) sp_nhc(SpTraj S, double* __restrict__ p, double* __restrict__ nhc, double kelvin,
                                              int nfree)
{
    __shared__ double sh[8];
    __shared__ double s_scale;
    const int t = blockIdx.x, nc = 3 * S.natoms, NB = S.nbeads;
    const size_t per = (size_t)NB * nc;
    double* pt = p + (size_t)t * per;
    double ek = 0.0;
    for (size_t i = threadIdx.x; i < per; i += blockDim.x) {
        const int a = (int)(i % nc) / 3;
        if (S.at_move[a]) ek += pt[i] * pt[i] / (2.0 * S.mass[a]) / NB / NB;
    }
    double eksum = sp_block_sum(ek, sh);
    if (threadIdx.x == 0) {
        constexpr float ektf = 1.380649E-23f / 4.3597447E-18f;   // REAL*4 division, nhc.f90:51
        const double ekt = (double)ektf * kelvin, dtc = S.dt / 5.0, nf = (double)nfree;
        double w[3], vnh[4], qnh[4], scale = 1.0, gn;
        w[0] = 1.0 / (2.0 - cbrt(2.0));
        w[1] = 1.0 - 2.0 * w[0];
        w[2] = w[0];
        for (int i = 0; i < 4; i++) {
            vnh[i] = nhc[(size_t)t * 8 + i];
            qnh[i] = nhc[(size_t)t * 8 + 4 + i];
        }
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
        for (int i = 0; i < 4; i++) nhc[(size_t)t * 8 + i] = vnh[i];
        s_scale = scale;
    }
    __syncthreads();
    const double scale = s_scale;
    for (size_t i = threadIdx.x; i < per; i += blockDim.x) {
        const int a = (int)(i % nc) / 3;
        pt[i] = S.at_move[a] ? scale * pt[i] : 0.0;
    }
}
// mdinit.f90:126-146
__global__ void sp_nhc_init(int ntraj, double* __restrict__ nhc, double kelvin, double nose_q, int nfree)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntraj) return;
    const double qterm = 0.316679e-5 * kelvin * nose_q * nose_q;
    for (int j = 0; j < 4; j++) {
        nhc[(size_t)t * 8 + 4 + j] = qterm;
        nhc[(size_t)t * 8 + j] = 0.0;
    }
    nhc[(size_t)t * 8 + 4] = (double)nfree * qterm;
}

// centroid of every trajectory (get_centroid.f90:67-82), beads summed in order
__global__ void sp_centroid(int ntraj, int natoms, int nbeads, const double* __restrict__ q, double* __restrict__ cen)
{
    const int nc = 3 * natoms;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)ntraj * nc) return;
    const size_t t = i / nc, c = i - t * nc;
    double s = 0.0;
    for (int b = 0; b < nbeads; b++) s += q[(t * nbeads + b) * nc + c];
    cen[i] = s / nbeads;
}

// sum xi, sum xi^2 over the steps of an umbrella sampling launch
__global__ void sp_accum_xi(int ntraj, const double* __restrict__ xr, double* __restrict__ s1, double* __restrict__ s2)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntraj) return;
    s1[t] += xr[t];
    s2[t] += xr[t] * xr[t];
}
// epot penalty of a failed SHAKE (verlet.f90:749-755)
// rpmd_check.f90:88-116 after a step of the biased / constrained modes (see Traj::step in traj_kernel.cuh)
__global__ void sp_rpmd_check(int ntraj, const double* __restrict__ epot, const double* __restrict__ xi_real,
                              const double* __restrict__ xi_ideal, double xi_ideal_s, double emax, double xi_tol,
                              int test_xi, int* __restrict__ status)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntraj) return;
    const double e = epot[t];
    int st = 0;
    if (e != e || e > 1.79769313486231570815e308) st |= CRCL_TRAJ_NAN;
    if (e > emax) st |= CRCL_TRAJ_ENERGY;
    if (test_xi && fabs(xi_real[t] - (xi_ideal ? xi_ideal[t] : xi_ideal_s)) > xi_tol) st |= CRCL_TRAJ_XI_RANGE;
    if (st) atomicOr(&status[t], st);
}

__global__ void sp_shake_penalty(int ntraj, const int* __restrict__ bad, double* __restrict__ epot)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ntraj && bad[t]) epot[t] += 100000.0;
}

// ---- recrossing children on the split path (recross_serial.f90:172-229) -------------------------
// trajectory t = 2 g + k: child k of pair pair0 + g starts from parent (pair mod nparent); both children
// of a pair draw the same momenta (stream keyed by the pair index), the second with the sign flipped
__global__ void sp_recross_init(int ntraj, int natoms, int nbeads, int nparent, int pair0,
                                const double* __restrict__ qpar, double* __restrict__ q, uint32_t* __restrict__ tid,
                                uint32_t* __restrict__ event)
{
    const size_t per = (size_t)nbeads * 3 * natoms;
    const int tiles = gridDim.x / ntraj, t = blockIdx.x / tiles;
    const size_t i = (size_t)(blockIdx.x - t * tiles) * blockDim.x + threadIdx.x;
    if (i >= per) return;
    const int pair = pair0 + (t >> 1);
    q[(size_t)t * per + i] = qpar[(size_t)(pair % nparent) * per + i];
    if (i == 0) {
        tid[t] = (uint32_t)pair;
        event[t] = 0u;
    }
}
__global__ void sp_flip_odd(int ntraj, size_t per, double* __restrict__ p)
{
    const int tiles = gridDim.x / ntraj, t = blockIdx.x / tiles;
    const size_t i = (size_t)(blockIdx.x - t * tiles) * blockDim.x + threadIdx.x;
    if (i < per && (t & 1)) p[(size_t)t * per + i] = -p[(size_t)t * per + i];
}
// v_s = sum dxi p / m / nbeads, f_s = sqrt(sum dxi^2/m / (2 pi beta)): one CTA per trajectory
__global__ void __launch_bounds__(256) sp_recross_weights(SpTraj S, const double* __restrict__ p, double* __restrict__ weight,
                                                          double* __restrict__ denom_part)
{
    __shared__ double sh[8];
    const int t = blockIdx.x, nc = 3 * S.natoms;
    const size_t per = (size_t)S.nbeads * nc;
    const double* d = S.dxi + (size_t)t * nc;
    const double* pt = p + (size_t)t * per;
    double vs = 0.0, fs = 0.0;
    for (size_t i = threadIdx.x; i < per; i += blockDim.x) {
        const int c = (int)(i % nc);
        vs += d[c] * pt[i] / S.mass[c / 3];
    }
    for (int c = threadIdx.x; c < nc; c += blockDim.x) fs += d[c] * d[c] / S.mass[c / 3];
    vs = sp_block_sum(vs, sh) / S.nbeads;
    fs = sqrt(sp_block_sum(fs, sh) / (2.0 * PI_UMBR * S.beta));
    if (threadIdx.x == 0) {
        const double w = vs / fs;
        weight[t] = w;
        denom_part[t] = (vs > 0) ? w : 0.0;
    }
}
// step_ctr (may be null): device-resident step index, so that a CUDA-graph replay of a step writes its own row
__global__ void sp_theta(int ntraj, const double* __restrict__ xr, unsigned char* __restrict__ theta,
                         const uint32_t* __restrict__ step_ctr)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (step_ctr) theta += (size_t)step_ctr[0] * ntraj;
    if (t < ntraj) theta[t] = (xr[t] > 0) ? 1 : 0;
}

}  // namespace crcl
