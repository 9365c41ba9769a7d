// qmdff_kernels.cu -- QMDFF energy and gradient on the device.
//
// Replaces, for one QMDFF (gradient.f90:341-362, nqmdff = 1):
//   ff_eg.f90:40-629    bonds, angles (cosine / linear form, abdamp.f90 damping), proper torsions
//                       (<= 4 cosine terms, erf switch, valijkl.f90 + dphidr.f90) and inversions
//                       (omega.f90 + domegadr.f90)                      -> qm_bonded_kernel
//   ff_nonb.f90:88-193,339-417   nci pair list: D3-BJ-like dispersion, exponential repulsion,
//                       Coulomb (Zahn / cut-off + exp_switch.f90 / plain)  -> qm_nci_kernel
//   ff_nonb.f90:198-332,421-512  inter-molecular O(N^2) loops             -> qm_inter_kernel
//   ff_hb.f90:35-280    hb list (eabhag.f90 analytic / eabxag.f90 + eabx.f90 numeric) and the
//                       on-the-fly donor x acceptor search for nmols > 1   -> qm_hb_list_kernel,
//                                                                             qm_hb_search_kernel
// The SPME/Ewald branch of ff_nonb is dead code in the reference (ewald=.false., :337) and has
// no counterpart.
//
// Mapping: one thread per (image, term) for the lists -- term-major with the image running fastest when a
// call carries many images of a small molecule --, gradients accumulated with FP64 atomics (red.global.add.f64; a term
// touches 2-4 atoms).  The inter-molecular part visits every unordered pair once: qm_inter_kernel, a warp per
// (image, atom i) sweeping j > i with an FP32 pre-filter on a packed copy and a per-warp queue of surviving pairs that is
// evaluated 32 at a time by qm_pair_exact (exact box_image loop and cut-off tests in FP64); for periodic boxes of at least
// 2.5 cut-offs, qm_cellsort_kernel + qm_inter_cell_kernel: atoms binned and sorted by cell, a CTA per home cell streaming
// the half shell of neighbour cells as contiguous runs, the same exact stage.  The donor x acceptor search of ff_hb uses
// the warp-queue scheme per (image, donor, 512-atom segment).
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <vector>
#include "qmdff.cuh"
#include "crcl_common.cuh"

namespace crcl {

__device__ __forceinline__ void box_image(const QmdffDev& D, double v[3])
{
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const double L = D.box[d], L2 = 0.5 * L;
        while (fabs(v[d]) > L2) v[d] -= (v[d] >= 0.0) ? L : -L;
    }
}
__device__ __forceinline__ void cross3(const double a[3], const double b[3], double c[3])
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3(const double a[3], const double b[3])
{
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
// abdamp.f90:35-50; rcut = 3.0*3.5710642*(rad_i+rad_j)^2 with REAL*4 literals (F3)
__device__ __forceinline__ void abdamp(const QmdffDev& D, int ti, int tj, double r2, double& damp, double& ddamp)
{
    const double rs = D.rad[ti] + D.rad[tj];
    const double rcut = (double)(3.0f * 3.5710642f) * (rs * rs);
    const double rr = (r2 / rcut) * (r2 / rcut);
    damp = 1.0 / (1.0 + rr);
    ddamp = -4.0 * rr / (r2 * ((1.0 + rr) * (1.0 + rr)));
}
__device__ __forceinline__ void ld3(const double* x, int a, double v[3])
{
    v[0] = x[3 * a];
    v[1] = x[3 * a + 1];
    v[2] = x[3 * a + 2];
}
__device__ __forceinline__ void add3(double* g, int a, const double v[3])
{
    atomicAdd(&g[3 * a], v[0]);
    atomicAdd(&g[3 * a + 1], v[1]);
    atomicAdd(&g[3 * a + 2], v[2]);
}
__device__ __forceinline__ double block_sum_to(double v, double* dst)
{
    __shared__ double sh[8];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += sh[w];
        if (t != 0.0) atomicAdd(dst, t);
    }
    __syncthreads();
    return v;
}

// (image, term) of this thread for the list kernels.  Two launch shapes: grid (term tiles, images) with
// a block reduction of the energy per image (large lists), or -- when a list is shorter than a CTA
// (gas-phase molecules x thousands of ring-polymer beads) -- one flat index over (image, term) so
// that CTAs stay full; the energy then goes out with one atomic per term.
__device__ __forceinline__ void term_index(int nterm, int nimg, int flat, int& t, int& img)
{
    if (flat == 2) {
        // term-major: the 32 threads of a warp evaluate the SAME term on 32 consecutive images -- one term type per warp
        // (bonds, angles and torsions of a small molecule no longer share warps and serialise their three bodies) and the
        // gradient reductions of a warp-instruction go to 32 different images instead of colliding on the few atoms of one
        const size_t gidx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        const size_t tt = gidx / (size_t)nimg;
        t = (tt < (size_t)nterm) ? (int)tt : nterm;
        img = (tt < (size_t)nterm) ? (int)(gidx - tt * nimg) : 0;
    } else if (flat) {
        const size_t gidx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        const size_t im = gidx / (size_t)nterm;
        t = (im < (size_t)nimg) ? (int)(gidx - im * nterm) : nterm;
        img = (im < (size_t)nimg) ? (int)im : 0;
    } else {
        t = blockIdx.x * blockDim.x + threadIdx.x;
        img = blockIdx.y;
    }
}
__device__ __forceinline__ void term_energy(double e, double* dst, int flat)
{
    if (flat) {
        if (e != 0.0) atomicAdd(dst, e);
    } else {
        block_sum_to(e, dst);
    }
}

// ---- bonded terms: term index t in [0, nbond+nangl+ntors) -------------------------------------------
__global__ void __launch_bounds__(128) qm_bonded_kernel(const QmdffDev D, const double* __restrict__ xyz,
                                                        double* __restrict__ V, double* __restrict__ g, int nimg,
                                                        int flat)
{
    constexpr double PI = 3.1415926535897932384626433832795029, PI2 = 6.28318530717958623199592693708837,
                     SPI = 1.77245385090551599275151910313925;
    const int nterm = D.nbond + D.nangl + D.ntors;
    int t, img;
    term_index(nterm, nimg, flat, t, img);
    const double* x = xyz + (size_t)img * 3 * D.n;
    double* gi = g + (size_t)img * 3 * D.n;
    double e = 0.0;
    if (t < D.nbond) {
        const int i = D.bond[2 * t], j = D.bond[2 * t + 1];
        double a[3], b[3], rb[3];
        ld3(x, i, a);
        ld3(x, j, b);
        for (int c = 0; c < 3; c++) rb[c] = a[c] - b[c];
        if (D.periodic) box_image(D, rb);
        const double r2 = dot3(rb, rb), r = sqrt(r2);
        const double rij = D.vbond[3 * t], kij = D.vbond[3 * t + 1], aai = D.vbond[3 * t + 2];
        const double ph = CRCL_POW(CRCL_DIV(rij, r), 0.5 * aai), pf = ph * ph;   // (rij/r)^(a/2), (rij/r)^a
        e = kij * (1.0 + pf - 2.0 * ph);
        const double fac = aai * kij * (-pf + ph) / r2;
        double v[3] = {fac * rb[0], fac * rb[1], fac * rb[2]};
        add3(gi, i, v);
        v[0] = -v[0];
        v[1] = -v[1];
        v[2] = -v[2];
        add3(gi, j, v);
    } else if (t < D.nbond + D.nangl) {
        const int m = t - D.nbond;
        const int j = D.angl[3 * m], i = D.angl[3 * m + 1], k = D.angl[3 * m + 2];
        const double c0 = D.vangl[2 * m], kijk = D.vangl[2 * m + 1];
        double va[3], vb[3], vc[3], vab[3], vcb[3], vp[3];
        ld3(x, i, va);
        ld3(x, j, vb);
        ld3(x, k, vc);
        for (int c = 0; c < 3; c++) {
            vab[c] = va[c] - vb[c];
            vcb[c] = vc[c] - vb[c];
        }
        if (D.periodic) {
            box_image(D, vab);
            box_image(D, vcb);
        }
        const double rab2 = dot3(vab, vab), rcb2 = dot3(vcb, vcb);
        cross3(vcb, vab, vp);
        const double rp = sqrt(dot3(vp, vp)) + 1.e-14;
        const double al = sqrt(rab2), bl = sqrt(rcb2);
        double cosa = (al > 0.0 && bl > 0.0) ? dot3(vab, vcb) / (al * bl) : 0.0;
        cosa = fmin(1.0, fmax(-1.0, cosa));
        const double theta = CRCL_ACOS(cosa);
        double dij, d2ij, djk, d2jk;
        abdamp(D, D.type[i], D.type[j], rab2, dij, d2ij);
        abdamp(D, D.type[k], D.type[j], rcb2, djk, d2jk);
        const double damp = dij * djk;
        double ea, deddt;
        if (PI - c0 < 1.e-6) {
            const double dt = theta - c0;
            ea = kijk * dt * dt;
            deddt = 2.0 * kijk * dt;
        } else {
            const double cc = cos(c0);
            ea = kijk * (cosa - cc) * (cosa - cc);
            deddt = 2.0 * kijk * sin(theta) * (cc - cosa);
        }
        e = ea * damp;
        double deda[3], dedc[3];
        cross3(vab, vp, deda);
        cross3(vcb, vp, dedc);
        const double rm1 = -deddt / (rab2 * rp), rm2 = deddt / (rcb2 * rp);
        double ga[3], gb[3], gc[3];
        for (int c = 0; c < 3; c++) {
            const double da = deda[c] * rm1, dc = dedc[c] * rm2;
            const double t1 = ea * d2ij * djk * vab[c], t2 = ea * d2jk * dij * vcb[c];
            ga[c] = da * damp + t1;
            gc[c] = dc * damp + t2;
            gb[c] = -(da + dc) * damp - t1 - t2;
        }
        add3(gi, i, ga);
        add3(gi, j, gb);
        add3(gi, k, gc);
    } else if (t < nterm) {
        const int m = t - D.nbond - D.nangl;
        const int* tr = D.tors + 6 * m;
        const double* vt = D.vtors + (size_t)D.ldvt * m;
        const int i = tr[0], j = tr[1], k = tr[2], l = tr[3], nt = tr[4];
        const double phi0 = vt[0];
        double xi[3], xj[3], xk[3], xl[3];
        ld3(x, i, xi);
        ld3(x, j, xj);
        ld3(x, k, xk);
        ld3(x, l, xl);
        double gA[3], gB[3], gC[3], gD[3];
        if (tr[5] != 2) {
            // proper torsion: ra = j-i, rb = k-j, rc = l-k (valijkl.f90, dphidr.f90)
            double ra[3], rb[3], rc[3];
            for (int c = 0; c < 3; c++) {
                ra[c] = xj[c] - xi[c];
                rb[c] = xk[c] - xj[c];
                rc[c] = xl[c] - xk[c];
            }
            if (D.periodic) {
                box_image(D, ra);
                box_image(D, rb);
                box_image(D, rc);
            }
            const double rij = dot3(ra, ra), rjk = dot3(rb, rb), rkl = dot3(rc, rc);
            double dij, d2ij, djk, d2jk, dkl, d2kl;
            abdamp(D, D.type[i], D.type[j], rij, dij, d2ij);
            abdamp(D, D.type[k], D.type[j], rjk, djk, d2jk);
            abdamp(D, D.type[k], D.type[l], rkl, dkl, d2kl);
            const double damp = djk * dij * dkl;
            double na[3], nb[3];
            cross3(ra, rb, na);
            cross3(rb, rc, nb);
            const double nan_ = sqrt(dot3(na, na)), nbn = sqrt(dot3(nb, nb));
            double sn = dot3(na, nb);
            if (nan_ > 1.e-14) sn /= nan_;
            if (nbn > 1.e-14) sn /= nbn;
            if (fabs(fabs(sn) - 1.0) < 1.0e-14) sn = (sn >= 0.0) ? 1.0 : -1.0;
            const double phi = CRCL_ACOS(sn);
            const double cosphi = cos(phi), sinphi = sin(phi);
            double dda[3] = {0, 0, 0}, ddb[3] = {0, 0, 0}, ddc[3] = {0, 0, 0}, ddd[3] = {0, 0, 0};
            const double nenner = nan_ * nbn * sinphi;
            if (!(fabs(nenner) < 1.e-14)) {
                const double on = 1.0 / nenner;
                double rapb[3], rbpc[3], rab[3], rba[3], rac[3], rbb[3], rbc[3], raa[3], rapba[3], rapbb[3],
                    rbpca[3], rbpcb[3];
                for (int c = 0; c < 3; c++) {
                    rapb[c] = ra[c] + rb[c];
                    rbpc[c] = rb[c] + rc[c];
                }
                cross3(na, rb, rab);
                cross3(nb, ra, rba);
                cross3(na, rc, rac);
                cross3(nb, rb, rbb);
                cross3(nb, rc, rbc);
                cross3(na, ra, raa);
                cross3(rapb, na, rapba);
                cross3(rapb, nb, rapbb);
                cross3(rbpc, na, rbpca);
                cross3(rbpc, nb, rbpcb);
                const double ba = nbn / nan_, ab = nan_ / nbn;
                for (int c = 0; c < 3; c++) {
                    dda[c] = on * (cosphi * ba * rab[c] - rbb[c]);
                    ddb[c] = on * (cosphi * (ba * rapba[c] + ab * rbc[c]) - (rac[c] + rapbb[c]));
                    ddc[c] = on * (cosphi * (ba * raa[c] + ab * rbpcb[c]) - (rba[c] + rbpca[c]));
                    ddd[c] = on * (cosphi * ab * rbb[c] - rab[c]);
                }
            }
            double et = 0.0, dd = 0.0;
            const double phipi = phi - PI, ef = erf(phipi), expo = exp(-phipi * phipi) / SPI;
            for (int it = 0; it < nt; it++) {
                const double rn = vt[2 + 3 * it], ph = vt[3 + 3 * it], vv = vt[4 + 3 * it];
                const double c1 = rn * (phi - phi0) + ph, c2 = rn * (phi + phi0 - PI2) + ph;
                double s1, co1, s2, co2;
                sincos(c1, &s1, &co1);
                sincos(c2, &s2, &co2);
                const double e1 = vv * (1.0 + co1), e2 = vv * (1.0 + co2);
                et += 0.5 * (1.0 - ef) * e1 + (0.5 + 0.5 * ef) * e2;
                dd += -expo * e1 - 0.5 * (1.0 - ef) * vv * s1 * rn + expo * e2 - (0.5 + 0.5 * ef) * vv * s2 * rn;
            }
            et *= vt[1];
            dd *= vt[1] * damp;
            e = et * damp;
            // the reference's vab = i-j = -ra, vcb = j-k = -rb, vdc = k-l = -rc
            for (int c = 0; c < 3; c++) {
                const double t1 = et * d2ij * djk * dkl * (-ra[c]);
                const double t2 = et * d2jk * dij * dkl * (-rb[c]);
                const double t3 = et * d2kl * dij * djk * (-rc[c]);
                gA[c] = dd * dda[c] + t1;
                gB[c] = dd * ddb[c] - t1 + t2;
                gC[c] = dd * ddc[c] + t3 - t2;
                gD[c] = dd * ddd[c] - t3;
            }
        } else {
            // inversion at centre j (omega.f90, domegadr.f90): re = i-j, rd = k-j, rv = l-i
            double re[3], rd[3], rv[3], vdl[3];
            for (int c = 0; c < 3; c++) {
                re[c] = xi[c] - xj[c];
                rd[c] = xk[c] - xj[c];
                rv[c] = xl[c] - xi[c];
                vdl[c] = xj[c] - xl[c];
            }
            if (D.periodic) {
                box_image(D, re);
                box_image(D, rd);
                box_image(D, rv);
                box_image(D, vdl);
            }
            const double rij = dot3(re, re), rjk = dot3(rd, rd), rjl = dot3(vdl, vdl);
            double dij, d2ij, djk, d2jk, djl, d2jl;
            abdamp(D, D.type[i], D.type[j], rij, dij, d2ij);
            abdamp(D, D.type[k], D.type[j], rjk, djk, d2jk);
            abdamp(D, D.type[j], D.type[l], rjl, djl, d2jl);
            const double damp = djk * dij * djl;
            double rn[3];
            cross3(re, rd, rn);
            const double rnn = sqrt(dot3(rn, rn)), rvn = sqrt(dot3(rv, rv));
            double sarg = dot3(rn, rv);
            if (rnn > 1.e-14) sarg /= rnn;
            if (rvn > 1.e-14) sarg /= rvn;
            const double om = asin(sarg);
            const double sinom = sin(om);
            double dda[3] = {0, 0, 0}, ddb[3] = {0, 0, 0}, ddc[3] = {0, 0, 0}, ddd[3] = {0, 0, 0};
            const double nenner = rnn * rvn * cos(om);
            if (fabs(nenner) > 1.e-14) {
                const double on = 1.0 / nenner;
                double rdme[3], rve[3], rne[3], rdv[3], rdn[3], rvdme[3], rndme[3];
                for (int c = 0; c < 3; c++) rdme[c] = rd[c] - re[c];
                cross3(rv, re, rve);
                cross3(rn, re, rne);
                cross3(rd, rv, rdv);
                cross3(rd, rn, rdn);
                cross3(rv, rdme, rvdme);
                cross3(rn, rdme, rndme);
                const double vn = rvn / rnn, nv = rnn / rvn;
                for (int c = 0; c < 3; c++) {
                    dda[c] = on * (rdv[c] - rn[c] - sinom * (vn * rdn[c] - nv * rv[c]));
                    ddb[c] = on * (rvdme[c] - sinom * vn * rndme[c]);
                    ddc[c] = on * (rve[c] - sinom * vn * rne[c]);
                    ddd[c] = on * (rn[c] - sinom * nv * rv[c]);
                }
            }
            double et, dd;
            if (vt[2] > 1.e-6) {
                const double c1 = (om - phi0) + PI;
                et = (1.0 + cos(c1)) * vt[1];
                dd = -sin(c1) * vt[1] * damp;
            } else {
                const double cd = cos(om) - cos(phi0);
                et = vt[1] * cd * cd;
                dd = 2.0 * vt[1] * sin(om) * (-cd) * damp;
            }
            e = et * damp;
            // the reference's vab = j-i = -re, vcb = j-k = -rd, vdc = j-l = vdl
            for (int c = 0; c < 3; c++) {
                const double t1 = et * d2ij * djk * djl * (-re[c]);
                const double t2 = et * d2jk * dij * djl * (-rd[c]);
                const double t3 = et * d2jl * dij * djk * vdl[c];
                gA[c] = dd * dda[c] - t1;
                gB[c] = dd * ddb[c] + t1 + t2 + t3;
                gC[c] = dd * ddc[c] - t2;
                gD[c] = dd * ddd[c] - t3;
            }
        }
        add3(gi, i, gA);
        add3(gi, j, gB);
        add3(gi, k, gC);
        add3(gi, l, gD);
    }
    term_energy(e, &V[img], flat);
}

// dispersion + repulsion of one pair (ff_nonb.f90:120-160): returns the energy, dr = gradient factor
// such that g_i1 += vab*dr, g_i2 -= vab*dr
__device__ __forceinline__ double vdw_pair(const QmdffDev& D, int t1, int t2, double c6, double r2, double r,
                                           double eps, double& dr)
{
    const double R0 = D.r094[t1][t2];
    const double r4 = r2 * r2, r6 = r4 * r2;
    const double R02 = R0 * R0, r06 = R02 * R02 * R02;
    const double t6 = r6 + r06, t8 = r6 * r2 + r06 * R02;
    const double c6t6 = c6 / t6, c6t8 = c6 / t8;
    const double t27 = D.sr42[t1][t2] * c6t8;
    double e = -(c6t6 + t27) * eps;
    dr = eps * (c6t6 * 6.0 * r4 / t6 + 8.0 * t27 * r6 / t8);
    if (r < 25.0) {
        const double alpha = D.r0ab[t1][t2];
        const double tt = D.zab[t1][t2] * CRCL_EXP(-alpha * r);
        const double oner = 1.0 / r;
        e += tt * oner * eps;
        dr -= eps * tt * (alpha * r + 1.0) * oner / r2;
    }
    return e;
}
// Coulomb of one pair (ff_nonb.f90:339-417): energy; gradient factor dr (g_i1 += vab*dr)
__device__ __forceinline__ double coul_pair(const QmdffDev& D, double qq, double r2, double r, double eps, double& dr)
{
    dr = 0.0;
    double sw = 1.0;
    if (D.periodic) {
        if (r > D.coul_cut) return 0.0;
        if (!D.zahn && r > D.cut_low) {
            const double xv = (r - D.cut_low) / (D.coul_cut - D.cut_low);
            sw = exp(1.0) * exp(1.0 / (xv - 1.0));
        }
    }
    if (r > D.coul_cut) return 0.0;
    const double oner = 1.0 / r;
    const double e0 = D.zahn ? qq * (erfc(D.zahn_a * r) * oner - D.zahn_par * (r - D.coul_cut)) : qq * oner * eps * sw;
    dr = -e0 / r2;   // the reference uses e0/r^2 for every Coulomb form (ff_nonb.f90:384,470)
    return e0;
}

__global__ void __launch_bounds__(128) qm_nci_kernel(const QmdffDev D, const double* __restrict__ xyz,
                                                     double* __restrict__ V, double* __restrict__ g, int nimg, int flat)
{
    int k, img;
    term_index(D.nnci, nimg, flat, k, img);
    const double* x = xyz + (size_t)img * 3 * D.n;
    double* gi = g + (size_t)img * 3 * D.n;
    double e = 0.0;
    if (k < D.nnci) {
        const int i1 = D.nci[3 * k], i2 = D.nci[3 * k + 1], nk = D.nci[3 * k + 2] - 1;
        double a[3], b[3], vab[3];
        ld3(x, i1, a);
        ld3(x, i2, b);
        for (int c = 0; c < 3; c++) vab[c] = a[c] - b[c];
        if (D.periodic) box_image(D, vab);
        const double r2 = dot3(vab, vab), r = sqrt(r2);
        double dr = 0.0, d1, d2;
        if (!(D.periodic && r > D.vdw_cut)) {
            const int lo = min(i1, i2), hi = max(i1, i2);
            e += vdw_pair(D, D.type[i1], D.type[i2], D.c6[(size_t)lo * D.n + hi], r2, r, D.eps2[nk], d1);
            dr += d1;
        }
        e += coul_pair(D, D.q[i1] * D.q[i2], r2, r, D.eps1[nk], d2);
        dr += d2;
        double v[3] = {vab[0] * dr, vab[1] * dr, vab[2] * dr};
        add3(gi, i1, v);
        v[0] = -v[0];
        v[1] = -v[1];
        v[2] = -v[2];
        add3(gi, i2, v);
    }
    term_energy(e, &V[img], flat);
}

// Per image: xs[img][c][n], an FP64 SoA copy of the positions (exact pair evaluation, coalesced
// gathers), and xf[img][n] = {x, y, z, molnum} packed as float4 (one 16-byte load per atom in the
// conservative FP32 pre-filters of qm_inter_kernel and qm_hb_search_kernel).
__global__ void qm_soa_kernel(const double* __restrict__ xyz, const int* __restrict__ molnum, int n, size_t natot,
                              double* __restrict__ xs, float4* __restrict__ xf)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // (image, atom)
    if (t >= natot) return;
    const size_t img = t / (size_t)n;
    const int a = (int)(t - img * n);
    const double* p = xyz + 3 * t;
    const double x = p[0], y = p[1], z = p[2];
    double* o = xs + img * 3 * n + a;
    o[0] = x;
    o[n] = y;
    o[2 * (size_t)n] = z;
    xf[t] = make_float4((float)x, (float)y, (float)z, __int_as_float(molnum[a]));
}

// FP32 minimum image for the pre-filters (any distance; the exact box_image.f90 loop runs on the survivors)
__device__ __forceinline__ float image_f(float v, float L, float iL) { return fmaf(-L, rintf(v * iL), v); }
// relative slack of the FP32 pre-filters: covers the rounding of coordinates up to ~1e4 bohr
#define QM_PREFILTER_SLACK 1.002f
#define QM_HB_SEG 512   // atoms j per warp in qm_hb_search_kernel

// Inter-molecular part of ff_nonb (ff_nonb.f90:198-332 dispersion/repulsion, :421-512 Coulomb), every
// pair i < j of different molecules once.  One WARP per (image, atom i): the lanes sweep j = i+1..n-1
// in chunks of 32 with a cheap test (other molecule, minimum image, r^2 inside the larger cut-off)
// and compact the survivors into a per-warp queue in shared memory; whenever 32 pairs are queued
// they are evaluated with all lanes busy, so that the expensive part (exp, erfc, reciprocals) never
// runs on a mostly-idle warp -- with 10 A cut-offs in a 31 A box only ~14 % of the tests survive and
// a lane-per-pair sweep without compaction executes the expensive path on almost every iteration.
// g_i is reduced over the warp in registers, g_j goes out as red.global.add.f64.
struct InterTables {
    double r094[QM_MAXTYPE][QM_MAXTYPE], sr42[QM_MAXTYPE][QM_MAXTYPE], r0ab[QM_MAXTYPE][QM_MAXTYPE],
        zab[QM_MAXTYPE][QM_MAXTYPE];
};

// One inter-molecular pair i < j evaluated exactly as ff_nonb.f90:198-332 (dispersion / repulsion) and :421-512
// (Coulomb) do: vab = x(i) - x(j), the box_image.f90 loop, the exact r > cut tests.  ep = energy of the pair,
// (fx,fy,fz) = its gradient contribution on atom i (atom j gets the negative).  Shared by the O(N^2) sweep
// (qm_inter_kernel) and the cell sweep (qm_inter_cell_kernel) so that a pair gives the same bits in both.
template <bool PER>
__device__ __forceinline__ void qm_pair_exact(const QmdffDev& D, const InterTables& tb, const double* __restrict__ Xx,
                                              const double* __restrict__ Xy, const double* __restrict__ Xz, int i, int j,
                                              double& ep, double& fx, double& fy, double& fz)
{
    constexpr bool per = PER;
    auto image = [&](double& v, double L) {
        const double L2 = 0.5 * L;
        while (fabs(v) > L2) v -= (v >= 0.0) ? L : -L;   // box_image.f90
    };
    double vab[3] = {Xx[i] - Xx[j], Xy[i] - Xy[j], Xz[i] - Xz[j]};   // x(i1) - x(i2), i1 < i2
    if (per) {
        image(vab[0], D.box[0]);
        image(vab[1], D.box[1]);
        image(vab[2], D.box[2]);
    }
    const double r2 = vab[0] * vab[0] + vab[1] * vab[1] + vab[2] * vab[2];
    double r, oner;
    sqrt_rsqrt(r2, r, oner);   // branch-free (crcl_common.cuh, fm::): a pair is ~1/3 fewer instructions without the library's slow paths
    const double oner2 = oner * oner;
    double dr = 0.0;
    ep = 0.0;
    if (!(per && r > D.vdw_cut)) {
        const int ti = D.type[i], tj = D.type[j];
        const double c6 = D.ncls ? __ldg(&D.c6c[D.cls[i] * D.ncls + D.cls[j]]) : __ldg(&D.c6[(size_t)i * D.n + j]);
        const double R0 = tb.r094[ti][tj];
        const double r4 = r2 * r2, r6 = r4 * r2, R02 = R0 * R0, r06 = R02 * R02 * R02;
        const double t6 = r6 + r06, t8 = r6 * r2 + r06 * R02;
        const double it6 = CRCL_RCP(t6), it8 = CRCL_RCP(t8);
        const double c6t6 = c6 * it6, t27 = tb.sr42[ti][tj] * (c6 * it8);
        ep -= c6t6 + t27;
        dr += c6t6 * 6.0 * r4 * it6 + 8.0 * t27 * r6 * it8;
        if (r < 25.0) {
            const double alpha = tb.r0ab[ti][tj];
            const double tt = tb.zab[ti][tj] * CRCL_EXP(-alpha * r);
            ep += tt * oner;
            dr -= tt * (alpha * r + 1.0) * oner * oner2;
        }
    }
    if (!(r > D.coul_cut)) {
        const double qq = D.q[i] * D.q[j];
        double e0;
        if (D.zahn) {
            e0 = qq * (erfc(D.zahn_a * r) * oner - D.zahn_par * (r - D.coul_cut));
        } else {
            double sw = 1.0;
            if (per && r > D.cut_low) {
                const double xv = (r - D.cut_low) / (D.coul_cut - D.cut_low);
                sw = exp(1.0) * exp(1.0 / (xv - 1.0));
            }
            e0 = qq * oner * sw;
        }
        ep += e0;
        dr -= e0 * oner2;   // the reference uses e0/r^2 for every Coulomb form (ff_nonb.f90:470)
    }
    fx = vab[0] * dr;
    fy = vab[1] * dr;
    fz = vab[2] * dr;
}

// ---- cell sweep of the inter-molecular part (periodic boxes of at least 2M+1 cells per dimension) -----------------
// The reference tests all n(n-1)/2 pairs per image (ff_nonb.f90:198,421: double loops over the atoms); with 10 A
// cut-offs in a 31 A box 86 % of those tests fail.  Here the atoms of an image are binned into cells of edge
// >= rc (1 + slack) / M (M = 2 or 3) and sorted by cell; a CTA owns one home cell and tests its atoms against the
// half shell of (2M+1)^3 neighbour cells (every unordered cell pair once; own cell: sorted position j > i) in FP32 on
// the wrapped coordinates, with the periodic shift a constant of the neighbour run instead of a rounding per pair.
// Survivors go to a CTA-wide queue in shared memory and are evaluated 128 at a time by qm_pair_exact on the
// original FP64 coordinates, ordered (min, max) by atom number as the reference's loops are: every pair inside the
// cut-offs gives the same bits as in the O(N^2) sweep, only the order of the FP64 accumulation differs (which the
// red.global.add of the O(N^2) sweep leaves open as well).
struct CellGrid {
    int nc[3], ncell, m;      // cells per dimension, total, neighbour range
    float rc2f;               // (larger cut-off)^2 with slack
};
#define QM_CELL_MAXCELL 8192   // shared-memory counters of qm_cellsort_kernel
#define QM_CELL_IB 32          // home-cell atoms per batch
#define QM_CELL_MAXRUN 80      // 1 + 3 x (M(2M+1) + M+1) segment slots, M <= 3

// one CTA per image: count, scan, scatter.  sf[pos] = {wrapped x, y, z, atom number}, smol[pos] = molnum,
// cstart[img][ncell+1]
__global__ void __launch_bounds__(1024) qm_cellsort_kernel(const QmdffDev D, const CellGrid G, const double* __restrict__ xs,
                                                           float4* __restrict__ sf, int* __restrict__ smol,
                                                           int* __restrict__ cstart)
{
    __shared__ int cnt[QM_CELL_MAXCELL + 1];
    __shared__ int part[1024];
    const int img = blockIdx.x, n = D.n, tid = threadIdx.x;
    const double* X = xs + (size_t)img * 3 * n;
    for (int c = tid; c <= G.ncell; c += 1024) cnt[c] = 0;
    __syncthreads();
    auto wrapped = [&](int a, float w[3], int& cell) {
        int ci[3];
        for (int d = 0; d < 3; d++) {
            const double L = D.box[d], v = X[(size_t)d * n + a];
            double u = v - L * floor(v / L);              // [0, L] (== L only by rounding)
            int k = (int)(u * ((double)G.nc[d] / L));
            k = k < 0 ? 0 : (k >= G.nc[d] ? G.nc[d] - 1 : k);
            ci[d] = k;
            w[d] = (float)u;
        }
        cell = (ci[2] * G.nc[1] + ci[1]) * G.nc[0] + ci[0];
    };
    for (int a = tid; a < n; a += 1024) {
        float w[3];
        int cell;
        wrapped(a, w, cell);
        atomicAdd(&cnt[cell], 1);
    }
    __syncthreads();
    // exclusive scan over the cells: thread t owns a contiguous chunk
    const int chunk = (G.ncell + 1023) / 1024, c0 = tid * chunk, c1 = min(G.ncell, c0 + chunk);
    int s = 0;
    for (int c = c0; c < c1; c++) s += cnt[c];
    part[tid] = s;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int t = 0; t < 1024; t++) {
            const int v = part[t];
            part[t] = run;
            run += v;
        }
    }
    __syncthreads();
    int run = part[tid];
    int* cs = cstart + (size_t)img * (G.ncell + 1);
    for (int c = c0; c < c1; c++) {
        const int v = cnt[c];
        cs[c] = run;
        cnt[c] = run;      // becomes the cell's cursor
        run += v;
    }
    if (tid == 0) cs[G.ncell] = n;
    __syncthreads();
    for (int a = tid; a < n; a += 1024) {
        float w[3];
        int cell;
        wrapped(a, w, cell);
        const int pos = atomicAdd(&cnt[cell], 1);
        sf[(size_t)img * n + pos] = make_float4(w[0], w[1], w[2], __int_as_float(a));
        smol[(size_t)img * n + pos] = D.molnum[a];
    }
}

__global__ void __launch_bounds__(128) qm_inter_cell_kernel(const QmdffDev D, const CellGrid G, const double* __restrict__ xs,
                                                            const float4* __restrict__ sf, const int* __restrict__ smol,
                                                            const int* __restrict__ cstart, double* __restrict__ V,
                                                            double* __restrict__ g)
{
    __shared__ InterTables tb;
    __shared__ int run_beg[QM_CELL_MAXRUN], run_pre[QM_CELL_MAXRUN + 1], run_len[QM_CELL_MAXRUN];
    __shared__ float run_sh[QM_CELL_MAXRUN][3];
    __shared__ float4 si[QM_CELL_IB];
    __shared__ int simol[QM_CELL_IB];
    __shared__ int2 queue[4][64];                          // per warp: < 32 left over + up to 32 new per test round
    const int img = blockIdx.y, n = D.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int t = tid; t < QM_MAXTYPE * QM_MAXTYPE; t += 128) {
        (&tb.r094[0][0])[t] = (&D.r094[0][0])[t];
        (&tb.sr42[0][0])[t] = (&D.sr42[0][0])[t];
        (&tb.r0ab[0][0])[t] = (&D.r0ab[0][0])[t];
        (&tb.zab[0][0])[t] = (&D.zab[0][0])[t];
    }
    const int* cs = cstart + (size_t)img * (G.ncell + 1);
    const int home = blockIdx.x, hb = cs[home], he = cs[home + 1];
    if (hb == he) return;                                  // empty home cell (uniform for the CTA)
    const int M = G.m, ncx = G.nc[0], ncy = G.nc[1], ncz = G.nc[2];
    const int cx = home % ncx, cy = (home / ncx) % ncy, cz = home / (ncx * ncy);
    // runs of the half shell: rows (oz, oy) with oz > 0, or oz == 0 and oy >= 0; the row (0, 0) only holds the
    // cells ox = 1..M (the home cell itself is run 0).  A row's cells are contiguous in the sorted order except
    // where the periodic wrap splits it: up to three segments, each with its own shift.  Thread t < nrows fills
    // the three segment slots of row t (empty segments keep length 0), thread 0 the prefix.
    const int nrows = M * (2 * M + 1) + (M + 1);
    if (tid < nrows) {
        int oz, oy;                                        // rows in the order (0; 0..M), (1..M; -M..M)
        if (tid <= M) {
            oz = 0;
            oy = tid;
        } else {
            const int t = tid - (M + 1);
            oz = 1 + t / (2 * M + 1);
            oy = t % (2 * M + 1) - M;
        }
        const float Lx = (float)D.box[0], Ly = (float)D.box[1], Lz = (float)D.box[2];
        int z2 = cz + oz, y2 = cy + oy;
        float shz = 0.f, shy = 0.f;
        if (z2 >= ncz) { z2 -= ncz; shz = Lz; }
        if (y2 < 0) { y2 += ncy; shy = -Ly; } else if (y2 >= ncy) { y2 -= ncy; shy = Ly; }
        const int row = (z2 * ncy + y2) * ncx;
        const int xa = cx + ((oz == 0 && oy == 0) ? 1 : -M), xb = cx + M;   // inclusive cell range, may leave [0, ncx)
        // segments: [xa, -1] -> +ncx, shift -Lx ; [max(xa,0), min(xb,ncx-1)] ; [ncx, xb] -> -ncx, shift +Lx
        const int lo[3] = {xa, xa > 0 ? xa : 0, ncx}, hi[3] = {-1 < xb ? -1 : xb, xb < ncx - 1 ? xb : ncx - 1, xb};
        const int off[3] = {ncx, 0, -ncx};
        const float shx[3] = {-Lx, 0.f, Lx};
        for (int sgm = 0; sgm < 3; sgm++) {
            const int k = 1 + 3 * tid + sgm;
            int b = 0, len = 0;
            if (lo[sgm] <= hi[sgm]) {
                b = cs[row + lo[sgm] + off[sgm]];
                len = cs[row + hi[sgm] + off[sgm] + 1] - b;
            }
            run_beg[k] = b;
            run_len[k] = len;
            run_sh[k][0] = shx[sgm];
            run_sh[k][1] = shy;
            run_sh[k][2] = shz;
        }
    }
    if (tid == 0) {
        run_beg[0] = hb;                                  // run 0: the home cell itself (pairs with jpos > ipos)
        run_len[0] = he - hb;
        run_sh[0][0] = run_sh[0][1] = run_sh[0][2] = 0.f;
    }
    __syncthreads();
    const int nrun = 1 + 3 * nrows;
    if (tid == 0) {
        int acc = 0;
        for (int k = 0; k < nrun; k++) {
            run_pre[k] = acc;
            acc += run_len[k];
        }
        run_pre[nrun] = acc;
    }
    __syncthreads();
    const int total = run_pre[nrun];
    const double* X = xs + (size_t)img * 3 * n;
    const double *Xx = X, *Xy = X + n, *Xz = X + 2 * (size_t)n;
    const float4* F = sf + (size_t)img * n;
    const int* Mo = smol + (size_t)img * n;
    double e = 0.0;
    int2* qu = queue[warp];
    int qn = 0;
    const unsigned lt = (1u << lane) - 1u;
    auto flush_one = [&](int2 pr) {
        const int i = pr.x < pr.y ? pr.x : pr.y, j = pr.x < pr.y ? pr.y : pr.x;
        double ep, fx, fy, fz;
        qm_pair_exact<true>(D, tb, Xx, Xy, Xz, i, j, ep, fx, fy, fz);
        e += ep;
        double* gi = g + (size_t)img * 3 * n + 3 * (size_t)i;
        double* gj = g + (size_t)img * 3 * n + 3 * (size_t)j;
        atomicAdd(gi, fx);
        atomicAdd(gi + 1, fy);
        atomicAdd(gi + 2, fz);
        atomicAdd(gj, -fx);
        atomicAdd(gj + 1, -fy);
        atomicAdd(gj + 2, -fz);
    };
    for (int ib = hb; ib < he; ib += QM_CELL_IB) {
        const int nb = min(QM_CELL_IB, he - ib);
        __syncthreads();                                   // the previous batch is no longer read
        if (tid < nb) {
            si[tid] = F[ib + tid];
            simol[tid] = Mo[ib + tid];
        }
        __syncthreads();
        // every warp sweeps its own 32 candidates of each 128-chunk against the batch and keeps its own queue: no
        // CTA-wide barrier inside the sweep, warps drift apart and hide each other's latencies
        int r = 0;
        for (int f0 = warp * 32; f0 < total; f0 += 128) {  // uniform per warp
            const int f = f0 + lane;
            const bool valid = f < total;
            float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
            int molj = -1, jpos = 0;
            bool own = false;
            if (valid) {
                while (f >= run_pre[r + 1]) r++;
                jpos = run_beg[r] + (f - run_pre[r]);
                c = F[jpos];
                molj = Mo[jpos];
                c.x += run_sh[r][0];
                c.y += run_sh[r][1];
                c.z += run_sh[r][2];
                own = (r == 0);
            }
            const int ja = __float_as_int(c.w);
            for (int ii = 0; ii < nb; ii++) {
                const float4 a = si[ii];
                const float dx = a.x - c.x, dy = a.y - c.y, dz = a.z - c.z;
                bool ok = valid && (dx * dx + dy * dy + dz * dz <= G.rc2f) && (simol[ii] != molj);
                if (own) ok = ok && (jpos > ib + ii);
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                if (m) {                                   // uniform
                    if (ok) qu[qn + __popc(m & lt)] = make_int2(__float_as_int(a.w), ja);
                    qn += __popc(m);
                    __syncwarp();
                    if (qn >= 32) {
                        qn -= 32;
                        const int2 pr = qu[qn + lane];
                        __syncwarp();
                        flush_one(pr);
                    }
                }
            }
        }
    }
    __syncwarp();
    if (lane < qn) flush_one(qu[lane]);
    block_sum_to(e, &V[img]);
}

template <bool PER>
__global__ void __launch_bounds__(128) qm_inter_kernel(const QmdffDev D, const double* __restrict__ xs,
                                                       const float4* __restrict__ xf, double* __restrict__ V,
                                                       double* __restrict__ g)
{
    __shared__ InterTables tb;
    __shared__ int queue[4][64];
    const int img = blockIdx.y, n = D.n, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t = threadIdx.x; t < QM_MAXTYPE * QM_MAXTYPE; t += 128) {
        (&tb.r094[0][0])[t] = (&D.r094[0][0])[t];
        (&tb.sr42[0][0])[t] = (&D.sr42[0][0])[t];
        (&tb.r0ab[0][0])[t] = (&D.r0ab[0][0])[t];
        (&tb.zab[0][0])[t] = (&D.zab[0][0])[t];
    }
    __syncthreads();
    const double* X = xs + (size_t)img * 3 * n;
    const double *Xx = X, *Xy = X + n, *Xz = X + 2 * (size_t)n;
    const int i = blockIdx.x * 4 + warp;
    double e = 0.0;
    if (i < n - 1) {
        const int mi = D.molnum[i];
        constexpr bool per = PER;
        const double Lx = D.box[0], Ly = D.box[1], Lz = D.box[2];
        const double rc = fmax(D.vdw_cut, D.coul_cut);
        double gx = 0.0, gy = 0.0, gz = 0.0;
        int* qu = queue[warp];
        int qn = 0;
        auto pair = [&](int j) {
            double ep, fx, fy, fz;
            qm_pair_exact<PER>(D, tb, Xx, Xy, Xz, i, j, ep, fx, fy, fz);
            e += ep;
            gx += fx;
            gy += fy;
            gz += fz;
            double* gj = g + (size_t)img * 3 * n + 3 * (size_t)j;
            atomicAdd(gj, -fx);
            atomicAdd(gj + 1, -fy);
            atomicAdd(gj + 2, -fz);
        };
        // cheap test in FP32 with slack (the exact r > cut tests follow in pair()); the next chunk
        // is loaded before the current one is tested / flushed
        const float4* F = xf + (size_t)img * n;
        const float4 fi = F[i];
        const float Lxf = (float)Lx, Lyf = (float)Ly, Lzf = (float)Lz;
        const float iLx = 1.0f / Lxf, iLy = 1.0f / Lyf, iLz = 1.0f / Lzf;
        const float rc2f = (float)(rc * rc) * QM_PREFILTER_SLACK;
        const float mif = __int_as_float(mi);
        // other molecule and inside the larger cut-off (with slack); branch-free
        auto near = [&](const float4& c) {
            bool ok = __float_as_int(c.w) != mi;
            if (per) {
                const float dx = image_f(fi.x - c.x, Lxf, iLx), dy = image_f(fi.y - c.y, Lyf, iLy),
                            dz = image_f(fi.z - c.z, Lzf, iLz);
                ok = ok & (dx * dx + dy * dy + dz * dz <= rc2f);
            }
            return ok;
        };
        const unsigned lt = (1u << lane) - 1u;
        auto push = [&](bool ok, int j) {
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (ok) qu[qn + __popc(m & lt)] = j;
            qn += __popc(m);
            __syncwarp();
            if (qn >= 32) {
                qn -= 32;
                pair(qu[qn + lane]);
                __syncwarp();
            }
        };
        const float4 self = make_float4(0.f, 0.f, 0.f, mif);   // fails the molecule test
        // first chunk (j > i) and last chunk (j < n) are guarded, the chunks in between are not;
        // the main loop walks a pointer, two chunks per trip (the loads of the next two are in flight
        // while the current two are tested / flushed)
        const int jstart = (i + 1) & ~31, jfull = n & ~31;
        {
            const int j = jstart + lane;
            const float4 c = (j > i && j < n) ? F[j] : self;
            push(near(c), j);
        }
        int j = jstart + 32 + lane;                      // this lane's atom in the current chunk
        const float4* pf = F + j;
        const int jend2 = jfull - 64;                     // last j0 for which two full chunks remain
        if (jstart + 32 <= jend2) {
            float4 c0 = pf[0], c1 = pf[32];
            for (;;) {
                const bool more = (j - lane) + 64 <= jend2;
                float4 n0 = self, n1 = self;
                if (more) {
                    n0 = pf[64];
                    n1 = pf[96];
                }
                push(near(c0), j);
                push(near(c1), j + 32);
                j += 64;
                pf += 64;
                if (!more) break;
                c0 = n0;
                c1 = n1;
            }
        }
        for (int j0 = j - lane; j0 < n; j0 += 32) {      // at most one full chunk and the partial one
            const int jj = j0 + lane;
            const float4 c = (jj < n) ? F[jj] : self;
            push(near(c), jj);
        }
        if (lane < qn) pair(qu[lane]);
        for (int o = 16; o > 0; o >>= 1) {
            gx += __shfl_xor_sync(0xffffffffu, gx, o);
            gy += __shfl_xor_sync(0xffffffffu, gy, o);
            gz += __shfl_xor_sync(0xffffffffu, gz, o);
        }
        if (lane == 0) {
            double* gi = g + (size_t)img * 3 * n + 3 * (size_t)i;
            atomicAdd(gi, gx);
            atomicAdd(gi + 1, gy);
            atomicAdd(gi + 2, gz);
        }
    }
    block_sum_to(e, &V[img]);
}

// ---- H/X-bond terms -------------------------------------------------------------------------------
// eabhag.f90:30-210 (analytic); only drah is imaged, as in the reference (F9)
__device__ __forceinline__ double eabhag_dev(const QmdffDev& D, const double* x, double* gi, int A, int B, int H,
                                             double ca, double cb)
{
    double xa[3], xb[3], xh[3], drah[3], drbh[3], drab[3];
    ld3(x, A, xa);
    ld3(x, B, xb);
    ld3(x, H, xh);
    for (int c = 0; c < 3; c++) {
        drah[c] = xa[c] - xh[c];
        drbh[c] = xb[c] - xh[c];
        drab[c] = xa[c] - xb[c];
    }
    if (D.periodic) box_image(D, drah);
    const double rab2 = dot3(drab, drab), rab = sqrt(rab2), rah2 = dot3(drah, drah), rah = sqrt(rah2),
                 rbh2 = dot3(drbh, drbh), rbh = sqrt(rbh2);
    const double ratio = pow(rab / 8.0, 12.0);
    const double rdampl = 1.0 / (1.0 + ratio) / rab2 / rab;
    const bool far = rah2 > rbh2;
    const double aprod = far ? 1.0 / rbh / rab : 1.0 / rah / rab;
    const double cosabh = far ? -dot3(drbh, drab) * aprod : dot3(drah, drab) * aprod;
    double aterm = 0.5 * (cosabh + 1.0);
    const double a2 = aterm * aterm, a5 = a2 * a2 * aterm;   // aterm**(alp3-1), alp3 = 6
    aterm = aterm * a5;
    const double apref = 3.0 * a5;
    const double rah4 = rah2 * rah2, rbh4 = rbh2 * rbh2, denom = 1.0 / (rah4 + rbh4);
    const double da = (ca * rah4 + cb * rbh4) * denom;
    const double eabh = -da * rdampl * aterm;
    if (eabh > -1.e-8) return 0.0;
    const double gia = -(4.0 * (ca - cb) * rah2 * rbh4 * denom * denom) * rdampl * aterm;
    const double gib = -(4.0 * (cb - ca) * rbh2 * rah4 * denom * denom) * rdampl * aterm;
    const double gid = rdampl * rdampl * rab * (3.0 + 15.0 * ratio) * da * aterm;
    const double gip = -da * rdampl * apref;
    double ga[3], gb[3], gh[3];
    for (int c = 0; c < 3; c++) {
        ga[c] = gia * drah[c];
        gb[c] = gib * drbh[c];
        gh[c] = -ga[c] - gb[c];
        const double dg = gid * drab[c];
        ga[c] += dg;
        gb[c] -= dg;
        if (far) {
            const double d1 = gip * (-aprod * drbh[c] - cosabh * drab[c] / rab2);
            const double d2 = gip * (aprod * drab[c] + cosabh * drbh[c] / rbh2);
            ga[c] += d1;
            gh[c] += d2;
            gb[c] -= d1 + d2;
        } else {
            const double d1 = gip * (-aprod * drah[c] + cosabh * drab[c] / rab2);
            const double d2 = gip * (-aprod * drab[c] + cosabh * drah[c] / rah2);
            gb[c] += d1;
            gh[c] += d2;
            ga[c] -= d1 + d2;
        }
    }
    add3(gi, A, ga);
    add3(gi, B, gb);
    add3(gi, H, gh);
    return eabh;
}
// eabx.f90:30-110 on explicit positions
__device__ __forceinline__ double eabx_dev(const QmdffDev& D, const double xa[3], const double xb[3],
                                           const double xh[3], double ca)
{
    double r[3];
    for (int c = 0; c < 3; c++) r[c] = xa[c] - xb[c];
    if (D.periodic) box_image(D, r);
    const double rab2 = dot3(r, r);
    const double dampl = 1.0 / (1.0 + pow(rab2 / 120.0, 6.0));
    for (int c = 0; c < 3; c++) r[c] = xa[c] - xh[c];
    if (D.periodic) box_image(D, r);
    const double d2ik = dot3(r, r);
    for (int c = 0; c < 3; c++) r[c] = xh[c] - xb[c];
    if (D.periodic) box_image(D, r);
    const double d2jk = dot3(r, r);
    const double term = (d2ik > d2jk) ? 0.5 * (rab2 + d2jk - d2ik) / sqrt(rab2 * d2jk)
                                      : 0.5 * (rab2 + d2ik - d2jk) / sqrt(rab2 * d2ik);
    return -ca * dampl * pow(0.5 * (term + 1.0), 6.0) / d2jk;
}
// eabxag.f90:30-150: central differences with step 1e-6; the reference assigns the scalar
// (er-el)*dum to all three components on every pass, so the z-derivative is added to x, y and z (F9)
__device__ __forceinline__ double eabxag_dev(const QmdffDev& D, const double* x, double* gi, int A, int B, int H,
                                             double ca)
{
    const double step = 1.e-6, dum = 1.0 / (2.0 * step);
    double p[3][3];
    ld3(x, A, p[0]);
    ld3(x, B, p[1]);
    ld3(x, H, p[2]);
    const double e0 = eabx_dev(D, p[0], p[1], p[2], ca);
    const int who[3] = {A, B, H};
    for (int w = 0; w < 3; w++) {
        double gl = 0.0;
        for (int j = 0; j < 3; j++) {
            p[w][j] = p[w][j] + step;
            const double er = eabx_dev(D, p[0], p[1], p[2], ca);
            p[w][j] = p[w][j] - step * 2.0;
            const double el = eabx_dev(D, p[0], p[1], p[2], ca);
            p[w][j] = p[w][j] + step;
            gl = (er - el) * dum;
        }
        const double v[3] = {gl, gl, gl};
        add3(gi, who[w], v);
    }
    return e0;
}
__device__ __forceinline__ double dist_dev(const QmdffDev& D, const double* x, int a, int b, bool image)
{
    double r[3] = {x[3 * a] - x[3 * b], x[3 * a + 1] - x[3 * b + 1], x[3 * a + 2] - x[3 * b + 2]};
    if (image && D.periodic) box_image(D, r);
    return sqrt(dot3(r, r));
}

__global__ void __launch_bounds__(128) qm_hb_list_kernel(const QmdffDev D, const double* __restrict__ xyz,
                                                         double* __restrict__ V, double* __restrict__ g, int nimg,
                                                         int flat)
{
    int k, img;
    term_index(D.nhb, nimg, flat, k, img);
    const double* x = xyz + (size_t)img * 3 * D.n;
    double* gi = g + (size_t)img * 3 * D.n;
    double e = 0.0;
    if (k < D.nhb) {
        const int A = D.hb[3 * k], B = D.hb[3 * k + 1], H = D.hb[3 * k + 2];
        if (!(dist_dev(D, x, A, B, true) > 15.0)) {
            if (D.isH[k])
                e = eabhag_dev(D, x, gi, A, B, H, D.vhb[2 * k], D.vhb[2 * k + 1]);
            else
                e = eabxag_dev(D, x, gi, A, B, H, D.vhb[2 * k]);
        }
    }
    term_energy(e, &V[img], flat);
}

// ff_hb.f90:90-274: every (donor bond, atom j of another molecule) pair.  One WARP per (image,
// donor): the lanes sweep all atoms j with a cheap test (other molecule, acceptor class, A..j
// inside 15 bohr in FP32 with slack) and compact the survivors into a per-warp queue; queued
// pairs go through the complete reference logic (exact distance tests included) 32 at a time.
__global__ void __launch_bounds__(128) qm_hb_search_kernel(const QmdffDev D, const double* __restrict__ xyz,
                                                           const float4* __restrict__ xf, double* __restrict__ V,
                                                           double* __restrict__ g)
{
    constexpr double c12 = (double)1.2f, c13 = (double)1.3f, b0 = (double)0.52917726f;
    __shared__ int queue[4][64];
    const int img = blockIdx.y, n = D.n, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d = blockIdx.x * 4 + warp;
    const double* x = xyz + (size_t)img * 3 * n;
    double* gi = g + (size_t)img * 3 * n;
    double e = 0.0;
    if (d < D.ndonor) {
        const int H = D.donor[3 * d], A = D.donor[3 * d + 1], kind = D.donor[3 * d + 2];
        const int mH = D.molnum[H];
        const double ri = dist_dev(D, x, A, H, true), dthr = D.dthr[d], dcoef = D.dcoef[d], dscal = D.dscal[d];
        const double radHal = D.rad[D.type[H]];
        auto pair = [&](int j) {
            // the reference's tests, in its order (ff_hb.f90:150-274)
            const double rj = dist_dev(D, x, j, H, true);
            if (kind == 1) {
                const double dum2 = c12 * (D.rad[D.type[j]] + radHal) / b0;
                if (ri < dthr || rj < dum2)
                    if (!(dist_dev(D, x, A, j, false) > 15.0)) e += eabxag_dev(D, x, gi, A, j, H, dcoef);
            } else {
                const double dum2 = c13 * (D.rad[D.type[j]] + D.radH) / b0;
                if (ri < dthr || rj < dum2)
                    if (!(dist_dev(D, x, A, j, true) > 15.0)) e += eabhag_dev(D, x, gi, j, A, H, D.acc_c1[j], dcoef);
            }
        };
        const float4* F = xf + (size_t)img * n;
        const float4 fa = F[A];
        // kind 1 tests the A..j distance WITHOUT the minimum image (ff_hb.f90:176-180)
        const bool img_aj = D.periodic && kind == 2;
        const float Lxf = (float)D.box[0], Lyf = (float)D.box[1], Lzf = (float)D.box[2];
        const float iLx = 1.0f / Lxf, iLy = 1.0f / Lyf, iLz = 1.0f / Lzf;
        const float r2max = 225.0f * QM_PREFILTER_SLACK;
        int* qu = queue[warp];
        int qn = 0;
        // blockIdx.z: segment of QM_HB_SEG atoms j (more warps in flight when there are few donors)
        const int jbeg = blockIdx.z * QM_HB_SEG, jend = min(n, jbeg + QM_HB_SEG);
        const float4 far4 = make_float4(0.f, 0.f, 0.f, __int_as_float(mH));   // fails the molecule test
        for (int j0 = jbeg; j0 < jend; j0 += 128) {
            // four chunks of 32 atoms: all loads first, then the tests (the sweep is latency-bound)
            float4 c[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int j = j0 + 32 * u + lane;
                c[u] = (j < jend) ? F[j] : far4;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int j = j0 + 32 * u + lane;
                float dx = fa.x - c[u].x, dy = fa.y - c[u].y, dz = fa.z - c[u].z;
                if (img_aj) {
                    dx = image_f(dx, Lxf, iLx);
                    dy = image_f(dy, Lyf, iLy);
                    dz = image_f(dz, Lzf, iLz);
                }
                bool ok = __float_as_int(c[u].w) != mH && dx * dx + dy * dy + dz * dz <= r2max;
                if (ok) ok = (kind == 1) ? D.acc_no[j] != 0 : dscal * D.acc_s[j] > 1e-6;
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                if (ok) qu[qn + __popc(m & ((1u << lane) - 1u))] = j;
                qn += __popc(m);
                __syncwarp();
                if (qn >= 32) {
                    qn -= 32;
                    pair(qu[qn + lane]);
                    __syncwarp();
                }
            }
        }
        if (lane < qn) pair(qu[lane]);
    }
    block_sum_to(e, &V[img]);
}

__global__ void qm_init_kernel(double* V, int nimg, double e_zero)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nimg) V[i] = e_zero;
}

cudaError_t qmdff_egrad(QmdffDev* D, const double* d_xyz, int nimg, double* d_V, double* d_g, cudaStream_t s,
                        long long* launches)
{
    if (nimg <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(d_g, 0, (size_t)nimg * 3 * D->n * sizeof(double), s);
    if (e != cudaSuccess) return e;
    const bool inter = !(D->nnci <= 1 && (D->is_two || D->nmols == 0)) && D->nmols > 1;
    const bool hbs = D->use_hb && D->nmols > 1 && D->ndonor > 0;
    if (inter || hbs) {
        const size_t natot = (size_t)nimg * D->n, total = 3 * natot;
        if (total > D->xs_cap) {
            // stream-ordered with the kernels that still read the old buffer
            if (D->xs && (e = cudaFreeAsync(D->xs, s)) != cudaSuccess) return e;
            if (D->xf && (e = cudaFreeAsync(D->xf, s)) != cudaSuccess) return e;
            D->xs = nullptr;
            D->xf = nullptr;
            D->xs_cap = 0;
            if ((e = cudaMallocAsync(&D->xs, total * sizeof(double), s)) != cudaSuccess) return e;
            if ((e = cudaMallocAsync(&D->xf, natot * sizeof(float4), s)) != cudaSuccess) return e;
            D->xs_cap = total;
        }
        qm_soa_kernel<<<(unsigned)((natot + 255) / 256), 256, 0, s>>>(d_xyz, D->molnum, D->n, natot, D->xs, D->xf);
        if (launches) ++*launches;
    }
    // cell sweep of the inter-molecular part: periodic boxes that hold at least 2M+1 cells of edge rc(1+slack)/M per
    // dimension (M = 2; 3 on request); everything else keeps the O(N^2) sweep.  CRCL_QM_CELLS=0
    // switches it off (A/B measurements, tests of the fallback)
    CellGrid G{};
    bool use_cells = false;
    if (inter && D->periodic && D->cells_enabled) {
        const double rc = std::max(D->vdw_cut, D->coul_cut), rcs = rc * std::sqrt((double)QM_PREFILTER_SLACK) * 1.0005;
        // M = 2 by default: measured on B200 (profiles/r2d_bench_qmdff_cells.json) the finer M = 3 grid loses more to its
        // short inner loops (4 atoms per home cell) than its tighter shell saves; CRCL_QM_CELL_M=3 selects it
        const char* em = getenv("CRCL_QM_CELL_M");
        const int m_hi = (em && em[0] == '3') ? 3 : 2;
        for (int m = m_hi; m >= 2 && !use_cells; m--) {
            bool ok = true;
            long long nct = 1;
            for (int d = 0; d < 3; d++) {
                G.nc[d] = (int)std::floor(D->box[d] * m / rcs);
                ok = ok && G.nc[d] >= 2 * m + 1;
                nct *= G.nc[d];
            }
            if (ok && nct <= QM_CELL_MAXCELL) {
                G.m = m;
                G.ncell = (int)nct;
                G.rc2f = (float)(rc * rc) * QM_PREFILTER_SLACK;
                use_cells = true;
            }
        }
    }
    if (use_cells) {
        const int ni_max = std::min(65535, nimg);
        const size_t need_a = (size_t)ni_max * D->n, need_c = (size_t)ni_max * (G.ncell + 1);
        if (need_a > D->cell_cap_a || need_c > D->cell_cap_c) {
            if (D->sf && (e = cudaFreeAsync(D->sf, s)) != cudaSuccess) return e;
            if (D->smol && (e = cudaFreeAsync(D->smol, s)) != cudaSuccess) return e;
            if (D->cstart && (e = cudaFreeAsync(D->cstart, s)) != cudaSuccess) return e;
            D->sf = nullptr;
            D->smol = nullptr;
            D->cstart = nullptr;
            D->cell_cap_a = D->cell_cap_c = 0;
            if ((e = cudaMallocAsync(&D->sf, need_a * sizeof(float4), s)) != cudaSuccess) return e;
            if ((e = cudaMallocAsync(&D->smol, need_a * sizeof(int), s)) != cudaSuccess) return e;
            if ((e = cudaMallocAsync(&D->cstart, need_c * sizeof(int), s)) != cudaSuccess) return e;
            D->cell_cap_a = need_a;
            D->cell_cap_c = need_c;
        }
    }
    qm_init_kernel<<<(nimg + 127) / 128, 128, 0, s>>>(d_V, nimg, D->e_zero);
    int nl = 1;
    // blockIdx.y carries the image: at most 65535 images per launch
    for (int i0 = 0; i0 < nimg; i0 += 65535) {
        const int ni = std::min(65535, nimg - i0);
        const double* x = d_xyz + (size_t)i0 * 3 * D->n;
        double* g = d_g + (size_t)i0 * 3 * D->n;
        double* V = d_V + i0;
        const int nterm = D->nbond + D->nangl + D->ntors;
        // CRCL_QM_TERM_MAJOR=0 keeps the image-major forms (A/B)
        static const bool tm_on = [] { const char* e = getenv("CRCL_QM_TERM_MAJOR"); return !(e && e[0] == '0'); }();
        const bool term_major = tm_on && ni >= 512 && D->n <= 128;
        if (nterm > 0) {
            if (term_major)   // many images of a small molecule: flat (term, image) index, see term_index
                qm_bonded_kernel<<<(unsigned)(((size_t)nterm * ni + 127) / 128), 128, 0, s>>>(*D, x, V, g, ni, 2);
            else if (nterm < 96)   // short lists: flat (image, term) index
                qm_bonded_kernel<<<(unsigned)(((size_t)nterm * ni + 127) / 128), 128, 0, s>>>(*D, x, V, g, ni, 1);
            else
                qm_bonded_kernel<<<dim3((nterm + 127) / 128, ni), 128, 0, s>>>(*D, x, V, g, ni, 0);
            nl++;
        }
        // early returns: ff_nonb.f90:74 (nnci <= 1 and nmols == 0), ff_nonb_two.f90:48 (nnci_two <= 1)
        if (!(D->nnci <= 1 && (D->is_two || D->nmols == 0))) {
            if (D->nnci > 0) {
                if (term_major)
                    qm_nci_kernel<<<(unsigned)(((size_t)D->nnci * ni + 127) / 128), 128, 0, s>>>(*D, x, V, g, ni, 2);
                else if (D->nnci < 96)
                    qm_nci_kernel<<<(unsigned)(((size_t)D->nnci * ni + 127) / 128), 128, 0, s>>>(*D, x, V, g, ni, 1);
                else
                    qm_nci_kernel<<<dim3((D->nnci + 127) / 128, ni), 128, 0, s>>>(*D, x, V, g, ni, 0);
                nl++;
            }
            if (D->nmols > 1) {
                if (use_cells) {
                    qm_cellsort_kernel<<<ni, 1024, 0, s>>>(*D, G, D->xs + (size_t)i0 * 3 * D->n, D->sf, D->smol, D->cstart);
                    qm_inter_cell_kernel<<<dim3(G.ncell, ni), 128, 0, s>>>(*D, G, D->xs + (size_t)i0 * 3 * D->n, D->sf,
                                                                           D->smol, D->cstart, V, g);
                    nl++;
                } else if (D->periodic)
                    qm_inter_kernel<true><<<dim3((D->n + 3) / 4, ni), 128, 0, s>>>(*D, D->xs + (size_t)i0 * 3 * D->n,
                                                                                   D->xf + (size_t)i0 * D->n, V, g);
                else
                    qm_inter_kernel<false><<<dim3((D->n + 3) / 4, ni), 128, 0, s>>>(*D, D->xs + (size_t)i0 * 3 * D->n,
                                                                                    D->xf + (size_t)i0 * D->n, V, g);
                nl++;
            }
        }
        if (D->use_hb && !(D->nhb < 1 && D->nmols == 0)) {   // ff_hb.f90:49 early return
            if (D->nhb > 0) {
                if (D->nhb < 96)
                    qm_hb_list_kernel<<<(unsigned)(((size_t)D->nhb * ni + 127) / 128), 128, 0, s>>>(*D, x, V, g, ni, 1);
                else
                    qm_hb_list_kernel<<<dim3((D->nhb + 127) / 128, ni), 128, 0, s>>>(*D, x, V, g, ni, 0);
                nl++;
            }
            if (D->nmols > 1 && D->ndonor > 0) {
                qm_hb_search_kernel<<<dim3((D->ndonor + 3) / 4, ni, (D->n + QM_HB_SEG - 1) / QM_HB_SEG), 128, 0, s>>>(
                    *D, x, D->xf + (size_t)i0 * D->n, V, g);
                nl++;
            }
        }
    }
    if (launches) *launches += nl;
    return cudaGetLastError();
}

template <class T>
static T* up(const T* h, size_t n, bool& ok)
{
    T* d = nullptr;
    if (!ok) return nullptr;
    if (cudaMalloc(&d, (n ? n : 1) * sizeof(T)) != cudaSuccess) {
        ok = false;
        return nullptr;
    }
    if (n && cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) ok = false;
    return d;
}

int qmdff_upload(const crcl_qmdff_tables* T, QmdffDev** out, const char** err, bool is_two)
{
    *out = nullptr;
    if (T->nhb > 0 && (!T->scalehb || !T->hb || !T->vhb)) {
        *err = "nhb > 0 needs hb, vhb and the scalehb/scalexb/q_glob tables";
        return CRCL_EINVAL;
    }
    const int n = T->n;
    QmdffDev* D = new QmdffDev();
    memset(D, 0, sizeof(*D));
    {
        const char* ev = getenv("CRCL_QM_CELLS");
        D->cells_enabled = !(ev && ev[0] == '0');
    }
    D->n = n;
    D->nbond = T->nbond;
    D->nangl = T->nangl;
    D->ntors = T->ntors;
    D->nnci = T->nnci;
    D->ldvt = T->ldvt;
    D->nmols = T->nmols;
    // element types
    int zmap[95];
    for (int z = 0; z < 95; z++) zmap[z] = -1;
    int ztab[QM_MAXTYPE];
    std::vector<int> type(n), mol(n);
    for (int a = 0; a < n; a++) {
        const int z = T->at[a];
        if (z < 1 || z > 94) {
            *err = "atomic number out of range 1..94";
            delete D;
            return CRCL_EINVAL;
        }
        if (zmap[z] < 0) {
            if (D->ntype == QM_MAXTYPE) {
                *err = "more than 12 distinct elements in one QMDFF";
                delete D;
                return CRCL_ENOSUP;
            }
            ztab[D->ntype] = z;
            zmap[z] = D->ntype++;
        }
        type[a] = zmap[z];
        mol[a] = T->molnum ? T->molnum[a] : 1;
    }
    for (int a = 0; a < D->ntype; a++) {
        D->rad[a] = T->rad[ztab[a] - 1];
        for (int b = 0; b < D->ntype; b++) {
            const size_t ix = (size_t)(ztab[a] - 1) + 94 * (size_t)(ztab[b] - 1);   // Fortran (94,94)
            D->r0ab[a][b] = T->r0ab[ix];
            D->zab[a][b] = T->zab[ix];
            D->r094[a][b] = T->r094[ix];
            D->sr42[a][b] = T->sr42[ix];
        }
    }
    for (int k = 0; k < 6; k++) {
        D->eps1[k] = T->eps1[k];
        D->eps2[k] = T->eps2[k];
    }
    D->periodic = T->periodic;
    D->zahn = T->zahn;
    for (int d = 0; d < 3; d++) D->box[d] = T->box[d];
    D->coul_cut = T->coul_cut;
    D->vdw_cut = T->vdw_cut;
    D->cut_low = T->cut_low;
    D->zahn_a = T->zahn_a;
    D->zahn_par = T->zahn_par;
    D->e_zero = T->e_zero;
    if (is_two) {
        // ff_eg_two / ff_nonb_two / ff_hb_two: no PBC, no inter-molecular loops, Coulomb without cut-off
        D->is_two = 1;
        D->periodic = 0;
        D->zahn = 0;
        D->nmols = 1;
        D->coul_cut = 1.0e300;
        for (int a = 0; a < n; a++) mol[a] = 1;
    }
    // lists -> 0-based
    auto idx_ok = [&](int v) { return v >= 1 && v <= n; };
    std::vector<int> bond(2 * (size_t)T->nbond), angl(3 * (size_t)T->nangl), tors(6 * (size_t)T->ntors),
        nci(3 * (size_t)T->nnci);
    bool good = true;
    for (size_t k = 0; k < bond.size(); k++) {
        good &= idx_ok(T->bond[k]);
        bond[k] = T->bond[k] - 1;
    }
    for (size_t k = 0; k < angl.size(); k++) {
        good &= idx_ok(T->angl[k]);
        angl[k] = T->angl[k] - 1;
    }
    for (int m = 0; m < T->ntors; m++) {
        for (int c = 0; c < 4; c++) {
            good &= idx_ok(T->tors[6 * m + c]);
            tors[6 * m + c] = T->tors[6 * m + c] - 1;
        }
        tors[6 * m + 4] = T->tors[6 * m + 4];
        tors[6 * m + 5] = T->tors[6 * m + 5];
        good &= T->tors[6 * m + 4] >= 0 && 2 + 3 * T->tors[6 * m + 4] <= T->ldvt;
    }
    for (int k = 0; k < T->nnci; k++) {
        good &= idx_ok(T->nci[3 * k]) && idx_ok(T->nci[3 * k + 1]) && T->nci[3 * k + 2] >= 1 && T->nci[3 * k + 2] <= 6;
        nci[3 * k] = T->nci[3 * k] - 1;
        nci[3 * k + 1] = T->nci[3 * k + 1] - 1;
        nci[3 * k + 2] = T->nci[3 * k + 2];
    }
    if (!good) {
        *err = "QMDFF list entry out of range";
        delete D;
        return CRCL_EINVAL;
    }
    // c6: the reference reads c66ab(i2,i1) with i1 < i2, i.e. element (max,min) of the Fortran array
    std::vector<double> c6((size_t)n * n);
    for (int a = 0; a < n; a++)
        for (int b = 0; b < n; b++) {
            const int lo = a < b ? a : b, hi = a < b ? b : a;
            c6[(size_t)a * n + b] = T->c6xy[(size_t)hi + (size_t)n * lo];
        }
    // c6 classes: atoms whose rows of the symmetric table coincide share a class (a box of identical solvent molecules
    // has one class per atom of the molecule).  c6(i,j) = c6c[cls(i)][cls(j)] is verified for EVERY pair before it is
    // used, so the pair kernels read the same doubles from a table that lives in L1 / L2 instead of gathering from the
    // n x n array (73 MB at 3030 atoms); more than QM_MAXCLS classes or any mismatch keeps the full table.
    std::vector<int> cls(n, 0);
    std::vector<double> c6c;
    int ncls = 0;
    {
        std::vector<int> rep;
        std::vector<unsigned long long> rep_hash;
        for (int a = 0; a < n && ncls >= 0; a++) {
            unsigned long long hsh = 1469598103934665603ull;   // FNV-1a over the row's bytes: rows are only compared on a hit
            const unsigned char* rb = reinterpret_cast<const unsigned char*>(&c6[(size_t)a * n]);
            for (size_t t = 0; t < sizeof(double) * (size_t)n; t++) hsh = (hsh ^ rb[t]) * 1099511628211ull;
            int found = -1;
            for (int k = 0; k < (int)rep.size(); k++)
                if (rep_hash[k] == hsh && memcmp(&c6[(size_t)a * n], &c6[(size_t)rep[k] * n], sizeof(double) * n) == 0) {
                    found = k;
                    break;
                }
            if (found < 0) {
                if ((int)rep.size() == QM_MAXCLS) {
                    ncls = -1;
                    break;
                }
                found = (int)rep.size();
                rep.push_back(a);
                rep_hash.push_back(hsh);
            }
            cls[a] = found;
        }
        if (ncls >= 0) {
            ncls = (int)rep.size();
            c6c.assign((size_t)ncls * ncls, 0.0);
            for (int k = 0; k < ncls; k++)
                for (int l = 0; l < ncls; l++) {
                    // a representative pair of (k, l): first atom of class l other than rep[k] if possible
                    int b = rep[l];
                    c6c[(size_t)k * ncls + l] = c6[(size_t)rep[k] * n + b];
                }
            // rows equal does not yet make c6(a,b) a function of the classes of BOTH atoms: check every entry
            bool same = true;
            for (int a = 0; a < n && same; a++)
                for (int b = 0; b < n; b++)
                    if (memcmp(&c6[(size_t)a * n + b], &c6c[(size_t)cls[a] * ncls + cls[b]], sizeof(double)) != 0) {
                        same = false;
                        break;
                    }
            if (!same) ncls = -1;
        }
        if (ncls < 0) ncls = 0;
    }
    D->ncls = ncls;
    // ff_hb tables: hb list, donor bonds (static topology) and per-atom acceptor data
    D->use_hb = T->scalehb ? 1 : 0;
    D->nhb = D->use_hb ? T->nhb : 0;
    std::vector<int> hb(3 * (size_t)D->nhb), isH(D->nhb), donor, accno(n, 0);
    std::vector<double> dthr, dcoef, dscal, accc1(n, 0.0), accs(n, 0.0);
    if (D->use_hb) {
        if (!T->scalexb || !T->q_glob) {
            *err = "scalehb given without scalexb / q_glob";
            delete D;
            return CRCL_EINVAL;
        }
        auto hbpara = [](double a, double b, double q) { return std::exp(-a * q) / (std::exp(-a * q) + b); };
        const double c12 = (double)1.2f, c13 = (double)1.3f, b0 = (double)0.52917726f;
        for (int k = 0; k < D->nhb; k++) {
            for (int c = 0; c < 3; c++) {
                if (!idx_ok(T->hb[3 * k + c])) good = false;
                hb[3 * k + c] = T->hb[3 * k + c] - 1;
            }
            if (good) isH[k] = T->at[hb[3 * k + 2]] == 1;
        }
        if (!good) {
            *err = "hb list entry out of range";
            delete D;
            return CRCL_EINVAL;
        }
        D->radH = T->rad[0];
        for (int a = 0; a < n; a++) {
            const int z = T->at[a];
            accs[a] = T->scalehb[z - 1];
            accc1[a] = hbpara(10.0, 5.0, T->q_glob[a]) * T->scalehb[z - 1];
            accno[a] = (z == 7 || z == 8);
        }
        auto isX = [](int z) { return z == 17 || z == 35 || z == 53 || z == 85; };
        auto isAcc = [](int z) { return z == 7 || z == 8 || z == 9 || z == 16 || z == 17; };
        for (int m = 0; m < T->nbond && T->nmols > 1; m++) {
            const int i1 = T->bond[2 * m] - 1, i2 = T->bond[2 * m + 1] - 1;
            const int z1 = T->at[i1], z2 = T->at[i2];
            int xh = -1, xa = -1;
            if (isX(z1)) {
                if (z2 != 1) xh = i1, xa = i2;
            } else if (isX(z2)) {
                if (z1 != 1) xh = i2, xa = i1;
            }
            if (xh >= 0) {
                donor.insert(donor.end(), {xh, xa, 1});
                dthr.push_back(c12 * (T->rad[T->at[xa] - 1] + T->rad[T->at[xh] - 1]) / b0);
                dcoef.push_back(T->scalexb[T->at[xh] - 1] * hbpara(-6.5, 1.0, T->q_glob[xh]));
                dscal.push_back(0.0);
            }
            int hh = -1, ha = -1;
            if (z1 == 1 && isAcc(z2)) hh = i1, ha = i2;
            if (z2 == 1 && isAcc(z1)) hh = i2, ha = i1;
            if (hh >= 0) {
                donor.insert(donor.end(), {hh, ha, 2});
                dthr.push_back(c13 * (T->rad[T->at[ha] - 1] + T->rad[0]) / b0);
                dcoef.push_back(hbpara(10.0, 5.0, T->q_glob[ha]) * T->scalehb[T->at[ha] - 1]);
                dscal.push_back(T->scalehb[T->at[ha] - 1]);
            }
        }
        D->ndonor = (int)dthr.size();
    }
    bool ok = true;
    D->hb = up(hb.data(), hb.size(), ok);
    D->vhb = up(D->nhb ? T->vhb : nullptr, 2 * (size_t)D->nhb, ok);
    D->isH = up(isH.data(), isH.size(), ok);
    D->donor = up(donor.data(), donor.size(), ok);
    D->dthr = up(dthr.data(), dthr.size(), ok);
    D->dcoef = up(dcoef.data(), dcoef.size(), ok);
    D->dscal = up(dscal.data(), dscal.size(), ok);
    D->acc_c1 = up(accc1.data(), accc1.size(), ok);
    D->acc_s = up(accs.data(), accs.size(), ok);
    D->acc_no = up(accno.data(), accno.size(), ok);
    D->type = up(type.data(), n, ok);
    D->molnum = up(mol.data(), n, ok);
    D->q = up(T->q, n, ok);
    D->bond = up(bond.data(), bond.size(), ok);
    D->vbond = up(T->vbond, 3 * (size_t)T->nbond, ok);
    D->angl = up(angl.data(), angl.size(), ok);
    D->vangl = up(T->vangl, 2 * (size_t)T->nangl, ok);
    D->tors = up(tors.data(), tors.size(), ok);
    D->vtors = up(T->vtors, (size_t)T->ldvt * T->ntors, ok);
    D->nci = up(nci.data(), nci.size(), ok);
    D->c6 = up(c6.data(), c6.size(), ok);
    D->cls = up(cls.data(), cls.size(), ok);
    D->c6c = up(c6c.data(), c6c.size(), ok);
    if (!ok) {
        qmdff_free(D);
        *err = "device allocation / upload of the QMDFF tables failed";
        return CRCL_ENOMEM;
    }
    *out = D;
    return CRCL_OK;
}

void qmdff_free(QmdffDev* D)
{
    if (!D) return;
    cudaFree(D->type);
    cudaFree(D->molnum);
    cudaFree(D->q);
    cudaFree(D->bond);
    cudaFree(D->vbond);
    cudaFree(D->angl);
    cudaFree(D->vangl);
    cudaFree(D->tors);
    cudaFree(D->vtors);
    cudaFree(D->nci);
    cudaFree(D->c6);
    cudaFree(D->cls);
    cudaFree(D->c6c);
    cudaFree(D->hb);
    cudaFree(D->vhb);
    cudaFree(D->isH);
    cudaFree(D->donor);
    cudaFree(D->dthr);
    cudaFree(D->dcoef);
    cudaFree(D->dscal);
    cudaFree(D->acc_c1);
    cudaFree(D->acc_s);
    cudaFree(D->acc_no);
    cudaFree(D->xs);
    cudaFree(D->xf);
    cudaFree(D->sf);
    cudaFree(D->smol);
    cudaFree(D->cstart);
    delete D;
}

}  // namespace crcl
