/*
 * qmdff.c -- CPU oracle: QMDFF bonded and non-bonded energy and gradient of one structure.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Parity UNPINNED by the reference.
 *
 * Literal restatement of
 *   ff_eg.f90:40-629       orc_ff_eg     bonds, angles, torsions, inversions (no virial, no
 *                                        eval_cutoff / debug branches)
 *   abdamp.f90:35-50       abdamp        (rcut = 3.0*3.5710642*(...) : REAL*4 literals, F3)
 *   valijkl.f90:34-102, dphidr.f90:34-126, omega.f90:35-67, domegadr.f90:35-114,
 *   crossprod.f90, crprod.f90, impsc.f90, vlen.f90, vecnorm.f90, box_image.f90
 *   ff_nonb.f90:33-512     orc_ff_nonb   nci pair list + inter-molecular O(N^2) loops,
 *                                        dispersion/repulsion and Coulomb (Zahn / cut-off /
 *                                        exp_switch.f90).  The Ewald/SPME branch is dead code
 *                                        (ewald=.false. at :337).
 *   ff_hb.f90:35-280       orc_ff_hb     hb list + on-the-fly donor/acceptor search (nmols > 1);
 *   eabhag.f90 (analytic H-bond, images only drah: F9), eabxag.f90 + eabx.f90 (X-bond, numeric
 *   gradient whose last component is broadcast to all three: F9), hbpara.f90
 *   gradient.f90:341-362   orc_qmdff_egrad  e = ff_eg + ff_nonb + ff_hb + E_zero1
 * Tables (c6xy, r0ab, zab, r094_mod, sr42, rad, eps1, eps2) are inputs, as they are for the
 * reference routines (built once by prepare.f90 / setnonb.f90); (94,94) and (n,n)
 * arrays are passed in Fortran order.  Atom indices are 1-based in the lists, as in the .qmdff file.
 */
#include <stdlib.h>
#include <string.h>
#include "oracle_real.h"
#include "qmdff.h"

#define PI_Q 3.1415926535897932384626433832795029
#define PI2_Q 6.28318530717958623199592693708837
#define SPI_Q 1.77245385090551599275151910313925

/* census hooks: the counting build (count_qmdff.cpp) drops the operations of a pair / donor-acceptor test that fails
 * its cut-off, so that the census holds algorithmic work only ("only pairs inside the cutoff") */
#ifndef ORC_MARK
#define ORC_MARK()
#define ORC_REJECT()
#endif

static void box_image(const orc_qmdff *f, double v[3])
{
    int d;
    for (d = 0; d < 3; d++) {
        const double L = f->box[d], L2 = f->box[d] * 0.5;
        while (fabs(v[d]) > L2) v[d] = v[d] - (v[d] >= 0 ? L : -L);
    }
}
static void cross(const double a[3], const double b[3], double c[3])
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
static double vlen(const double a[3]) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
static double vecnorm(double r[3], int inorm)
{
    double sp = r[0] * r[0] + r[1] * r[1] + r[2] * r[2], rn = sqrt(sp);
    if (inorm > 0 && fabs(rn) > 1.e-14) {
        double o = 1.0 / rn;
        r[0] = o * r[0];
        r[1] = o * r[1];
        r[2] = o * r[2];
    }
    return rn;
}
/* abdamp.f90 */
static void abdamp(const orc_qmdff *f, int ati, int atj, double r2, double *damp, double *ddamp)
{
    const double rs = f->rad[ati - 1] + f->rad[atj - 1];
    const double rcut = (double)(3.0f * 3.5710642f) * (rs * rs);
    const double rr = (r2 / rcut) * (r2 / rcut);
    *damp = 1.0 / (1.0 + rr);
    *ddamp = -2.0 * 2 * rr / (r2 * ((1.0 + rr) * (1.0 + rr)));
}
#define X(i, c) xyz[3 * ((i) - 1) + (c)]
#define G(i, c) g[3 * ((i) - 1) + (c)]

/* valijkl.f90 */
static double valijkl(const orc_qmdff *f, const double *xyz, int i, int j, int k, int l)
{
    double ra[3], rb[3], rc[3], na[3], nb[3], snanb = 0.0;
    int c;
    for (c = 0; c < 3; c++) {
        ra[c] = X(j, c) - X(i, c);
        rb[c] = X(k, c) - X(j, c);
        rc[c] = X(l, c) - X(k, c);
    }
    if (f->periodic) {
        box_image(f, ra);
        box_image(f, rb);
        box_image(f, rc);
    }
    cross(ra, rb, na);
    cross(rb, rc, nb);
    vecnorm(na, 1);
    vecnorm(nb, 1);
    for (c = 0; c < 3; c++) snanb += na[c] * nb[c];
    if (fabs(fabs(snanb) - 1.0) < 1.0e-14) snanb = (snanb >= 0) ? 1.0 : -1.0;
    return acos(snanb);
}
/* dphidr.f90 */
static void dphidr(const orc_qmdff *f, const double *xyz, int i, int j, int k, int l, double phi, double di[3],
                   double dj[3], double dk[3], double dl[3])
{
    double ra[3], rb[3], rc[3], rapb[3], rbpc[3], na[3], nb[3], rab[3], rba[3], rac[3], rbb[3], rbc[3], raa[3],
        rapba[3], rapbb[3], rbpca[3], rbpcb[3];
    double cosphi = cos(phi), sinphi = sin(phi), nan_, nbn, nenner, on;
    int c;
    for (c = 0; c < 3; c++) {
        ra[c] = X(j, c) - X(i, c);
        rb[c] = X(k, c) - X(j, c);
        rc[c] = X(l, c) - X(k, c);
    }
    if (f->periodic) {
        box_image(f, ra);
        box_image(f, rb);
        box_image(f, rc);
    }
    for (c = 0; c < 3; c++) {
        rapb[c] = ra[c] + rb[c];
        rbpc[c] = rb[c] + rc[c];
    }
    cross(ra, rb, na);
    cross(rb, rc, nb);
    nan_ = vecnorm(na, 0);
    nbn = vecnorm(nb, 0);
    nenner = nan_ * nbn * sinphi;
    if (fabs(nenner) < 1.e-14) {
        for (c = 0; c < 3; c++) di[c] = dj[c] = dk[c] = dl[c] = 0.0;
        return;
    }
    on = 1.0 / nenner;
    cross(na, rb, rab);
    cross(nb, ra, rba);
    cross(na, rc, rac);
    cross(nb, rb, rbb);
    cross(nb, rc, rbc);
    cross(na, ra, raa);
    cross(rapb, na, rapba);
    cross(rapb, nb, rapbb);
    cross(rbpc, na, rbpca);
    cross(rbpc, nb, rbpcb);
    for (c = 0; c < 3; c++) {
        di[c] = on * (cosphi * nbn / nan_ * rab[c] - rbb[c]);
        dj[c] = on * (cosphi * (nbn / nan_ * rapba[c] + nan_ / nbn * rbc[c]) - (rac[c] + rapbb[c]));
        dk[c] = on * (cosphi * (nbn / nan_ * raa[c] + nan_ / nbn * rbpcb[c]) - (rba[c] + rbpca[c]));
        dl[c] = on * (cosphi * nan_ / nbn * rbb[c] - rab[c]);
    }
}
/* omega.f90 */
static double omega(const orc_qmdff *f, const double *xyz, int i, int j, int k, int l)
{
    double rd[3], re[3], rn[3], rv[3];
    int c;
    for (c = 0; c < 3; c++) {
        re[c] = X(i, c) - X(j, c);
        rd[c] = X(k, c) - X(j, c);
        rv[c] = X(l, c) - X(i, c);
    }
    if (f->periodic) {
        box_image(f, re);
        box_image(f, rd);
        box_image(f, rv);
    }
    cross(re, rd, rn);
    vecnorm(rn, 1);
    vecnorm(rv, 1);
    return asin(rn[0] * rv[0] + rn[1] * rv[1] + rn[2] * rv[2]);
}
/* domegadr.f90 */
static void domegadr(const orc_qmdff *f, const double *xyz, int i, int j, int k, int l, double om, double di[3],
                     double dj[3], double dk[3], double dl[3])
{
    double rn[3], rv[3], rd[3], re[3], rdme[3], rve[3], rne[3], rdv[3], rdn[3], rvdme[3], rndme[3];
    double sinomega = sin(om), rvn, rnn, nenner, on;
    int c;
    for (c = 0; c < 3; c++) {
        rv[c] = X(l, c) - X(i, c);
        rd[c] = X(k, c) - X(j, c);
        re[c] = X(i, c) - X(j, c);
    }
    if (f->periodic) {
        box_image(f, rv);
        box_image(f, rd);
        box_image(f, re);
    }
    for (c = 0; c < 3; c++) rdme[c] = rd[c] - re[c];
    cross(re, rd, rn);
    rvn = vecnorm(rv, 0);
    rnn = vecnorm(rn, 0);
    cross(rv, re, rve);
    cross(rn, re, rne);
    cross(rd, rv, rdv);
    cross(rd, rn, rdn);
    cross(rv, rdme, rvdme);
    cross(rn, rdme, rndme);
    nenner = rnn * rvn * cos(om);
    if (fabs(nenner) > 1.e-14) {
        on = 1.0 / nenner;
        for (c = 0; c < 3; c++) {
            di[c] = on * (rdv[c] - rn[c] - sinomega * (rvn / rnn * rdn[c] - rnn / rvn * rv[c]));
            dj[c] = on * (rvdme[c] - sinomega * rvn / rnn * rndme[c]);
            dk[c] = on * (rve[c] - sinomega * rvn / rnn * rne[c]);
            dl[c] = on * (rn[c] - sinomega * rnn / rvn * rv[c]);
        }
    } else {
        for (c = 0; c < 3; c++) di[c] = dj[c] = dk[c] = dl[c] = 0.0;
    }
}

/* ff_eg.f90:40-629 */
void orc_ff_eg(const orc_qmdff *f, const double *xyz, double *e_out, double *g)
{
    const int n = f->n;
    double e = 0.0;
    int m, c;
    memset(g, 0, sizeof(double) * 3 * n);
    for (m = 0; m < f->nbond; m++) {
        const int i = f->bond[2 * m], j = f->bond[2 * m + 1];
        double rb[3], r2, r, rij, kij, aai, aai2, fac;
        for (c = 0; c < 3; c++) rb[c] = X(i, c) - X(j, c);
        if (f->periodic) box_image(f, rb);
        r2 = rb[0] * rb[0] + rb[1] * rb[1] + rb[2] * rb[2];
        r = sqrt(r2);
        rij = f->vbond[3 * m];
        kij = f->vbond[3 * m + 1];
        aai = f->vbond[3 * m + 2];
        aai2 = aai / 2;
        e = e + kij * (1. + pow(rij / r, aai) - 2. * pow(rij / r, aai2));
        fac = aai * kij * (-pow(rij / r, aai) + pow(rij / r, aai2)) / r2;
        for (c = 0; c < 3; c++) {
            G(i, c) += fac * rb[c];
            G(j, c) -= fac * rb[c];
        }
    }
    for (m = 0; m < f->nangl; m++) {
        const int j = f->angl[3 * m], i = f->angl[3 * m + 1], k = f->angl[3 * m + 2];
        const double c0 = f->vangl[2 * m], kijk = f->vangl[2 * m + 1];
        double vab[3], vcb[3], vp[3], deda[3], dedc[3], dedb[3], term1[3], term2[3];
        double rab2, rcb2, rp, cosa, theta, dampij, damp2ij, dampjk, damp2jk, damp, ea, deddt, rmul1, rmul2, al, bl;
        for (c = 0; c < 3; c++) {
            vab[c] = X(i, c) - X(j, c);
            vcb[c] = X(k, c) - X(j, c);
        }
        if (f->periodic) {
            box_image(f, vab);
            box_image(f, vcb);
        }
        rab2 = vab[0] * vab[0] + vab[1] * vab[1] + vab[2] * vab[2];
        rcb2 = vcb[0] * vcb[0] + vcb[1] * vcb[1] + vcb[2] * vcb[2];
        cross(vcb, vab, vp);
        rp = vlen(vp) + 1.e-14;
        al = vlen(vab);
        bl = vlen(vcb);
        cosa = (al > 0.0 && bl > 0.0) ? (vab[0] * vcb[0] + vab[1] * vcb[1] + vab[2] * vcb[2]) / (al * bl) : 0.0;
        cosa = (cosa > 1.0) ? 1.0 : (cosa < -1.0 ? -1.0 : cosa);
        theta = acos(cosa);
        abdamp(f, f->at[i - 1], f->at[j - 1], rab2, &dampij, &damp2ij);
        abdamp(f, f->at[k - 1], f->at[j - 1], rcb2, &dampjk, &damp2jk);
        damp = dampij * dampjk;
        if (PI_Q - c0 < 1.e-6) {
            const double dt = theta - c0;
            ea = kijk * dt * dt;
            deddt = 2.0 * kijk * dt;
        } else {
            ea = kijk * (cosa - cos(c0)) * (cosa - cos(c0));
            deddt = 2. * kijk * sin(theta) * (cos(c0) - cosa);
        }
        e = e + ea * damp;
        cross(vab, vp, deda);
        rmul1 = -deddt / (rab2 * rp);
        cross(vcb, vp, dedc);
        rmul2 = deddt / (rcb2 * rp);
        for (c = 0; c < 3; c++) {
            deda[c] *= rmul1;
            dedc[c] *= rmul2;
            dedb[c] = deda[c] + dedc[c];
            term1[c] = ea * damp2ij * dampjk * vab[c];
            term2[c] = ea * damp2jk * dampij * vcb[c];
            G(i, c) += deda[c] * damp + term1[c];
            G(j, c) += -dedb[c] * damp - term1[c] - term2[c];
            G(k, c) += dedc[c] * damp + term2[c];
        }
    }
    for (m = 0; m < f->ntors; m++) {
        const int *t = f->tors + 6 * m;
        const double *vt = f->vtors + (size_t)f->ldvt * m;
        const int i = t[0], j = t[1], k = t[2], l = t[3], nt = t[4];
        const double phi0 = vt[0];
        double vab[3], vcb[3], vdc[3], dda[3], ddb[3], ddc[3], ddd[3], term1[3], term2[3], term3[3];
        double rij, rjk, rkl, dampij, damp2ij, dampjk, damp2jk, dampkl, damp2kl, damp, phi, et, dij;
        if (t[5] != 2) {
            int it, mm = 2;
            for (c = 0; c < 3; c++) {
                vab[c] = X(i, c) - X(j, c);
                vcb[c] = X(j, c) - X(k, c);
                vdc[c] = X(k, c) - X(l, c);
            }
            if (f->periodic) {
                box_image(f, vab);
                box_image(f, vcb);
                box_image(f, vdc);
            }
            rij = vab[0] * vab[0] + vab[1] * vab[1] + vab[2] * vab[2];
            rjk = vcb[0] * vcb[0] + vcb[1] * vcb[1] + vcb[2] * vcb[2];
            rkl = vdc[0] * vdc[0] + vdc[1] * vdc[1] + vdc[2] * vdc[2];
            abdamp(f, f->at[i - 1], f->at[j - 1], rij, &dampij, &damp2ij);
            abdamp(f, f->at[k - 1], f->at[j - 1], rjk, &dampjk, &damp2jk);
            abdamp(f, f->at[k - 1], f->at[l - 1], rkl, &dampkl, &damp2kl);
            damp = dampjk * dampij * dampkl;
            phi = valijkl(f, xyz, i, j, k, l);
            dphidr(f, xyz, i, j, k, l, phi, dda, ddb, ddc, ddd);
            et = 0;
            dij = 0;
            for (it = 1; it <= nt; it++) {
                const double rn = vt[mm], ph = vt[mm + 1], vv = vt[mm + 2];
                const double dphi1 = phi - phi0, dphi2 = phi + phi0 - PI2_Q;
                const double c1 = rn * dphi1 + ph, c2 = rn * dphi2 + ph;
                const double x1cos = cos(c1), x2cos = cos(c2), x1sin = sin(c1), x2sin = sin(c2);
                const double phipi = phi - PI_Q, ef = erf(phipi);
                const double e1 = vv * (1. + x1cos), e2 = vv * (1. + x2cos);
                const double expo = exp(-phipi * phipi) / SPI_Q;
                et = et + 0.5 * (1. - ef) * e1 + (0.5 + 0.5 * ef) * e2;
                dij = dij - expo * e1 - 0.5 * (1. - ef) * vv * x1sin * rn + expo * e2 -
                      (0.5 + 0.5 * ef) * vv * x2sin * rn;
                mm += 3;
            }
            et = et * vt[1];
            dij = dij * vt[1] * damp;
            for (c = 0; c < 3; c++) {
                term1[c] = et * damp2ij * dampjk * dampkl * vab[c];
                term2[c] = et * damp2jk * dampij * dampkl * vcb[c];
                term3[c] = et * damp2kl * dampij * dampjk * vdc[c];
                G(i, c) += dij * dda[c] + term1[c];
                G(j, c) += dij * ddb[c] - term1[c] + term2[c];
                G(k, c) += dij * ddc[c] + term3[c] - term2[c];
                G(l, c) += dij * ddd[c] - term3[c];
            }
            e = e + et * damp;
        } else {
            double rjl, dampjl, damp2jl, rn;
            for (c = 0; c < 3; c++) {
                vab[c] = X(j, c) - X(i, c);
                vcb[c] = X(j, c) - X(k, c);
                vdc[c] = X(j, c) - X(l, c);
            }
            if (f->periodic) {
                box_image(f, vab);
                box_image(f, vcb);
                box_image(f, vdc);
            }
            rij = vab[0] * vab[0] + vab[1] * vab[1] + vab[2] * vab[2];
            rjk = vcb[0] * vcb[0] + vcb[1] * vcb[1] + vcb[2] * vcb[2];
            rjl = vdc[0] * vdc[0] + vdc[1] * vdc[1] + vdc[2] * vdc[2];
            abdamp(f, f->at[i - 1], f->at[j - 1], rij, &dampij, &damp2ij);
            abdamp(f, f->at[k - 1], f->at[j - 1], rjk, &dampjk, &damp2jk);
            abdamp(f, f->at[j - 1], f->at[l - 1], rjl, &dampjl, &damp2jl);
            damp = dampjk * dampij * dampjl;
            phi = omega(f, xyz, i, j, k, l);
            domegadr(f, xyz, i, j, k, l, phi, dda, ddb, ddc, ddd);
            rn = vt[2];
            if (rn > 1.e-6) {
                const double c1 = (phi - phi0) + PI_Q;
                et = (1. + cos(c1)) * vt[1];
                dij = -sin(c1) * vt[1] * damp;
            } else {
                et = vt[1] * (cos(phi) - cos(phi0)) * (cos(phi) - cos(phi0));
                dij = 2. * vt[1] * sin(phi) * (cos(phi0) - cos(phi)) * damp;
            }
            for (c = 0; c < 3; c++) {
                term1[c] = et * damp2ij * dampjk * dampjl * vab[c];
                term2[c] = et * damp2jk * dampij * dampjl * vcb[c];
                term3[c] = et * damp2jl * dampij * dampjk * vdc[c];
                G(i, c) += dij * dda[c] - term1[c];
                G(j, c) += dij * ddb[c] + term1[c] + term2[c] + term3[c];
                G(k, c) += dij * ddc[c] - term2[c];
                G(l, c) += dij * ddd[c] - term3[c];
            }
            e = e + et * damp;
        }
    }
    *e_out = e;
}

#define T94(tab, a, b) ((tab)[((a) - 1) + 94 * ((b) - 1)])

static void nonb_vdw_pair(const orc_qmdff *f, const double *xyz, int i1, int i2, double eps, double *e, double *g)
{
    double vab[3], r2, r, oner, R0, c6, r4, r6, r06, t6, t8, c6t6, c6t8, t27, e0, drij, ga[3];
    const int iz1 = f->at[i1 - 1], iz2 = f->at[i2 - 1];
    int c;
    ORC_MARK();
    for (c = 0; c < 3; c++) vab[c] = X(i1, c) - X(i2, c);
    if (f->periodic) box_image(f, vab);
    r2 = vab[0] * vab[0] + vab[1] * vab[1] + vab[2] * vab[2];
    r = sqrt(r2);
    if (f->periodic && r > f->vdw_cut) {
        ORC_REJECT();
        return;
    }
    oner = 1.0 / r;
    R0 = T94(f->r094, iz1, iz2);
    c6 = f->c6xy[(i2 - 1) + (size_t)f->n * (i1 - 1)];
    r4 = r2 * r2;
    r6 = r4 * r2;
    r06 = R0 * R0 * R0 * R0 * R0 * R0;
    t6 = r6 + r06;
    t8 = r6 * r2 + r06 * R0 * R0;
    c6t6 = c6 / t6;
    c6t8 = c6 / t8;
    t27 = T94(f->sr42, iz1, iz2) * c6t8;
    e0 = c6t6 + t27;
    *e = *e - e0 * eps;
    drij = eps * (c6t6 * 6.0 * r4 / t6 + 8.0 * t27 * r6 / t8);
    for (c = 0; c < 3; c++) ga[c] = vab[c] * drij;
    if (r < 25) {
        const double x = T94(f->zab, iz1, iz2), alpha = T94(f->r0ab, iz1, iz2);
        t27 = x * exp(-alpha * r);
        e0 = t27 * oner;
        *e = *e + e0 * eps;
        drij = eps * t27 * (alpha * r + 1.0) * oner / r2;
        for (c = 0; c < 3; c++) ga[c] = ga[c] - vab[c] * drij;
    }
    for (c = 0; c < 3; c++) {
        G(i1, c) += ga[c];
        G(i2, c) -= ga[c];
    }
}

static void nonb_coul_pair(const orc_qmdff *f, const double *xyz, int i1, int i2, double eps, double *e, double *g)
{
    double vab[3], r2, r, sw = 1.0, oner, e0, drij;
    int c;
    ORC_MARK();
    for (c = 0; c < 3; c++) vab[c] = X(i1, c) - X(i2, c);
    if (f->periodic) box_image(f, vab);
    r2 = vab[0] * vab[0] + vab[1] * vab[1] + vab[2] * vab[2];
    r = sqrt(r2);
    if (f->periodic) {
        if (r > f->coul_cut) {
            ORC_REJECT();
            return;
        }
        if (!f->zahn && r > f->cut_low) {
            const double xv = (r - f->cut_low) / (f->coul_cut - f->cut_low);
            sw = exp(1.0) * exp(1.0 / (xv - 1.0));
        }
    }
    if (r > f->coul_cut) {
        ORC_REJECT();
        return;
    }
    oner = 1.0 / r;
    if (f->zahn)
        e0 = f->q[i1 - 1] * f->q[i2 - 1] * ((erfc(f->zahn_a * r) * oner) - f->zahn_par * (r - f->coul_cut));
    else
        e0 = f->q[i1 - 1] * f->q[i2 - 1] * oner * eps * sw;
    *e = *e + e0;
    drij = e0 / r2;
    for (c = 0; c < 3; c++) {
        G(i1, c) += -vab[c] * drij;
        G(i2, c) -= -vab[c] * drij;
    }
}

/* ff_nonb.f90:33-512: ADDS to e and g (as the reference does after ff_eg) */
void orc_ff_nonb(const orc_qmdff *f, const double *xyz, double *e_io, double *g)
{
    double e = 0.0;
    int k, i, j;
    if (f->nnci <= 1 && f->nmols == 0) return;
    for (k = 0; k < f->nnci; k++)
        nonb_vdw_pair(f, xyz, f->nci[3 * k], f->nci[3 * k + 1], f->eps2[f->nci[3 * k + 2] - 1], &e, g);
    if (f->nmols > 1)
        for (i = 1; i <= f->n - 1; i++)
            for (j = i + 1; j <= f->n; j++)
                if (f->molnum[i - 1] != f->molnum[j - 1]) nonb_vdw_pair(f, xyz, i, j, 1.0, &e, g);
    for (k = 0; k < f->nnci; k++)
        nonb_coul_pair(f, xyz, f->nci[3 * k], f->nci[3 * k + 1], f->eps1[f->nci[3 * k + 2] - 1], &e, g);
    if (f->nmols > 1)
        for (i = 1; i <= f->n - 1; i++)
            for (j = i + 1; j <= f->n; j++)
                if (f->molnum[i - 1] != f->molnum[j - 1]) nonb_coul_pair(f, xyz, i, j, 1.0, &e, g);
    *e_io = *e_io + e;
}

/* eabhag.f90:30-210 */
static void eabhag(const orc_qmdff *f, const double *xyz, int A, int B, int H, double ca, double cb, double *energy,
                   double *g)
{
    const double longcut = 8.0, alp7 = 12.0, alp3 = 6.0;
    double drah[3], drbh[3], drab[3], ga[3], gb[3], gh[3], dg[3], dg2[3];
    double rab2, rab, rah2, rah, rbh2, rbh, ratio, rdampl, aprod, cosabh, aterm, apref, rah4, rbh4, denom, da, eabh, gi;
    int c;
    for (c = 0; c < 3; c++) {
        drah[c] = X(A, c) - X(H, c);
        drbh[c] = X(B, c) - X(H, c);
        drab[c] = X(A, c) - X(B, c);
    }
    if (f->periodic) box_image(f, drah); /* only drah is imaged (eabhag.f90:60-62) */
    rab2 = drab[0] * drab[0] + drab[1] * drab[1] + drab[2] * drab[2];
    rab = sqrt(rab2);
    rah2 = drah[0] * drah[0] + drah[1] * drah[1] + drah[2] * drah[2];
    rah = sqrt(rah2);
    rbh2 = drbh[0] * drbh[0] + drbh[1] * drbh[1] + drbh[2] * drbh[2];
    rbh = sqrt(rbh2);
    ratio = pow(rab / longcut, alp7);
    rdampl = 1.0 / (1.0 + ratio);
    rdampl = rdampl / rab2 / rab;
    if (rah2 > rbh2) {
        aprod = 1.0 / rbh / rab;
        cosabh = -(drbh[0] * drab[0] + drbh[1] * drab[1] + drbh[2] * drab[2]) * aprod;
    } else {
        aprod = 1.0 / rah / rab;
        cosabh = (drah[0] * drab[0] + drah[1] * drab[1] + drah[2] * drab[2]) * aprod;
    }
    aterm = 0.5 * (cosabh + 1.0);
    apref = pow(aterm, alp3 - 1);
    aterm = aterm * apref;
    apref = alp3 * 0.5 * apref;
    rah4 = rah2 * rah2;
    rbh4 = rbh2 * rbh2;
    denom = 1.0 / (rah4 + rbh4);
    da = (ca * rah4 + cb * rbh4) * denom;
    eabh = -da * rdampl * aterm;
    if (eabh > -1.e-8) return;
    *energy = *energy + eabh;
    gi = 4.0 * (ca - cb) * rah2 * rbh4 * denom * denom;
    gi = -gi * rdampl * aterm;
    for (c = 0; c < 3; c++) ga[c] = gi * drah[c];
    gi = 4.0 * (cb - ca) * rbh2 * rah4 * denom * denom;
    gi = -gi * rdampl * aterm;
    for (c = 0; c < 3; c++) {
        gb[c] = gi * drbh[c];
        gh[c] = -ga[c] - gb[c];
    }
    gi = rdampl * rdampl * rab * (3.0 + (3.0 + alp7) * ratio);
    gi = gi * da * aterm;
    for (c = 0; c < 3; c++) {
        dg[c] = gi * drab[c];
        ga[c] = ga[c] + dg[c];
        gb[c] = gb[c] - dg[c];
    }
    gi = -da * rdampl * apref;
    if (rah2 > rbh2) {
        for (c = 0; c < 3; c++) {
            dg[c] = gi * (-aprod * drbh[c] - cosabh * drab[c] / rab2);
            dg2[c] = gi * (aprod * drab[c] + cosabh * drbh[c] / rbh2);
            ga[c] = ga[c] + dg[c];
            gh[c] = gh[c] + dg2[c];
            gb[c] = gb[c] - dg[c] - dg2[c];
        }
    } else {
        for (c = 0; c < 3; c++) {
            dg[c] = gi * (-aprod * drah[c] + cosabh * drab[c] / rab2);
            dg2[c] = gi * (-aprod * drab[c] + cosabh * drah[c] / rah2);
            gb[c] = gb[c] + dg[c];
            gh[c] = gh[c] + dg2[c];
            ga[c] = ga[c] - dg[c] - dg2[c];
        }
    }
    for (c = 0; c < 3; c++) {
        G(A, c) += ga[c];
        G(B, c) += gb[c];
        G(H, c) += gh[c];
    }
}

/* eabx.f90:30-110 */
static double eabx(const orc_qmdff *f, const double *xyz, int A, int B, int H, double ca)
{
    const double longcut = 120., alp7 = 6, alp3 = 6;
    double rb[3], rab2, dampl, d2ik, d2jk, xy, term, aterm;
    int c;
    for (c = 0; c < 3; c++) rb[c] = X(A, c) - X(B, c);
    if (f->periodic) box_image(f, rb);
    rab2 = rb[0] * rb[0] + rb[1] * rb[1] + rb[2] * rb[2];
    dampl = 1. / (1. + pow(rab2 / longcut, alp7));
    for (c = 0; c < 3; c++) rb[c] = X(A, c) - X(H, c);
    if (f->periodic) box_image(f, rb);
    d2ik = rb[0] * rb[0] + rb[1] * rb[1] + rb[2] * rb[2];
    for (c = 0; c < 3; c++) rb[c] = X(H, c) - X(B, c);
    if (f->periodic) box_image(f, rb);
    d2jk = rb[0] * rb[0] + rb[1] * rb[1] + rb[2] * rb[2];
    if (d2ik > d2jk) {
        xy = sqrt(rab2 * d2jk);
        term = 0.5 * (rab2 + d2jk - d2ik) / xy;
    } else {
        xy = sqrt(rab2 * d2ik);
        term = 0.5 * (rab2 + d2ik - d2jk) / xy;
    }
    aterm = pow(0.5 * (term + 1.0), alp3);
    return -ca * dampl * aterm / d2jk;
}

/* eabxag.f90:30-150: numeric gradient, step 1e-6; g_local_* = (er-el)*dum assigns the scalar to all
 * three components on every pass, so the z-derivative is what gets added to x, y and z (F9).  The
 * coordinates are perturbed in place and restored by +step (eabxag.f90:108-113), as here. */
static void eabxag(const orc_qmdff *f, double *xyz, int A, int B, int H, double ca, double *energy, double *g)
{
    const double step = 1.e-6, dum = 1. / (2.0 * step);
    const int who[3] = {A, B, H};
    double er, el, gl[3] = {0, 0, 0};
    int w, j, c;
    er = eabx(f, xyz, A, B, H, ca);
    *energy = *energy + er;
    for (w = 0; w < 3; w++) {
        const int at = who[w];
        for (j = 0; j < 3; j++) {
            X(at, j) = X(at, j) + step;
            er = eabx(f, xyz, A, B, H, ca);
            X(at, j) = X(at, j) - step * 2.0;
            el = eabx(f, xyz, A, B, H, ca);
            X(at, j) = X(at, j) + step;
            gl[w] = (er - el) * dum;
        }
        for (c = 0; c < 3; c++) G(at, c) += gl[w];
    }
}

static double hbpara(double a, double b, double q) { return exp(-a * q) / (exp(-a * q) + b); }

static double dist_img(const orc_qmdff *f, const double *xyz, int a, int b, int image)
{
    double r[3];
    int c;
    for (c = 0; c < 3; c++) r[c] = X(a, c) - X(b, c);
    if (image && f->periodic) box_image(f, r);
    return sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
}

/* ff_hb.f90:35-280: ADDS to e and g.  xyz is not const: eabxag perturbs and restores it. */
void orc_ff_hb(const orc_qmdff *f, double *xyz, double *e_io, double *g)
{
    double eh = *e_io;
    int k, i, j;
    if (f->nhb < 1 && f->nmols == 0) return;
    for (k = 0; k < f->nhb; k++) {
        const int i1 = f->hb[3 * k], i2 = f->hb[3 * k + 1], ih = f->hb[3 * k + 2];
        if (dist_img(f, xyz, i1, i2, 1) > 15.0) continue;
        if (f->at[ih - 1] == 1)
            eabhag(f, xyz, i1, i2, ih, f->vhb[2 * k], f->vhb[2 * k + 1], &eh, g);
        else
            eabxag(f, xyz, i1, i2, ih, f->vhb[2 * k], &eh, g);
    }
    if (f->nmols > 1) {
        const double c12 = (double)1.2f, c13 = (double)1.3f, b0 = (double)0.52917726f;
        for (i = 0; i < f->nbond; i++) {
            const int i1 = f->bond[2 * i], i2 = f->bond[2 * i + 1];
            const int z1 = f->at[i1 - 1], z2 = f->at[i2 - 1];
            int at_H = 0, at_A = 0, halogen = 0, hydrogen = 0;
            if (z1 == 17 || z1 == 35 || z1 == 53 || z1 == 85) {
                if (z2 != 1) {
                    at_H = i1;
                    at_A = i2;
                    halogen = 1;
                }
            } else if (z2 == 17 || z2 == 35 || z2 == 53 || z2 == 85) {
                if (z1 != 1) {
                    at_H = i2;
                    at_A = i1;
                    halogen = 1;
                }
            }
            if (halogen)
                for (j = 1; j <= f->n; j++) {
                    const int zj = f->at[j - 1];
                    double ri, rj, dum1, dum2, r;
                    if (f->molnum[at_H - 1] == f->molnum[j - 1]) continue;
                    if (!(zj == 7 || zj == 8)) continue;
                    ORC_MARK();
                    ri = dist_img(f, xyz, at_A, at_H, 1);
                    rj = dist_img(f, xyz, j, at_H, 1);
                    dum1 = c12 * (f->rad[f->at[at_A - 1] - 1] + f->rad[f->at[at_H - 1] - 1]) / b0;
                    dum2 = c12 * (f->rad[zj - 1] + f->rad[f->at[at_H - 1] - 1]) / b0;
                    if (ri < dum1 || rj < dum2) {
                        r = dist_img(f, xyz, at_A, j, 0); /* not imaged (ff_hb.f90:159-161) */
                        if (r > 15.0) {
                            ORC_REJECT();
                            continue;
                        }
                        dum1 = hbpara(-6.5, 1.0, f->q_glob[at_H - 1]);
                        eabxag(f, xyz, at_A, j, at_H, f->scalexb[f->at[at_H - 1] - 1] * dum1, &eh, g);
                    } else {
                        ORC_REJECT();
                    }
                }
            if (z1 == 1 && (z2 == 7 || z2 == 8 || z2 == 9 || z2 == 16 || z2 == 17)) {
                at_H = i1;
                at_A = i2;
                hydrogen = 1;
            }
            if (z2 == 1 && (z1 == 7 || z1 == 8 || z1 == 9 || z1 == 16 || z1 == 17)) {
                at_H = i2;
                at_A = i1;
                hydrogen = 1;
            }
            if (hydrogen)
                for (j = 1; j <= f->n; j++) {
                    const int zj = f->at[j - 1];
                    double cpar, ri, rj, dum1, dum2, r, c1, c2;
                    if (f->molnum[at_H - 1] == f->molnum[j - 1]) continue;
                    cpar = f->scalehb[f->at[at_A - 1] - 1] * f->scalehb[zj - 1];
                    if (!(cpar > 1e-6)) continue;
                    ORC_MARK();
                    ri = dist_img(f, xyz, at_A, at_H, 1);
                    rj = dist_img(f, xyz, j, at_H, 1);
                    dum1 = c13 * (f->rad[f->at[at_A - 1] - 1] + f->rad[0]) / b0;
                    dum2 = c13 * (f->rad[zj - 1] + f->rad[0]) / b0;
                    if (ri < dum1 || rj < dum2) {
                        r = dist_img(f, xyz, at_A, j, 1);
                        if (r > 15.0) {
                            ORC_REJECT();
                            continue;
                        }
                        c2 = hbpara(10.0, 5.0, f->q_glob[at_A - 1]) * f->scalehb[f->at[at_A - 1] - 1];
                        c1 = hbpara(10.0, 5.0, f->q_glob[j - 1]) * f->scalehb[zj - 1];
                        eabhag(f, xyz, j, at_A, at_H, c1, c2, &eh, g);
                    } else {
                        ORC_REJECT();
                    }
                }
        }
    }
    *e_io = eh;
}

/* gradient.f90:341-362, nqmdff = 1 */
void orc_qmdff_egrad(const orc_qmdff *f, const double *xyz, int nimg, double *V, double *g)
{
    int s;
    for (s = 0; s < nimg; s++) {
        double e = 0.0;
        orc_ff_eg(f, xyz + (size_t)s * 3 * f->n, &e, g + (size_t)s * 3 * f->n);
        orc_ff_nonb(f, xyz + (size_t)s * 3 * f->n, &e, g + (size_t)s * 3 * f->n);
        if (f->scalehb) {
            /* eabxag perturbs xyz in place: work on a copy so the caller's array stays const */
            double *tmp = (double *)malloc(sizeof(double) * 3 * f->n);
            memcpy(tmp, xyz + (size_t)s * 3 * f->n, sizeof(double) * 3 * f->n);
            orc_ff_hb(f, tmp, &e, g + (size_t)s * 3 * f->n);
            free(tmp);
        }
        V[s] = e + f->e_zero;
    }
}
