/*
 * dgevb.c -- CPU oracle: second QMDFF (the *_two routines) and the DG-EVB coupling / mixing.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Parity UNPINNED by the reference.
 *
 * Literal restatement of
 *   ff_eg_two.f90:40-344, ff_nonb_two.f90:33-158, ff_hb_two.f90:30-71   orc_qmdff_two_egrad
 *       (same terms as ff_eg / ff_nonb / ff_hb on the second table set, never periodic, no
 *        inter-molecular loops, Coulomb q_i q_j eps1/r without cut-off; corr_nonb off)
 *   gradient.f90:365-537          orc_dgevb_egrad   E = (E1+E2)/2 - sqrt(((E1-E2)/2)^2 + V12^2)
 *   xyz_2int.f90:33-79, dist.f90, ang.f90, dihed.f90, oop.f90            internal coordinates
 *   calc_wilson.f90:114-178       Wilson B matrix -- ALWAYS numeric: init_int.f90:142 sets
 *                                 num_wilson=.true. unconditionally; central differences, shift 1e-3
 *   int2grad.f90:30-96            g_x = B^T g_q
 *   sum_v12.f90:30-144, sum_dv12.f90:30-186, deltaq.f90   distributed Gaussians, modes 1-3
 *       (sum_dv12 mode 3 uses `inc` before it is initialised, :157 -- F9; restated with inc = 0,
 *        which is what an untouched stack slot gives)
 */
#include <stdlib.h>
#include <string.h>
#include "oracle_real.h"
#include "qmdff.h"
#include "dgevb.h"

#define X(i, c) xyz[3 * ((i) - 1) + (c)]
#define G(i, c) g[3 * ((i) - 1) + (c)]
#define T94(tab, a, b) ((tab)[((a) - 1) + 94 * ((b) - 1)])

/* ff_nonb_two.f90:33-158 (ADDS to e and g) */
static void ff_nonb_two(const orc_qmdff *f, const double *xyz, double *e_io, double *g)
{
    double e = 0.0;
    int k;
    if (f->nnci <= 1) return;
    for (k = 0; k < f->nnci; k++) {
        const int i1 = f->nci[3 * k], i2 = f->nci[3 * k + 1], nk = f->nci[3 * k + 2] - 1;
        const double dx = X(i1, 0) - X(i2, 0), dy = X(i1, 1) - X(i2, 1), dz = X(i1, 2) - X(i2, 2);
        const double r2 = dx * dx + dy * dy + dz * dz, r = sqrt(r2), oner = 1.0 / r;
        const int iz1 = f->at[i1 - 1], iz2 = f->at[i2 - 1];
        const double R0 = T94(f->r094, iz1, iz2), c6 = f->c6xy[(i2 - 1) + (size_t)f->n * (i1 - 1)];
        const double r4 = r2 * r2, r6 = r4 * r2, r06 = R0 * R0 * R0 * R0 * R0 * R0;
        const double t6 = r6 + r06, t8 = r6 * r2 + r06 * R0 * R0, c6t6 = c6 / t6, c6t8 = c6 / t8;
        double t27 = T94(f->sr42, iz1, iz2) * c6t8, e0 = c6t6 + t27, drij;
        e = e - e0 * f->eps2[nk];
        drij = f->eps2[nk] * (c6t6 * 6.0 * r4 / t6 + 8.0 * t27 * r6 / t8);
        e0 = f->q[i1 - 1] * f->q[i2 - 1] * oner * f->eps1[nk];
        G(i1, 0) += dx * drij;
        G(i1, 1) += dy * drij;
        G(i1, 2) += dz * drij;
        G(i2, 0) -= dx * drij;
        G(i2, 1) -= dy * drij;
        G(i2, 2) -= dz * drij;
        e = e + e0;
        drij = e0 / r2;
        G(i1, 0) -= dx * drij;
        G(i1, 1) -= dy * drij;
        G(i1, 2) -= dz * drij;
        G(i2, 0) += dx * drij;
        G(i2, 1) += dy * drij;
        G(i2, 2) += dz * drij;
        if (r < 25) {
            const double x = T94(f->zab, iz1, iz2), alpha = T94(f->r0ab, iz1, iz2);
            t27 = x * exp(-alpha * r);
            e0 = t27 * oner;
            e = e + e0 * f->eps2[nk];
            drij = f->eps2[nk] * t27 * (alpha * r + 1.0) * oner / r2;
            G(i1, 0) -= dx * drij;
            G(i1, 1) -= dy * drij;
            G(i1, 2) -= dz * drij;
            G(i2, 0) += dx * drij;
            G(i2, 1) += dy * drij;
            G(i2, 2) += dz * drij;
        }
    }
    *e_io = *e_io + e;
}

/* second QMDFF of one structure: ff_eg_two + ff_nonb_two + ff_hb_two (gradient.f90:364-368).
 * The table set must be non-periodic with nmols <= 1 (that is what the *_two routines assume). */
void orc_qmdff_two_one(const orc_qmdff *f2, const double *xyz, double *e_out, double *g)
{
    double e = 0.0;
    orc_ff_eg(f2, xyz, &e, g);            /* ff_eg_two == ff_eg without the periodic branches */
    ff_nonb_two(f2, xyz, &e, g);
    if (f2->scalehb && f2->nhb >= 1) {    /* ff_hb_two: list part only */
        orc_qmdff tmp = *f2;
        double *xt = (double *)malloc(sizeof(double) * 3 * f2->n);
        tmp.nmols = 1;
        memcpy(xt, xyz, sizeof(double) * 3 * f2->n);
        orc_ff_hb(&tmp, xt, &e, g);
        free(xt);
    }
    *e_out = e;
}

/* ---- internal coordinates (dist.f90, ang.f90, dihed.f90, oop.f90) ---- */
static double ic_dist(const double *xyz, int a1, int a2)
{
    return sqrt((X(a2, 0) - X(a1, 0)) * ((X(a2, 0) - X(a1, 0))) +
                ((X(a2, 1) - X(a1, 1)) * ((X(a2, 1) - X(a1, 1))) + ((X(a2, 2) - X(a1, 2)) * ((X(a2, 2) - X(a1, 2))))));
}
static double ic_ang(const double *xyz, int a1, int a2, int a3)
{
    const double ax = X(a1, 0) - X(a2, 0), ay = X(a1, 1) - X(a2, 1), az = X(a1, 2) - X(a2, 2);
    const double bx = X(a3, 0) - X(a2, 0), by = X(a3, 1) - X(a2, 1), bz = X(a3, 2) - X(a2, 2);
    return acos((ax * bx + ay * by + az * bz) / (sqrt(ax * ax + ay * ay + az * az) * sqrt(bx * bx + by * by + bz * bz)));
}
static void cr(const double a[3], const double b[3], double c[3])
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
static double dt3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double ic_dihed(const double *xyz, int a1, int a2, int a3, int a4)
{
    double u[3], v[3], w[3], uxw[3], vxw[3], ul, vl, wl, su, sv, cv;
    int c;
    for (c = 0; c < 3; c++) {
        u[c] = X(a1, c) - X(a2, c);
        v[c] = X(a4, c) - X(a3, c);
        w[c] = X(a3, c) - X(a2, c);
    }
    ul = sqrt(dt3(u, u));
    vl = sqrt(dt3(v, v));
    wl = sqrt(dt3(w, w));
    for (c = 0; c < 3; c++) {
        u[c] = u[c] / ul;
        v[c] = v[c] / vl;
        w[c] = w[c] / wl;
    }
    cr(u, w, uxw);
    cr(v, w, vxw);
    su = sqrt(1.0 - dt3(u, w) * dt3(u, w));
    sv = sqrt(1.0 - dt3(v, w) * dt3(v, w));
    cv = dt3(uxw, vxw) / (su * sv);
    if (cv >= 1.0)
        cv = 1.0;
    else if (cv <= -1.0)
        cv = -1.0;
    return acos(cv);
}
static double ic_oop(const double *xyz, int a1, int a2, int a3, int a4)
{
    double v41[3], v42[3], v43[3], c12[3], c23[3], c31[3], nv[3], l1, l2, l3;
    int c;
    for (c = 0; c < 3; c++) {
        v41[c] = X(a4, c) - X(a1, c);
        v42[c] = X(a4, c) - X(a2, c);
        v43[c] = X(a4, c) - X(a3, c);
    }
    l1 = sqrt(dt3(v41, v41));
    l2 = sqrt(dt3(v42, v42));
    l3 = sqrt(dt3(v43, v43));
    for (c = 0; c < 3; c++) {
        v41[c] = v41[c] / l1;
        v42[c] = v42[c] / l2;
        v43[c] = v43[c] / l3;
    }
    cr(v41, v42, c12);
    cr(v42, v43, c23);
    cr(v43, v41, c31);
    for (c = 0; c < 3; c++) nv[c] = c12[c] + c23[c] + c31[c];
    return dt3(v41, c23) / sqrt(dt3(nv, nv));
}
static double ic_eval(const orc_dgevb *d, const double *xyz, int i)
{
    const int *cd = d->coord_def + 5 * i;
    switch (cd[0]) {
    case 1: return ic_dist(xyz, cd[1], cd[2]);
    case 2: return ic_ang(xyz, cd[1], cd[2], cd[3]);
    case 3: return ic_dihed(xyz, cd[1], cd[2], cd[3], cd[4]);
    case 4: return ic_oop(xyz, cd[1], cd[2], cd[3], cd[4]);
    }
    return 0.0;
}
void orc_xyz_2int(const orc_dgevb *d, const double *xyz, double *internal)
{
    int i;
    for (i = 0; i < d->nat6; i++) internal[i] = ic_eval(d, xyz, i);
}

/* calc_wilson.f90:114-178 (num_wilson branch): B(nat6, 3*natoms), stored B[i*3n + col] */
static void wilson_num(const orc_dgevb *d, const double *xyz5, double *B)
{
    const int n3 = 3 * d->natoms;
    const double shift = 0.001;
    double *xyz = (double *)malloc(sizeof(double) * n3);
    int i, j, k;
    memcpy(xyz, xyz5, sizeof(double) * n3);
    memset(B, 0, sizeof(double) * n3 * d->nat6);
    for (i = 0; i < d->nat6; i++) {
        const int *cd = d->coord_def + 5 * i;
        const int act_num = (cd[0] == 1) ? 2 : (cd[0] == 2 ? 3 : 4);
        int l = 0, act_atom = cd[1];
        for (j = 1; j <= 3 * act_num; j++) {
            const int m = j - ((j - 1) / 3 * 3);
            double lo = 0, hi = 0;
            if ((j + 2) % 3 == 0) {
                l = l + 1;
                act_atom = cd[l];
            }
            for (k = 1; k <= 2; k++) {
                if (k == 1)
                    X(act_atom, m - 1) = X(act_atom, m - 1) - shift;
                else
                    X(act_atom, m - 1) = X(act_atom, m - 1) + 2 * shift;
                if (k == 1)
                    lo = ic_eval(d, xyz, i);
                else
                    hi = ic_eval(d, xyz, i);
            }
            B[(size_t)i * n3 + (act_atom - 1) * 3 + (m - 1)] = (hi - lo) / (2 * shift);
            memcpy(xyz, xyz5, sizeof(double) * n3);
        }
    }
    free(xyz);
}

/* sum_v12.f90 */
static double sum_v12(const orc_dgevb *d, const double *act)
{
    const int nat6 = d->nat6, mode = d->mode;
    const double *b = d->b_vec - 1; /* 1-based */
    double *qq = (double *)malloc(sizeof(double) * nat6);
    double V12 = 0.0;
    int j, k, l;
    for (j = 1; j <= d->npoints; j++) {
        const double *pt = d->point_int + (size_t)(j - 1) * nat6;
        const double al = d->alph[j - 1];
        double d_p = 0.0, expo;
        for (k = 0; k < nat6; k++) {
            qq[k] = act[k] - pt[k];
            d_p += qq[k] * qq[k];
        }
        expo = exp(-0.5 * al * d_p);
        if (expo < d->g_thres) continue;
        if (mode == 1) {
            V12 = V12 + b[j] * (1 + 0.5 * al * d_p) * expo;
        } else if (mode == 2) {
            V12 = V12 + b[(j - 1) * nat6 + j] * (1 + 0.5 * al * d_p) * expo;
            for (k = 1; k <= nat6; k++) V12 = V12 + b[j + (j - 1) * nat6 + k] * qq[k - 1] * expo;
        } else {
            const int block = 1 + nat6 + nat6 * (nat6 + 1) / 2, first = 1 + nat6;
            int inc = 0;
            V12 = V12 + b[(j - 1) * (block - 1) + j] * (1 + 0.5 * al * d_p) * expo;
            for (k = 1; k <= nat6; k++) V12 = V12 + b[k + (j - 1) * block + 1] * qq[k - 1] * expo;
            for (k = 1; k <= nat6; k++)
                for (l = k; l <= nat6; l++) {
                    inc = inc + 1;
                    if (k == l)
                        V12 = V12 + b[inc + (j - 1) * block + first] * 0.5 * qq[k - 1] * qq[l - 1] * expo;
                    else
                        V12 = V12 + b[inc + (j - 1) * block + first] * qq[k - 1] * qq[l - 1] * expo;
                }
        }
    }
    free(qq);
    return V12;
}

/* sum_dv12.f90 */
static void sum_dv12(const orc_dgevb *d, const double *act, double *gq)
{
    const int nat6 = d->nat6, mode = d->mode;
    const double *b = d->b_vec - 1;
    double *qq = (double *)malloc(sizeof(double) * nat6);
    int j, k, l, m, inc = 0;
    memset(gq, 0, sizeof(double) * nat6);
    for (j = 1; j <= d->npoints; j++) {
        const double *pt = d->point_int + (size_t)(j - 1) * nat6;
        const double al = d->alph[j - 1];
        double d_p = 0.0, expo;
        for (k = 0; k < nat6; k++) {
            qq[k] = act[k] - pt[k];
            d_p += qq[k] * qq[k];
        }
        expo = exp(-0.5 * al * d_p);
        if (expo < d->g_thres) continue;
        if (mode == 1) {
            for (k = 1; k <= nat6; k++) gq[k - 1] = gq[k - 1] - 0.5 * al * al * b[j] * d_p * qq[k - 1] * expo;
        } else if (mode == 2) {
            for (l = 1; l <= nat6; l++) {
                gq[l - 1] = gq[l - 1] - 0.5 * al * al * b[(j - 1) * nat6 + j] * d_p * qq[l - 1] * expo;
                for (k = 1; k <= nat6; k++) {
                    if (k == l)
                        gq[l - 1] = gq[l - 1] - b[j + (j - 1) * nat6 + k] * (al * qq[k - 1] * qq[l - 1] - 1.0) * expo;
                    else
                        gq[l - 1] = gq[l - 1] - b[j + (j - 1) * nat6 + k] * al * qq[l - 1] * qq[k - 1] * expo;
                }
            }
        } else {
            const int block = 1 + nat6 + nat6 * (nat6 + 1) / 2, first = 1 + nat6;
            for (l = 1; l <= nat6; l++) {
                gq[l - 1] = gq[l - 1] - 0.5 * al * al * b[(j - 1) * (block - 1) + j] * d_p * qq[l - 1] * expo;
                for (k = 1; k <= nat6; k++) {
                    if (k == l)
                        gq[l - 1] = gq[l - 1] - b[k + (j - 1) * block + 1] * (al * qq[k - 1] * qq[l - 1] - 1.0) * expo;
                    else
                        gq[l - 1] = gq[l - 1] - b[k + (j - 1) * block + 1] * al * qq[l - 1] * qq[k - 1] * expo;
                }
                for (k = 1; k <= nat6; k++)
                    for (m = k; m <= nat6; m++) {
                        const double bb = b[inc + 1 + (j - 1) * block + first];
                        inc = inc + 1;
                        if (m == k) {
                            if (l == k)
                                gq[l - 1] = gq[l - 1] - 0.5 * bb * qq[k - 1] * (al * qq[k - 1] * qq[k - 1] - 2.0) * expo;
                            else
                                gq[l - 1] = gq[l - 1] - 0.5 * bb * al * qq[k - 1] * qq[k - 1] * qq[l - 1] * expo;
                        } else {
                            if (l == k)
                                gq[l - 1] = gq[l - 1] - bb * qq[m - 1] * (al * qq[k - 1] * qq[k - 1] - 1.0) * expo;
                            else if (l == m)
                                gq[l - 1] = gq[l - 1] - bb * qq[k - 1] * (al * qq[m - 1] * qq[m - 1] - 1.0) * expo;
                            else
                                gq[l - 1] = gq[l - 1] - bb * al * qq[m - 1] * qq[k - 1] * qq[l - 1] * expo;
                        }
                    }
                inc = 0;
            }
        }
    }
    free(qq);
}

/* gradient.f90:365-537, dg_evb branch, analytic (num_grad = .false.) */
void orc_dgevb_egrad(const orc_qmdff *f1, const orc_qmdff *f2, const orc_dgevb *d, const double *xyz_all, int nimg,
                     double *V, double *g_all)
{
    const int n = f1->n, n3 = 3 * n;
    double *g1 = (double *)malloc(sizeof(double) * n3), *g2 = (double *)malloc(sizeof(double) * n3);
    double *internal = (double *)malloc(sizeof(double) * d->nat6), *gq = (double *)malloc(sizeof(double) * d->nat6);
    double *B = (double *)malloc(sizeof(double) * n3 * d->nat6), *gv = (double *)malloc(sizeof(double) * n3);
    int s, i, j;
    for (s = 0; s < nimg; s++) {
        const double *xyz = xyz_all + (size_t)s * n3;
        double *g = g_all + (size_t)s * n3;
        double e1, e2, e1s, e2s, ediff, V12, off4, root2 = 0.0, root;
        int unset;
        orc_qmdff_egrad(f1, xyz, 1, &e1, g1); /* returns e + E_zero1 */
        e1s = e1;
        orc_qmdff_two_one(f2, xyz, &e2, g2);
        e2s = e2 + f2->e_zero;
        ediff = e1s - e2s;
        orc_xyz_2int(d, xyz, internal);
        V12 = sum_v12(d, internal);
        sum_dv12(d, internal, gq);
        wilson_num(d, xyz, B);
        for (j = 0; j < n3; j++) {
            double acc = 0.0;
            for (i = 0; i < d->nat6; i++) acc += B[(size_t)i * n3 + j] * gq[i];
            gv[j] = acc;
        }
        off4 = 4.0 * V12;
        unset = (ediff * ediff + off4 < 0.0);
        if (!unset) root2 = sqrt(ediff * ediff + off4);
        for (j = 0; j < n3; j++) {
            const double deldiscr = ediff * (g1[j] - g2[j]) + 2.0 * gv[j];
            const double delsqrt = unset ? 0.0 : deldiscr / root2;
            g[j] = 0.5 * (g1[j] + g2[j] - delsqrt);
        }
        root = (0.5 * (e1s - e2s)) * (0.5 * (e1s - e2s)) + V12;
        V[s] = (root <= 0) ? 0.5 * (e1s + e2s) : 0.5 * (e1s + e2s) - sqrt(root);
    }
    free(g1);
    free(g2);
    free(internal);
    free(gq);
    free(B);
    free(gv);
}
