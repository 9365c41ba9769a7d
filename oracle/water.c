/*
 * water.c -- CPU oracle: the flexible SPC water box of pes WATER_SPC.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_real.h).  Parity UNPINNED by the reference (no vectors, cannot be
 * compiled here); pinned in tests/ by finite differences of the intramolecular part and of the plain-Coulomb
 * form, rigid-motion invariance, the water monomer minimum the parameters encode and the dimer well.
 *
 * Literal restatement of
 *   /root/reference/src/egrad_water.f90:36-333   orc_water_egrad (called per structure from gradient.f90:212-213;
 *                                                the virial blocks belong to the NPT barostat, out of scope)
 *   /root/reference/src/water_init.f90:75-107    orc_water_default_pars (+ the charges the caller lays out)
 *   /root/reference/src/box_image.f90            wat_box_image
 * Reproduced as written (SURVEY.md F9 style quirks):
 *   - the Lennard-Jones block tests name(i) twice (egrad_water.f90:295: `name(i) .eq. "O" .and. name(i) .eq. "O"`),
 *     so every pair whose FIRST atom is an oxygen gets the O-O Lennard-Jones term, O-H pairs included;
 *   - the Coulomb gradient of the Zahn form reuses e0/r^2 (:279), as ff_nonb.f90:384 does;
 *   - non-periodic systems have no cut-off (`rij .lt. coul_cut .or. .not. periodic`, :268);
 *   - water_pars(7) = 111.70765 and water_pars(11) = 0.1554 are REAL*4 literals (water_init.f90:82,86; F3).
 */
#include <math.h>
#include <string.h>
#include "oracle_real.h"
#include "water.h"

void orc_water_default_pars(double p[11])
{
    const double bohr = 0.52917721092, hartree = 627.5094743; /* general.f90:256-257 */
    p[0] = 1.0;          /* r_0 (Angstrom) */
    p[1] = 1.633;        /* r_0HH */
    p[2] = 101.9188;     /* D_e (kcal/mol) */
    p[3] = 2.567;        /* a (1/Angstrom) */
    p[4] = 328.645606;   /* k_theta */
    p[5] = -211.4672;    /* k_rtheta */
    p[6] = F(111.70765); /* k_rr */
    p[7] = 0.41;         /* e_H */
    p[8] = -0.82;        /* e_O */
    p[9] = 3.166;        /* sigma_OO */
    p[10] = F(0.1554);   /* eps_OO */
    p[0] = p[0] / bohr;
    p[1] = p[1] / bohr;
    p[2] = p[2] / hartree;
    p[3] = p[3] * bohr;
    p[4] = p[4] / hartree * bohr * bohr;
    p[5] = p[5] / hartree * bohr * bohr;
    p[6] = p[6] / hartree * bohr * bohr;
    p[9] = p[9] / bohr;
    p[10] = p[10] / hartree;
}

static void wat_box_image(const orc_water *w, double v[3])
{
    int d;
    for (d = 0; d < 3; d++) {
        const double L = w->box[d], L2 = 0.5 * w->box[d];
        while (fabs(v[d]) > L2) v[d] = v[d] - (v[d] >= 0 ? L : -L);
    }
}
static double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

static void water_one(const orc_water *w, const double *x, double *e_out, double *g)
{
    const double r_zero = w->pars[0], r_0HH = w->pars[1], d_e = w->pars[2], a_par = w->pars[3], k_theta = w->pars[4],
                 k_rtheta = w->pars[5], k_rr = w->pars[6], sigma_OO = w->pars[9], eps_OO = w->pars[10];
    const int n = w->natoms, nwater = n / 3;
    double e_act = 0.0;
    int i, j, d;
    memset(g, 0, sizeof(double) * 3 * n);
#define X(a, d) x[3 * (a) + (d)]
#define G(a, d) g[3 * (a) + (d)]
    for (i = 0; i < nwater; i++) {
        const int iO = 3 * i, iH1 = 3 * i + 1, iH2 = 3 * i + 2;
        double OH1[3], OH2[3], HH[3], ga[3], gb[3];
        double r_OH1, r_OH2, r_HH, dr_OH1, dr_OH2, dr_HH, exp1, exp2, de1, de2;
        for (d = 0; d < 3; d++) OH1[d] = X(iO, d) - X(iH1, d);
        if (w->periodic) wat_box_image(w, OH1);
        for (d = 0; d < 3; d++) OH2[d] = X(iO, d) - X(iH2, d);
        if (w->periodic) wat_box_image(w, OH2);
        for (d = 0; d < 3; d++) HH[d] = X(iH1, d) - X(iH2, d);
        if (w->periodic) wat_box_image(w, HH);
        r_OH1 = sqrt(dot3(OH1, OH1));
        dr_OH1 = r_OH1 - r_zero;
        r_OH2 = sqrt(dot3(OH2, OH2));
        dr_OH2 = r_OH2 - r_zero;
        r_HH = sqrt(dot3(HH, HH));
        dr_HH = r_HH - r_0HH;
        exp1 = exp(a_par * dr_OH1);
        exp2 = exp(a_par * dr_OH2);
        e_act = e_act + d_e * ((1.0 - exp1) * (1.0 - exp1));
        e_act = e_act + d_e * ((1.0 - exp2) * (1.0 - exp2));
        de1 = -2.0 * a_par * d_e * exp1 * (1.0 - exp1) / r_OH1;
        for (d = 0; d < 3; d++) {
            ga[d] = de1 * OH1[d];
            G(iH1, d) = G(iH1, d) - ga[d];
            G(iO, d) = G(iO, d) + ga[d];
        }
        de2 = -2.0 * a_par * d_e * exp2 * (1.0 - exp2) / r_OH2;
        for (d = 0; d < 3; d++) {
            ga[d] = de2 * OH2[d];
            G(iH2, d) = G(iH2, d) - ga[d];
            G(iO, d) = G(iO, d) + ga[d];
        }
        e_act = e_act + 0.5 * k_theta * (dr_HH * dr_HH);
        for (d = 0; d < 3; d++) {
            ga[d] = k_theta * HH[d] * dr_HH / r_HH;
            G(iH1, d) = G(iH1, d) + ga[d];
            G(iH2, d) = G(iH2, d) - ga[d];
        }
        e_act = e_act + k_rtheta * dr_HH * (dr_OH1 + dr_OH2);
        for (d = 0; d < 3; d++) {
            ga[d] = dr_HH / r_OH1 * OH1[d] * k_rtheta;
            gb[d] = dr_HH / r_OH2 * OH2[d] * k_rtheta;
        }
        for (d = 0; d < 3; d++) {
            G(iH1, d) = G(iH1, d) + k_rtheta * 1.0 * (dr_OH1 + dr_OH2) / r_HH * HH[d] - ga[d];
            G(iH2, d) = G(iH2, d) - k_rtheta * 1.0 * (dr_OH1 + dr_OH2) / r_HH * HH[d] - gb[d];
            G(iO, d) = G(iO, d) + ga[d] + gb[d];
        }
        e_act = e_act + k_rr * dr_OH1 * dr_OH2;
        for (d = 0; d < 3; d++) {
            ga[d] = k_rr * (dr_OH2 / r_OH1 * OH1[d]);
            gb[d] = k_rr * (dr_OH1 / r_OH2 * OH2[d]);
        }
        for (d = 0; d < 3; d++) {
            G(iH1, d) = G(iH1, d) - ga[d];
            G(iH2, d) = G(iH2, d) - gb[d];
            G(iO, d) = G(iO, d) + gb[d] + ga[d];
        }
    }
    for (i = 0; i < n; i++) {
        for (j = i + 1; j < n; j++) {
            double dv[3], rij;
            if (i / 3 == j / 3) continue; /* water_act(i) .ne. water_act(j) */
            for (d = 0; d < 3; d++) dv[d] = X(i, d) - X(j, d);
            if (w->periodic) wat_box_image(w, dv);
            rij = sqrt(dot3(dv, dv));
            if (rij < w->coul_cut || !w->periodic) {
                const double oner = 1.0 / rij;
                double e0, gv[3];
                if (w->zahn)
                    e0 = w->q[i] * w->q[j] * ((erfc(w->zahn_a * rij) * oner) - w->zahn_par * (rij - w->coul_cut));
                else
                    e0 = w->q[i] * w->q[j] * oner;
                e_act = e_act + e0;
                for (d = 0; d < 3; d++) {
                    gv[d] = e0 * oner * oner * dv[d];
                    G(i, d) = G(i, d) - gv[d];
                    G(j, d) = G(j, d) + gv[d];
                }
                if (w->is_O[i] && w->is_O[i]) { /* sic: egrad_water.f90:295 tests name(i) twice */
                    const double s2 = (sigma_OO * oner) * (sigma_OO * oner), sigr6 = s2 * s2 * s2;
                    e_act = e_act + 4.0 * eps_OO * (sigr6 * sigr6 - sigr6);
                    for (d = 0; d < 3; d++) {
                        gv[d] = 24.0 * eps_OO * oner * oner * dv[d] * sigr6 * (2.0 * sigr6 - 1.0);
                        G(i, d) = G(i, d) - gv[d];
                        G(j, d) = G(j, d) + gv[d];
                    }
                }
            }
        }
    }
#undef X
#undef G
    *e_out = e_act;
}

void orc_water_egrad(const orc_water *w, const double *xyz, int nimg, double *V, double *g)
{
    int k;
    for (k = 0; k < nimg; k++) water_one(w, xyz + (size_t)k * 3 * w->natoms, &V[k], g + (size_t)k * 3 * w->natoms);
}
