/*
 * pes_ch4h.c -- CPU oracle: CBE-2009 CH4 + H -> CH3 + H2 potential energy surface
 * (Corchado, Bravo, Espinosa-Garcia, J. Chem. Phys. 130, 184314 (2009); POTLIB form).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_real.h).  Parity UNPINNED by the reference
 * (no golden vectors, cannot be compiled here); pinned by per-term finite differences,
 * permutation symmetry and the literature barrier/asymptotes in tests/.
 *
 * Literal restatement of /root/reference/src/egrad_ch4h.f.  COMMON blocks become the
 * struct ch4h_state; arrays keep the reference's 1-based indices (element 0 unused):
 *   egrad_ch4h       :74-132     oracle_egrad_ch4h_real
 *   POT_ch4h         :161-285    ch4h_pot   (CARTOU/CARTTOR/EUNITZERO/RTOCART/DEDCOU are
 *                                identity maps for NFLAG(1)=NFLAG(2)=1, ICARTR=1, ANUZERO=0:
 *                                util_ch4h.f:683-731,737-815,904-916,1150-1206,1265-1275,1507)
 *   coorden_ch4h     :287-365    ch4h_coorden
 *   refangles_ch4h   :367-504    ch4h_refangles
 *   stretch_ch4h     :506-711    ch4h_stretch
 *   opbend_ch4h      :713-863    ch4h_opbend
 *   ipbend_ch4h      :865-984    ch4h_ipbend
 *   calcdelta_ch4h   :986-1238   ch4h_calcdelta
 *   opforce_ch4h     :1240-1348  ch4h_opforce
 *   ipforce_ch4h     :1350-1543  ch4h_ipforce
 *   switchf_ch4h     :1545-1685  ch4h_switchf
 *   PREPOT_ch4h      :1687-1796  ch4h_prepot (one-shot unit scaling applied at init here)
 *   BLOCK DATA       :1798-1888  constants (all d0 -> clean doubles)
 * The reference does not clamp acos / 1/sqrt(1-x^2) arguments; neither does the oracle.
 *
 * With -DCBE_CH4OH (set by pes_ch4oh.c, which includes this file) the same routines restate
 * /root/reference/src/egrad_ch4oh.f (Espinosa-Garcia, Corchado, J. Chem. Phys. 112, 5731 (2000);
 * CH4 + OH -> CH3 + H2O, 7 atoms): that file is this template with the "b" atom an oxygen, its own
 * BLOCK DATA (:2066-2106), and three additions marked CBE_CH4OH below -- the C-O triplet depth
 * switched on the mean C-H distance (:590-592, :681-699), the O-H Morse bond (:616-621, :644-659, :769-774)
 * and the four H-O-H bends with a tanh-switched force constant (:1059-1171).
 * With -DCBE_GEH4OH in addition (pes_geh4oh.c) they restate /root/reference/src/egrad_geh4oh.f (GeH4 + OH ->
 * GeH3 + H2O): the CH4 + OH file again with its own BLOCK DATA (:2002-2042), the in-plane reference angle built
 * on taugeh = 0.678 pi instead of pi/2 (:368, :382-440) and sphi evaluated at every distance (:1793-1804).
 * With -DCBE_CH4CN in addition to CBE_CH4OH (pes_ch4cn.c) they restate /root/reference/src/egrad_ch4cn.f (CH4 + CN ->
 * CH3 + HCN; Espinosa-Garcia, Rangel, Suleimanov, PCCP 19, 19341 (2017)): the CH4 + OH file with the abstracting atom
 * the carbon of CN and the seventh atom its nitrogen, its own BLOCK DATA (:2074-2114), and the Morse term of the
 * seventh atom on literal constants (:625-627, :659: r0 = 1.172, a = 0.80, D = 80.0 -- D is NOT scaled by fact1).
 */
#include "oracle_real.h"
#include "oracle.h"

#define CH4H_PI 3.141592653589793

typedef struct {
    /* /POTCM_ch4h/ after PREPOT scaling */
    double r0ch, d1ch, d3ch, a1ch, b1ch, c1ch, r0hh, d1hh, d3hh, ahh, r0cb, d1cb, d3cb, acb, a3s,
        b3s, aphi, bphi, cphi, atheta, btheta, ctheta, fch3, hch3, fkinf, ak, bk, aa1, aa2, aa3,
        aa4;
    int nc[4], nhb[4], nh[5][4]; /* /ndx/ */
#ifdef CBE_CH4OH
    double d3cbi, a3cb, b3cb, rcbsp, fkh2oeq, alph2o, anh2oeq; /* egrad_ch4oh.f:2080,2101-2106 */
    int no[4];
#endif
} ch4h_par;

typedef struct {
    real theta0[5][5], dtheta0[5][5][5];                  /* /angles/  */
    real rcb, rch[5], rbh[5];                             /* /bonds/   */
    real tcb[4], tch[5][4], tbh[5][4];                    /* /coords/  */
    real fdelta[5], hdelta[5];                            /* /delta1/  */
    real dfdelta[5][5], dhdelta[5][5];                    /* /delta2/  */
    real fk0[5][5], f1[5], dfdc[5][5][5], dfdh[5][5][5];  /* /force1/  */
    real a1s, b1s, a2s, b2s;                              /* /fsw1/    */
    real s1[5], ds1[5], s2[5], ds2[5];                    /* /ip1/     */
    real s3[5], ds3[5];                                   /* /op1/     */
#ifdef CBE_CH4OH
    real q[22], pdot[22];                                 /* /qpdot_pl/ */
    real rno, tno[4];                                     /* /bonds/, /coords/ */
#else
    real q[19], pdot[19];                                 /* /qpdot_pl/ */
#endif
    real sphi[5], dsphi[5], stheta[5], dstheta[5];        /* /switch1/ */
} ch4h_state;

/* BLOCK DATA PTPACM_ch4h (:1852-1886) followed by PREPOT_ch4h (:1758-1793) */
static void ch4h_prepot(ch4h_par *p)
{
    const int nnc = 2, nnb = 6;
    const int nnh[5] = {0, 3, 4, 5, 1};
    const double fact1 = 0.041840, fact2 = 6.022045;
    int ind, i, icount;
#ifdef CBE_CH4OH
    const int nno = 7;
    const double fact3 = 2.0 * CH4H_PI / 360.0;
#endif
#ifdef CBE_CH4OH
    /* BLOCK DATA PTPACM_ch4oh (egrad_ch4oh.f:2066-2106), PREPOT_ch4oh (:1972-2002) */
    p->r0ch = 1.09397;
    p->d1ch = 112.17000;
    p->d3ch = 32.65328;
    p->a1ch = 1.78000;
    p->b1ch = 0.15000;
    p->c1ch = 15.00000;
    p->r0hh = 0.97060;
    p->d1hh = 125.44000;
    p->d3hh = 20.41017;
    p->ahh = 2.15000;
    p->r0cb = 1.49092;
    p->d1cb = 91.47526;
    p->d3cbi = 112.69509;
    p->acb = 2.98688;
    p->a3s = 0.1419100;
    p->b3s = -0.3068400;
    p->aphi = 0.5287900;
    p->bphi = 0.4006600;
    p->cphi = 1.9209900;
    p->atheta = 0.9078700;
    p->btheta = 0.3548900;
    p->ctheta = 1.8915500;
    p->fch3 = 0.0740000;
    p->hch3 = 0.1915000;
    p->fkinf = 0.4400000;
    p->ak = 0.1260000;
    p->bk = 10.7132;
    p->aa1 = 0.303746;
    p->aa2 = 1.599960;
    p->aa3 = 3.216595;
    p->aa4 = 11.569980;
    p->fkh2oeq = 0.7300000;
    p->alph2o = 1.1080000;
    p->anh2oeq = 104.7132000;
    p->a3cb = 0.000000;
    p->b3cb = 0.900000;
    p->rcbsp = 2.606485;
    p->d3cb = 0.0; /* a function of the geometry here, see ch4h_stretch */
    for (ind = 1; ind <= 3; ind++) p->no[ind] = 3 * nno + ind - 3;
#ifdef CBE_GEH4OH
    /* BLOCK DATA PTPACM_geh4oh (egrad_geh4oh.f:2002-2042): the entries that differ from CH4 + OH */
    p->r0ch = 1.52500;
    p->d1ch = 86.50000;
    p->d3ch = 41.50000;
    p->a1ch = 1.43925;
    p->b1ch = 0.12330;
    p->c1ch = 2.00400;
    p->d1hh = 120.94800;
    p->d3hh = 31.86417;
    p->ahh = 2.18200;
    p->r0cb = 1.90035;
    p->d1cb = 41.50283;
    p->d3cbi = 10.50589;
    p->acb = 0.67621;
    p->a3s = 0.2019100;
    p->b3s = -0.6068400;
    p->cphi = 11.8809900;
    p->fch3 = 0.0150000;
    p->fkinf = 0.3060000;
    p->bk = 50.7132;
    p->aa1 = 0.173746;
    p->aa3 = 2.166595;
#endif
#ifdef CBE_CH4CN
    /* BLOCK DATA PTPACM_ch4cn (egrad_ch4cn.f:2074-2114): the entries that differ from CH4 + OH */
    p->d3ch = 14.65328;
    p->a1ch = 1.75000;
    p->b1ch = 0.12000;
    p->c1ch = 5.00000;
    p->r0hh = 1.06497;
    p->d1hh = 132.17000;
    p->d3hh = 44.63017;
    p->ahh = 1.70000;
    p->r0cb = 1.72592;
    p->d3cbi = 88.69509;
    p->acb = 3.08688;
    p->aa1 = 0.273746;
    p->fkh2oeq = 0.2600000;
    p->alph2o = 3.1080000;
    p->anh2oeq = 180.0000000;
#endif
    p->d3cbi = p->d3cbi * fact1;
    p->a3cb = p->a3cb * fact1;
    p->fkh2oeq = p->fkh2oeq * fact2;  /* egrad_ch4oh.f:2001, egrad_geh4oh.f:1937 */
    p->anh2oeq = p->anh2oeq * fact3;
#else
    p->r0ch = 1.08898;
    p->d1ch = 111.266;
    p->d3ch = 48.96226;
    p->a1ch = 1.78374;
    p->b1ch = 0.14201;
    p->c1ch = 2.21773;
    p->r0hh = 0.74239;
    p->d1hh = 108.382;
    p->d3hh = 38.42657;
    p->ahh = 1.9589;
    p->r0cb = 1.08898;
    p->d1cb = 56.505;
    p->d3cb = 19.612;
    p->acb = 1.4173200;
    p->a3s = 0.1475300;
    p->b3s = -2.9926300;
    p->aphi = 0.5307000;
    p->bphi = 0.4012200;
    p->cphi = 1.9235100;
    p->atheta = 0.9119800;
    p->btheta = 0.3537500;
    p->ctheta = 1.8970500;
    p->fch3 = 0.0693700;
    p->hch3 = 0.1387400;
    p->fkinf = 0.4291400;
    p->ak = 0.1353000;
    p->bk = 10.7132;
    p->aa1 = 1.265960;
    p->aa2 = 0.000710;
    p->aa3 = 0.985920;
    p->aa4 = 2.785060;
#endif
    for (ind = 1; ind <= 3; ind++) {
        icount = ind - 3;
        p->nc[ind] = 3 * nnc + icount;
        p->nhb[ind] = 3 * nnb + icount;
        for (i = 1; i <= 4; i++) p->nh[i][ind] = 3 * nnh[i] + icount;
    }
    p->d1ch = p->d1ch * fact1;
    p->d3ch = p->d3ch * fact1;
    p->d1cb = p->d1cb * fact1;
    p->d3cb = p->d3cb * fact1;
    p->d1hh = p->d1hh * fact1;
    p->d3hh = p->d3hh * fact1;
    p->fch3 = p->fch3 * fact2;
    p->hch3 = p->hch3 * fact2;
    p->fkinf = p->fkinf * fact2;
    p->ak = p->ak * fact2;
}

/* ---- coorden_ch4h (:287-365) ---- */
static void ch4h_coorden(const ch4h_par *p, ch4h_state *s)
{
    int ind, i;
    for (ind = 1; ind <= 3; ind++) {
        s->tcb[ind] = s->q[p->nc[ind]] - s->q[p->nhb[ind]];
        for (i = 1; i <= 4; i++) {
            s->tch[i][ind] = s->q[p->nc[ind]] - s->q[p->nh[i][ind]];
            s->tbh[i][ind] = s->q[p->nhb[ind]] - s->q[p->nh[i][ind]];
        }
    }
    s->rcb = sqrt(s->tcb[1] * s->tcb[1] + s->tcb[2] * s->tcb[2] + s->tcb[3] * s->tcb[3]);
#ifdef CBE_CH4OH
    for (ind = 1; ind <= 3; ind++) s->tno[ind] = s->q[p->no[ind]] - s->q[p->nhb[ind]]; /* :351 */
    s->rno = sqrt(s->tno[1] * s->tno[1] + s->tno[2] * s->tno[2] + s->tno[3] * s->tno[3]); /* :361 */
#endif
    for (i = 1; i <= 4; i++) {
        s->rch[i] = sqrt(s->tch[i][1] * s->tch[i][1] + s->tch[i][2] * s->tch[i][2] +
                         s->tch[i][3] * s->tch[i][3]);
        s->rbh[i] = sqrt(s->tbh[i][1] * s->tbh[i][1] + s->tbh[i][2] * s->tbh[i][2] +
                         s->tbh[i][3] * s->tbh[i][3]);
    }
}

static real ipow(real x, int n)
{
    /* gfortran expands x**n (small integer n) by repeated squaring/multiplication */
    real r = 1.0, b = x;
    while (n > 0) {
        if (n & 1) r = r * b;
        n >>= 1;
        if (n) b = b * b;
    }
    return r;
}

/* ---- switchf_ch4h (:1545-1685) ---- */
static void ch4h_switchf(const ch4h_par *p, ch4h_state *s)
{
    const real argmax = 19.0;
    int i;
#ifdef CBE_CH4OH
    s->a1s = 1.5313681e-7; /* egrad_ch4oh.f:1808-1811 */
    s->b1s = -4.6696246;
    s->a2s = 1.0147402e-7;
    s->b2s = -12.363798;
#else
    s->a1s = 1.5132681e-7;
    s->b1s = -4.3792246;
    s->a2s = 1.9202402e-7;
    s->b2s = -12.323018;
#endif
    for (i = 1; i <= 4; i++) {
        real rch = s->rch[i];
        real args1, args2, args3;
        args1 = s->a1s * (rch - p->r0ch) * ipow(rch - s->b1s, 8);
        if (args1 < argmax) {
            s->s1[i] = 1.0 - tanh(args1);
            s->ds1[i] = s->a1s * (ipow(rch - s->b1s, 8) +
                                  8.0 * (rch - p->r0ch) * ipow(rch - s->b1s, 7));
            s->ds1[i] = -s->ds1[i] / ipow(cosh(args1), 2);
        } else {
            s->s1[i] = 0.0;
            s->ds1[i] = 0.0;
        }
        args2 = s->a2s * (rch - p->r0ch) * ipow(rch - s->b2s, 6);
        if (args2 < argmax) {
            s->s2[i] = 1.0 - tanh(args2);
            s->ds2[i] = s->a2s * (ipow(rch - s->b2s, 6) +
                                  6.0 * (rch - p->r0ch) * ipow(rch - s->b2s, 5));
            s->ds2[i] = -s->ds2[i] / ipow(cosh(args2), 2);
        } else {
            s->s2[i] = 0.0;
            s->ds2[i] = 0.0;
        }
        args3 = p->a3s * (rch - p->r0ch) * ipow(rch - p->b3s, 2);
        if (args3 < argmax) {
            s->s3[i] = 1.0 - tanh(args3);
            s->ds3[i] = p->a3s * (3.0 * ipow(rch, 2) - 2.0 * rch * (p->r0ch + 2.0 * p->b3s) +
                                  p->b3s * (p->b3s + 2.0 * p->r0ch));
            s->ds3[i] = -s->ds3[i] / ipow(cosh(args3), 2);
        } else {
            s->s3[i] = 0.0;
            s->ds3[i] = 0.0;
        }
#ifdef CBE_GEH4OH
        if (1) {   /* the rch < 3.8 test is commented out for sphi (egrad_geh4oh.f:1793-1804), kept for stheta */
#else
        if (rch < 3.8) {
#endif
            real argsphi = p->aphi * (rch - p->r0ch) * exp(p->bphi * ipow(rch - p->cphi, 3));
            s->sphi[i] = 1.0 - tanh(argsphi);
            s->dsphi[i] =
                p->aphi * (1.0 + 3.0 * p->bphi * (rch - p->r0ch) * ipow(rch - p->cphi, 2));
            s->dsphi[i] = s->dsphi[i] * exp(p->bphi * ipow(rch - p->cphi, 3));
            s->dsphi[i] = -s->dsphi[i] / ipow(cosh(argsphi), 2);
        } else {
            s->sphi[i] = 0.0;
            s->dsphi[i] = 0.0;
        }
        if (rch < 3.8) {
            real argstheta =
                p->atheta * (rch - p->r0ch) * exp(p->btheta * ipow(rch - p->ctheta, 3));
            s->stheta[i] = 1.0 - tanh(argstheta);
            s->dstheta[i] = p->atheta * (1.0 + 3.0 * p->btheta * (rch - p->r0ch) *
                                                   ipow(rch - p->ctheta, 2));
            s->dstheta[i] = s->dstheta[i] * exp(p->btheta * ipow(rch - p->ctheta, 3));
            s->dstheta[i] = -s->dstheta[i] / ipow(cosh(argstheta), 2);
        } else {
            s->stheta[i] = 0.0;
            s->dstheta[i] = 0.0;
        }
    }
}

/* ---- refangles_ch4h (:367-504) ---- */
static void ch4h_refangles(ch4h_state *s)
{
    const real pi = CH4H_PI;
    real tau = acos((real)(-1.0 / 3.0));
    real halfpi = 0.5 * pi;
    real twopi = 2.0 * pi;
#ifdef CBE_GEH4OH
    real ta = tau - 0.678 * pi;    /* (tau-taugeh), taugeh=0.678d0*pi: pyramidal GeH3 (egrad_geh4oh.f:368,:382-440) */
#else
    real ta = tau - halfpi;        /* (tau-halfpi)        */
#endif
    real tb = tau - twopi / 3.0;   /* (tau-twopi/3.0d0)   */
    real *sphi = s->sphi, *dsphi = s->dsphi, *stheta = s->stheta, *dstheta = s->dstheta;
    int i, j, k;
    for (i = 1; i <= 4; i++) {
        s->theta0[i][i] = 0.0;
        for (k = 1; k <= 4; k++) s->dtheta0[i][i][k] = 0.0;
    }
    s->theta0[1][2] = tau + ta * (sphi[1] * sphi[2] - 1.0) + tb * (stheta[3] * stheta[4] - 1.0);
    s->theta0[1][3] = tau + ta * (sphi[1] * sphi[3] - 1.0) + tb * (stheta[2] * stheta[4] - 1.0);
    s->theta0[1][4] = tau + ta * (sphi[1] * sphi[4] - 1.0) + tb * (stheta[2] * stheta[3] - 1.0);
    s->theta0[2][3] = tau + ta * (sphi[2] * sphi[3] - 1.0) + tb * (stheta[1] * stheta[4] - 1.0);
    s->theta0[2][4] = tau + ta * (sphi[2] * sphi[4] - 1.0) + tb * (stheta[1] * stheta[3] - 1.0);
    s->theta0[3][4] = tau + ta * (sphi[3] * sphi[4] - 1.0) + tb * (stheta[1] * stheta[2] - 1.0);
    /* wrt rch(1) */
    s->dtheta0[1][2][1] = ta * dsphi[1] * sphi[2];
    s->dtheta0[1][3][1] = ta * dsphi[1] * sphi[3];
    s->dtheta0[1][4][1] = ta * dsphi[1] * sphi[4];
    s->dtheta0[2][3][1] = tb * dstheta[1] * stheta[4];
    s->dtheta0[2][4][1] = tb * dstheta[1] * stheta[3];
    s->dtheta0[3][4][1] = tb * dstheta[1] * stheta[2];
    /* wrt rch(2) */
    s->dtheta0[1][2][2] = ta * sphi[1] * dsphi[2];
    s->dtheta0[1][3][2] = tb * dstheta[2] * stheta[4];
    s->dtheta0[1][4][2] = tb * dstheta[2] * stheta[3];
    s->dtheta0[2][3][2] = ta * dsphi[2] * sphi[3];
    s->dtheta0[2][4][2] = ta * dsphi[2] * sphi[4];
    s->dtheta0[3][4][2] = tb * stheta[1] * dstheta[2];
    /* wrt rch(3) */
    s->dtheta0[1][2][3] = tb * dstheta[3] * stheta[4];
    s->dtheta0[1][3][3] = ta * sphi[1] * dsphi[3];
    s->dtheta0[1][4][3] = tb * stheta[2] * dstheta[3];
    s->dtheta0[2][3][3] = ta * sphi[2] * dsphi[3];
    s->dtheta0[2][4][3] = tb * stheta[1] * dstheta[3];
    s->dtheta0[3][4][3] = ta * dsphi[3] * sphi[4];
    /* wrt rch(4) */
    s->dtheta0[1][2][4] = tb * stheta[3] * dstheta[4];
    s->dtheta0[1][3][4] = tb * stheta[2] * dstheta[4];
    s->dtheta0[1][4][4] = ta * sphi[1] * dsphi[4];
    s->dtheta0[2][3][4] = tb * stheta[1] * dstheta[4];
    s->dtheta0[2][4][4] = ta * sphi[2] * dsphi[4];
    s->dtheta0[3][4][4] = ta * sphi[3] * dsphi[4];
    for (i = 1; i <= 3; i++)
        for (j = i + 1; j <= 4; j++) {
            s->theta0[j][i] = s->theta0[i][j];
            for (k = 1; k <= 4; k++) s->dtheta0[j][i][k] = s->dtheta0[i][j][k];
        }
}

/* ---- stretch_ch4h (:506-711) ---- */
static void ch4h_stretch(const ch4h_par *p, ch4h_state *s, real *vstr_out)
{
    real vqch[5], vjch[5], vqbh[5], vjbh[5], vq[5], vj[5], achdc[4], achdh[5][4];
    real rav, vstr, arga, ach, dumach, e1, e3, vqcb, vjcb, dumqcb;
    const double r0ch = p->r0ch, r0cb = p->r0cb, r0hh = p->r0hh, acb = p->acb, ahh = p->ahh,
                 d1cb = p->d1cb, d1ch = p->d1ch, d3ch = p->d3ch, d1hh = p->d1hh, d3hh = p->d3hh;
#ifdef CBE_CH4OH
    real d3cb, dd3cb, texp, dt, expterm, vno, deddt, de, ded[4], addd3;
#else
    const double d3cb = p->d3cb;
#endif
    real *rch = s->rch, *rbh = s->rbh, rcb = s->rcb, *pdot = s->pdot;
    const int *nc = p->nc, *nhb = p->nhb;
    int i, ind, j, k;
    rav = (rch[1] + rch[2] + rch[3] + rch[4]) / 4.0;
    vstr = 0.0;
#ifdef CBE_CH4OH
    /* d3cb and dd3cb (:590-592); x**4.d0 and x**3.d0 are real powers of a possibly negative base */
    texp = exp(-pow(4.0 * (rav - p->rcbsp) / p->b3cb, 4.0));
    d3cb = (p->d3cbi - p->a3cb) + p->a3cb * texp;
    dd3cb = -4.0 * p->a3cb * texp * pow(rav - p->rcbsp, 3.0) * pow(4.0 / p->b3cb, 4.0);
#endif
    arga = p->c1ch * (rav - r0ch);
    if (arga < 19.0) {
        ach = p->a1ch + p->b1ch * (tanh(arga) + 1.0) * 0.5;
        dumach = p->b1ch * p->c1ch / (2.0 * ipow(cosh(arga), 2));
    } else {
        ach = p->a1ch + p->b1ch;
        dumach = 0.0;
    }
    e1 = d1cb * (exp(-2.0 * acb * (rcb - r0cb)) - 2.0 * exp(-acb * (rcb - r0cb)));
    e3 = d3cb * (exp(-2.0 * acb * (rcb - r0cb)) + 2.0 * exp(-acb * (rcb - r0cb)));
    vqcb = (e1 + e3) * 0.5;
    vjcb = (e1 - e3) * 0.5;
#ifdef CBE_CH4OH
#ifdef CBE_CH4CN
    /* C-N Morse term on literal constants (egrad_ch4cn.f:625-627) */
    dt = (s->rno - 1.172);
    expterm = exp(-0.80 * dt);
    vno = 80.0 * ((1.0 - expterm) * (1.0 - expterm));
#else
    /* O-H Morse term (:616-621) */
    dt = (s->rno - r0hh);
    expterm = exp(-ahh * dt);
    vno = d1hh * ((1.0 - expterm) * (1.0 - expterm));
#endif
#endif
    for (i = 1; i <= 4; i++) {
        e1 = d1ch * (exp(-2.0 * ach * (rch[i] - r0ch)) - 2.0 * exp(-ach * (rch[i] - r0ch)));
        e3 = d3ch * (exp(-2.0 * ach * (rch[i] - r0ch)) + 2.0 * exp(-ach * (rch[i] - r0ch)));
        vqch[i] = (e1 + e3) * 0.5;
        vjch[i] = (e1 - e3) * 0.5;
        e1 = d1hh * (exp(-2.0 * ahh * (rbh[i] - r0hh)) - 2.0 * exp(-ahh * (rbh[i] - r0hh)));
        e3 = d3hh * (exp(-2.0 * ahh * (rbh[i] - r0hh)) + 2.0 * exp(-ahh * (rbh[i] - r0hh)));
        vqbh[i] = (e1 + e3) * 0.5;
        vjbh[i] = (e1 - e3) * 0.5;
        vq[i] = vqch[i] + vqcb + vqbh[i];
        vj[i] = -sqrt((ipow(vjch[i] - vjcb, 2) + ipow(vjcb - vjbh[i], 2) +
                       ipow(vjbh[i] - vjch[i], 2)) *
                      0.5);
        vstr = vstr + vq[i] + vj[i];
    }
#ifdef CBE_CH4OH
    vstr = vstr + vno;                                      /* :646 */
#ifdef CBE_CH4CN
    deddt = 2.0 * 0.8 * 80.0 * (1.0 - expterm) * expterm;   /* egrad_ch4cn.f:659-660 */
#else
    deddt = 2.0 * ahh * d1hh * (1.0 - expterm) * expterm;   /* :652-659 */
#endif
    de = deddt / s->rno;
    for (i = 1; i <= 3; i++) ded[i] = de * s->tno[i];
#endif
    for (ind = 1; ind <= 3; ind++) {
        achdc[ind] = dumach *
                     (s->tch[1][ind] / rch[1] + s->tch[2][ind] / rch[2] + s->tch[3][ind] / rch[3] +
                      s->tch[4][ind] / rch[4]) /
                     4.0;
        for (i = 1; i <= 4; i++) achdh[i][ind] = -dumach * s->tch[i][ind] / rch[i] / 4.0;
    }
    dumqcb = -acb *
             ((d1cb + d3cb) * exp(-2.0 * acb * (rcb - r0cb)) -
              (d1cb - d3cb) * exp(-acb * (rcb - r0cb))) /
             rcb;
    for (i = 1; i <= 4; i++) {
        real dumqbh, factj, dumjcb, dumjbh;
        dumqbh = -ahh *
                 ((d1hh + d3hh) * exp(-2.0 * ahh * (rbh[i] - r0hh)) -
                  (d1hh - d3hh) * exp(-ahh * (rbh[i] - r0hh))) /
                 rbh[i];
#ifdef CBE_CH4OH
        /* "adding the derv of D3cb wrt r" as written (:681-686) */
        addd3 = dd3cb * (exp(-2.0 * acb * (rcb - r0cb)) + 2.0 * exp(-acb * (rcb - r0cb)));
        addd3 = addd3 / 4.0 * 0.5 / rcb;
        dumqbh = dumqbh + addd3;
#endif
        factj = 0.5 / vj[i];
        dumjcb = -acb *
                 ((d1cb - d3cb) * exp(-2.0 * acb * (rcb - r0cb)) -
                  (d1cb + d3cb) * exp(-acb * (rcb - r0cb))) *
                 factj / rcb;
        dumjbh = -ahh *
                 ((d1hh - d3hh) * exp(-2.0 * ahh * (rbh[i] - r0hh)) -
                  (d1hh + d3hh) * exp(-ahh * (rbh[i] - r0hh))) *
                 factj / rbh[i];
#ifdef CBE_CH4OH
        addd3 = dd3cb * (exp(-2.0 * acb * (rcb - r0cb)) + 2.0 * exp(-acb * (rcb - r0cb))); /* :694-699 */
        addd3 = addd3 / 4.0 * 0.5 * factj / rcb;
        dumjbh = dumjbh - addd3;
#endif
        for (ind = 1; ind <= 3; ind++) {
            real dumqch, dumqhi, dumjch, dumjhi;
            real tcb = s->tcb[ind], tbh = s->tbh[i][ind], tch = s->tch[i][ind];
            /* deriv wrt hb */
            pdot[nhb[ind]] = pdot[nhb[ind]] - tcb * dumqcb + tbh * dumqbh +
                             (vjch[i] - vjcb) * (dumjcb * tcb) +
                             (vjcb - vjbh[i]) * (-dumjcb * tcb - dumjbh * tbh) +
                             (vjbh[i] - vjch[i]) * dumjbh * tbh;
            /* dvqch(i)/dc */
            dumqch = -(ach * tch / rch[i] + achdc[ind] * (rch[i] - r0ch)) *
                     ((d1ch + d3ch) * exp(-2.0 * ach * (rch[i] - r0ch)) -
                      (d1ch - d3ch) * exp(-ach * (rch[i] - r0ch)));
            pdot[nc[ind]] = pdot[nc[ind]] + dumqch + tcb * dumqcb;
            /* dvqch(i)/dh(i) */
            dumqhi = (ach * tch / rch[i] - achdh[i][ind] * (rch[i] - r0ch)) *
                     ((d1ch + d3ch) * exp(-2.0 * ach * (rch[i] - r0ch)) -
                      (d1ch - d3ch) * exp(-ach * (rch[i] - r0ch)));
            pdot[p->nh[i][ind]] = pdot[p->nh[i][ind]] + dumqhi - tbh * dumqbh;
            /* dvjch(i)/dc */
            dumjch = -(ach * tch / rch[i] + achdc[ind] * (rch[i] - r0ch)) *
                     ((d1ch - d3ch) * exp(-2.0 * ach * (rch[i] - r0ch)) -
                      (d1ch + d3ch) * exp(-ach * (rch[i] - r0ch))) *
                     factj;
            pdot[nc[ind]] = pdot[nc[ind]] + (vjch[i] - vjcb) * (dumjch - dumjcb * tcb) +
                            (vjcb - vjbh[i]) * dumjcb * tcb - (vjbh[i] - vjch[i]) * dumjch;
            /* dvjch(i)/dh(i) */
            dumjhi = (ach * tch / rch[i] - achdh[i][ind] * (rch[i] - r0ch)) *
                     ((d1ch - d3ch) * exp(-2.0 * ach * (rch[i] - r0ch)) -
                      (d1ch + d3ch) * exp(-ach * (rch[i] - r0ch))) *
                     factj;
            pdot[p->nh[i][ind]] = pdot[p->nh[i][ind]] + (vjch[i] - vjcb) * dumjhi +
                                  (vjcb - vjbh[i]) * dumjbh * tbh +
                                  (vjbh[i] - vjch[i]) * (-dumjbh * tbh - dumjhi);
            /* dv(i)/dh(j) */
            for (k = 1; k <= 3; k++) {
                real dumqhj, dumjhj;
                j = i + k;
                if (j > 4) j = j - 4;
                dumqhj = -achdh[j][ind] * (rch[i] - r0ch) *
                         ((d1ch + d3ch) * exp(-2.0 * ach * (rch[i] - r0ch)) -
                          (d1ch - d3ch) * exp(-ach * (rch[i] - r0ch)));
                dumjhj = -achdh[j][ind] * (rch[i] - r0ch) *
                         ((d1ch - d3ch) * exp(-2.0 * ach * (rch[i] - r0ch)) -
                          (d1ch + d3ch) * exp(-ach * (rch[i] - r0ch))) *
                         factj;
                pdot[p->nh[j][ind]] = pdot[p->nh[j][ind]] + dumqhj +
                                      (vjch[i] - vjcb) * dumjhj - (vjbh[i] - vjch[i]) * dumjhj;
            }
        }
    }
#ifdef CBE_CH4OH
    for (ind = 1; ind <= 3; ind++) {                         /* :769-774 */
        pdot[nhb[ind]] = pdot[nhb[ind]] - ded[ind];
        pdot[p->no[ind]] = pdot[p->no[ind]] + ded[ind];
    }
#endif
    *vstr_out = vstr;
}

/* ---- opforce_ch4h (:1240-1348) ---- */
static void ch4h_opforce(const ch4h_par *p, ch4h_state *s)
{
    real sw[5], dsw[5][5];
    real *s3 = s->s3, *ds3 = s->ds3;
    int i, j;
    sw[1] = (1.0 - s3[1]) * s3[2] * s3[3] * s3[4];
    sw[2] = (1.0 - s3[2]) * s3[3] * s3[4] * s3[1];
    sw[3] = (1.0 - s3[3]) * s3[4] * s3[1] * s3[2];
    sw[4] = (1.0 - s3[4]) * s3[1] * s3[2] * s3[3];
    dsw[1][1] = -ds3[1] * s3[2] * s3[3] * s3[4];
    dsw[1][2] = (1.0 - s3[1]) * ds3[2] * s3[3] * s3[4];
    dsw[1][3] = (1.0 - s3[1]) * s3[2] * ds3[3] * s3[4];
    dsw[1][4] = (1.0 - s3[1]) * s3[2] * s3[3] * ds3[4];
    dsw[2][1] = (1.0 - s3[2]) * s3[3] * s3[4] * ds3[1];
    dsw[2][2] = -ds3[2] * s3[3] * s3[4] * s3[1];
    dsw[2][3] = (1.0 - s3[2]) * ds3[3] * s3[4] * s3[1];
    dsw[2][4] = (1.0 - s3[2]) * s3[3] * ds3[4] * s3[1];
    dsw[3][1] = (1.0 - s3[3]) * s3[4] * ds3[1] * s3[2];
    dsw[3][2] = (1.0 - s3[3]) * s3[4] * s3[1] * ds3[2];
    dsw[3][3] = -ds3[3] * s3[4] * s3[1] * s3[2];
    dsw[3][4] = (1.0 - s3[3]) * ds3[4] * s3[1] * s3[2];
    dsw[4][1] = (1.0 - s3[4]) * ds3[1] * s3[2] * s3[3];
    dsw[4][2] = (1.0 - s3[4]) * s3[1] * ds3[2] * s3[3];
    dsw[4][3] = (1.0 - s3[4]) * s3[1] * s3[2] * ds3[3];
    dsw[4][4] = -ds3[4] * s3[1] * s3[2] * s3[3];
    for (i = 1; i <= 4; i++) {
        s->fdelta[i] = sw[i] * p->fch3;
        s->hdelta[i] = sw[i] * p->hch3;
        for (j = 1; j <= 4; j++) {
            s->dfdelta[i][j] = dsw[i][j] * p->fch3;
            s->dhdelta[i][j] = dsw[i][j] * p->hch3;
        }
    }
}

/* ---- calcdelta_ch4h (:986-1238) ---- */
static void ch4h_calcdelta(const ch4h_par *p, ch4h_state *s, int i, int j, int k, int l,
                           real *sum2_out, real *sum4_out)
{
    real delta[5], a[4], b[4], axb[4], c[5][4], argd[5], daxb[5][4][4], cdot[5][4][4],
        atemp2[4];
    real norma, sum2, sum4, deldot, atemp1, atemp3, atemp4, atemp5;
    int in[4], ii, ind, jj;
    real *q = s->q, *pdot = s->pdot, *rch = s->rch;
    sum2 = 0.0;
    sum4 = 0.0;
    in[1] = j;
    in[2] = k;
    in[3] = l;
    for (ind = 1; ind <= 3; ind++) {
        a[ind] = q[p->nh[k][ind]] - q[p->nh[j][ind]];
        b[ind] = q[p->nh[l][ind]] - q[p->nh[j][ind]];
    }
    axb[1] = a[2] * b[3] - a[3] * b[2];
    axb[2] = a[3] * b[1] - a[1] * b[3];
    axb[3] = a[1] * b[2] - a[2] * b[1];
    norma = axb[1] * axb[1] + axb[2] * axb[2] + axb[3] * axb[3];
    norma = sqrt(norma);
    for (ii = 1; ii <= 3; ii++)
        for (ind = 1; ind <= 3; ind++) c[in[ii]][ind] = -s->tch[in[ii]][ind] / rch[in[ii]];
    for (ii = 1; ii <= 3; ii++) {
        argd[in[ii]] = axb[1] * c[in[ii]][1] + axb[2] * c[in[ii]][2] + axb[3] * c[in[ii]][3];
        argd[in[ii]] = argd[in[ii]] / norma;
        delta[in[ii]] = acos(argd[in[ii]]) - s->theta0[i][in[ii]];
        sum2 = sum2 + ipow(delta[in[ii]], 2);
        sum4 = sum4 + ipow(delta[in[ii]], 4);
    }
    /* derivatives of axb wrt hj */
    daxb[j][1][1] = 0.0;
    daxb[j][1][2] = b[3] - a[3];
    daxb[j][1][3] = -b[2] + a[2];
    daxb[j][2][1] = -b[3] + a[3];
    daxb[j][2][2] = 0.0;
    daxb[j][2][3] = b[1] - a[1];
    daxb[j][3][1] = b[2] - a[2];
    daxb[j][3][2] = -b[1] + a[1];
    daxb[j][3][3] = 0.0;
    /* wrt hk */
    daxb[k][1][1] = 0.0;
    daxb[k][1][2] = -b[3];
    daxb[k][1][3] = b[2];
    daxb[k][2][1] = b[3];
    daxb[k][2][2] = 0.0;
    daxb[k][2][3] = -b[1];
    daxb[k][3][1] = -b[2];
    daxb[k][3][2] = b[1];
    daxb[k][3][3] = 0.0;
    /* wrt hl */
    daxb[l][1][1] = 0.0;
    daxb[l][1][2] = a[3];
    daxb[l][1][3] = -a[2];
    daxb[l][2][1] = -a[3];
    daxb[l][2][2] = 0.0;
    daxb[l][2][3] = a[1];
    daxb[l][3][1] = a[2];
    daxb[l][3][2] = -a[1];
    daxb[l][3][3] = 0.0;
    for (ii = 1; ii <= 3; ii++) {
        int m = in[ii];
        real r2 = ipow(rch[m], 2);
        cdot[m][1][1] = 1.0 / rch[m] + s->tch[m][1] * c[m][1] / r2;
        cdot[m][1][2] = s->tch[m][2] * c[m][1] / r2;
        cdot[m][1][3] = s->tch[m][3] * c[m][1] / r2;
        cdot[m][2][1] = s->tch[m][1] * c[m][2] / r2;
        cdot[m][2][2] = 1.0 / rch[m] + s->tch[m][2] * c[m][2] / r2;
        cdot[m][2][3] = s->tch[m][3] * c[m][2] / r2;
        cdot[m][3][1] = s->tch[m][1] * c[m][3] / r2;
        cdot[m][3][2] = s->tch[m][2] * c[m][3] / r2;
        cdot[m][3][3] = 1.0 / rch[m] + s->tch[m][3] * c[m][3] / r2;
    }
    for (ii = 1; ii <= 3; ii++) {
        int mi = in[ii];
        for (ind = 1; ind <= 3; ind++) {
            deldot = -s->dtheta0[i][mi][i];
            deldot = -deldot * s->tch[i][ind] / rch[i];
            pdot[p->nh[i][ind]] = pdot[p->nh[i][ind]] + 2.0 * s->fdelta[i] * delta[mi] * deldot +
                                  4.0 * s->hdelta[i] * ipow(delta[mi], 3) * deldot;
            deldot = -deldot;
            pdot[p->nc[ind]] = pdot[p->nc[ind]] + 2.0 * s->fdelta[i] * delta[mi] * deldot +
                               4.0 * s->hdelta[i] * ipow(delta[mi], 3) * deldot;
            for (jj = 1; jj <= 3; jj++) {
                int mj = in[jj];
                atemp1 = axb[1] * daxb[mj][ind][1] + axb[2] * daxb[mj][ind][2] +
                         axb[3] * daxb[mj][ind][3];
                atemp1 = atemp1 / ipow(norma, 3);
                atemp2[1] = daxb[mj][ind][1] / norma - atemp1 * axb[1];
                atemp2[2] = daxb[mj][ind][2] / norma - atemp1 * axb[2];
                atemp2[3] = daxb[mj][ind][3] / norma - atemp1 * axb[3];
                atemp3 = atemp2[1] * c[mi][1] + atemp2[2] * c[mi][2] + atemp2[3] * c[mi][3];
                atemp4 = 0.0;
                if (ii == jj) {
                    atemp4 = axb[1] * cdot[mi][1][ind] + axb[2] * cdot[mi][2][ind] +
                             axb[3] * cdot[mi][3][ind];
                    atemp4 = atemp4 / norma;
                }
                atemp5 = -s->dtheta0[i][mi][mj];
                atemp5 = -atemp5 * s->tch[mj][ind] / rch[mj];
                deldot = atemp3 + atemp4;
                deldot = -1.0 / sqrt(1.0 - ipow(argd[mi], 2)) * deldot;
                deldot = deldot + atemp5;
                pdot[p->nh[mj][ind]] = pdot[p->nh[mj][ind]] +
                                       2.0 * s->fdelta[i] * delta[mi] * deldot +
                                       4.0 * s->hdelta[i] * ipow(delta[mi], 3) * deldot;
                deldot = 1.0 / sqrt(1.0 - ipow(argd[mi], 2)) * atemp4;
                deldot = deldot - atemp5;
                pdot[p->nc[ind]] = pdot[p->nc[ind]] + 2.0 * s->fdelta[i] * delta[mi] * deldot +
                                   4.0 * s->hdelta[i] * ipow(delta[mi], 3) * deldot;
            }
        }
    }
    *sum2_out = sum2;
    *sum4_out = sum4;
}

/* ---- opbend_ch4h (:713-863) ---- */
static void ch4h_opbend(const ch4h_par *p, ch4h_state *s, real *vop_out)
{
    real sumd2[5], sumd4[5], a[4], b[4], axb[4], c[5][4], argd[5];
    real norma, vop, sum2, sum4, ddr;
    int in[4], i, j, k, l, ii, ind, itemp;
    real *q = s->q, *pdot = s->pdot, *rch = s->rch;
    vop = 0.0;
    ch4h_opforce(p, s);
    for (i = 1; i <= 4; i++) {
        j = i + 1;
        if (j > 4) j = j - 4;
        k = j + 1;
        if (k > 4) k = k - 4;
        l = k + 1;
        if (l > 4) l = l - 4;
        in[1] = j;
        in[2] = k;
        in[3] = l;
        for (ind = 1; ind <= 3; ind++) {
            a[ind] = q[p->nh[k][ind]] - q[p->nh[j][ind]];
            b[ind] = q[p->nh[l][ind]] - q[p->nh[j][ind]];
        }
        axb[1] = a[2] * b[3] - a[3] * b[2];
        axb[2] = a[3] * b[1] - a[1] * b[3];
        axb[3] = a[1] * b[2] - a[2] * b[1];
        norma = axb[1] * axb[1] + axb[2] * axb[2] + axb[3] * axb[3];
        norma = sqrt(norma);
        for (ii = 1; ii <= 3; ii++)
            for (ind = 1; ind <= 3; ind++) c[in[ii]][ind] = -s->tch[in[ii]][ind] / rch[in[ii]];
        /* right-handedness test: toggles the loop-local k,l using the stale in(:) (:821-833) */
        for (ii = 1; ii <= 3; ii++) {
            argd[in[ii]] = axb[1] * c[in[ii]][1] + axb[2] * c[in[ii]][2] + axb[3] * c[in[ii]][3];
            argd[in[ii]] = argd[in[ii]] / norma;
            if (argd[in[ii]] > 0.0) {
                itemp = k;
                k = l;
                l = itemp;
            }
        }
        ch4h_calcdelta(p, s, i, j, k, l, &sum2, &sum4);
        sumd2[i] = sum2;
        sumd4[i] = sum4;
        vop = vop + s->fdelta[i] * sumd2[i] + s->hdelta[i] * sumd4[i];
    }
    for (i = 1; i <= 4; i++)
        for (j = 1; j <= 4; j++) {
            ddr = s->dfdelta[i][j] * sumd2[i] + s->dhdelta[i][j] * sumd4[i];
            for (ind = 1; ind <= 3; ind++) {
                pdot[p->nh[j][ind]] = pdot[p->nh[j][ind]] - s->tch[j][ind] * ddr / rch[j];
                pdot[p->nc[ind]] = pdot[p->nc[ind]] + s->tch[j][ind] * ddr / rch[j];
            }
        }
    *vop_out = vop;
}

/* ---- ipforce_ch4h (:1350-1543) ---- */
static void ch4h_ipforce(const ch4h_par *p, ch4h_state *s)
{
    real dfk0[5][5][5], df1dc[5], df1dh[5];
    real f0, f2;
    real *s1 = s->s1, *ds1 = s->ds1, *s2 = s->s2, *ds2 = s->ds2, *f1 = s->f1;
    real *rch = s->rch, *rbh = s->rbh;
    int i;
    f0 = p->fkinf + p->ak;
    f2 = p->fkinf;
    s->fk0[1][2] = f0 + f0 * (s1[1] * s1[2] - 1.0) + (f0 - f2) * (s2[3] * s2[4] - 1.0);
    s->fk0[1][3] = f0 + f0 * (s1[1] * s1[3] - 1.0) + (f0 - f2) * (s2[2] * s2[4] - 1.0);
    s->fk0[1][4] = f0 + f0 * (s1[1] * s1[4] - 1.0) + (f0 - f2) * (s2[2] * s2[3] - 1.0);
    s->fk0[2][3] = f0 + f0 * (s1[2] * s1[3] - 1.0) + (f0 - f2) * (s2[1] * s2[4] - 1.0);
    s->fk0[2][4] = f0 + f0 * (s1[2] * s1[4] - 1.0) + (f0 - f2) * (s2[1] * s2[3] - 1.0);
    s->fk0[3][4] = f0 + f0 * (s1[3] * s1[4] - 1.0) + (f0 - f2) * (s2[1] * s2[2] - 1.0);
    dfk0[1][2][1] = f0 * ds1[1] * s1[2];
    dfk0[1][2][2] = f0 * s1[1] * ds1[2];
    dfk0[1][2][3] = (f0 - f2) * ds2[3] * s2[4];
    dfk0[1][2][4] = (f0 - f2) * s2[3] * ds2[4];
    dfk0[1][3][1] = f0 * ds1[1] * s1[3];
    dfk0[1][3][2] = (f0 - f2) * ds2[2] * s2[4];
    dfk0[1][3][3] = f0 * s1[1] * ds1[3];
    dfk0[1][3][4] = (f0 - f2) * s2[2] * ds2[4];
    dfk0[1][4][1] = f0 * ds1[1] * s1[4];
    dfk0[1][4][2] = (f0 - f2) * ds2[2] * s2[3];
    dfk0[1][4][3] = (f0 - f2) * s2[2] * ds2[3];
    dfk0[1][4][4] = f0 * s1[1] * ds1[4];
    dfk0[2][3][1] = (f0 - f2) * ds2[1] * s2[4];
    dfk0[2][3][2] = f0 * ds1[2] * s1[3];
    dfk0[2][3][3] = f0 * s1[2] * ds1[3];
    dfk0[2][3][4] = (f0 - f2) * s2[1] * ds2[4];
    dfk0[2][4][1] = (f0 - f2) * ds2[1] * s2[3];
    dfk0[2][4][2] = f0 * ds1[2] * s1[4];
    dfk0[2][4][3] = (f0 - f2) * s2[1] * ds2[3];
    dfk0[2][4][4] = f0 * s1[2] * ds1[4];
    dfk0[3][4][1] = (f0 - f2) * ds2[1] * s2[2];
    dfk0[3][4][2] = (f0 - f2) * s2[1] * ds2[2];
    dfk0[3][4][3] = f0 * ds1[3] * s1[4];
    dfk0[3][4][4] = f0 * s1[3] * ds1[4];
    for (i = 1; i <= 4; i++) {
        real arga1, arga2, a1, a2, duma1, duma2;
        arga1 = p->aa1 * rbh[i] * rbh[i];
        arga2 = p->aa4 * (rbh[i] - p->r0hh) * (rbh[i] - p->r0hh);
        a1 = 1.0 - exp(-arga1);
        a2 = p->aa2 + p->aa3 * exp(-arga2);
        f1[i] = a1 * exp(-a2 * ipow(rch[i] - p->r0ch, 2));
        duma1 = 2.0 * p->aa1 * rbh[i] * exp(-arga1);
        duma2 = -2.0 * p->aa3 * p->aa4 * (rbh[i] - p->r0hh) * exp(-arga2);
        df1dc[i] = -2.0 * (rch[i] - p->r0ch) * a1 * a2 * exp(-a2 * ipow(rch[i] - p->r0ch, 2));
        df1dh[i] = duma1 * exp(-a2 * ipow(rch[i] - p->r0ch, 2)) -
                   duma2 * ipow(rch[i] - p->r0ch, 2) * a1 * exp(-a2 * ipow(rch[i] - p->r0ch, 2));
    }
    s->dfdc[1][2][1] = dfk0[1][2][1] * f1[1] * f1[2] + s->fk0[1][2] * df1dc[1] * f1[2];
    s->dfdc[1][2][2] = dfk0[1][2][2] * f1[1] * f1[2] + s->fk0[1][2] * f1[1] * df1dc[2];
    s->dfdc[1][2][3] = dfk0[1][2][3] * f1[1] * f1[2];
    s->dfdc[1][2][4] = dfk0[1][2][4] * f1[1] * f1[2];
    s->dfdc[1][3][1] = dfk0[1][3][1] * f1[1] * f1[3] + s->fk0[1][3] * df1dc[1] * f1[3];
    s->dfdc[1][3][2] = dfk0[1][3][2] * f1[1] * f1[3];
    s->dfdc[1][3][3] = dfk0[1][3][3] * f1[1] * f1[3] + s->fk0[1][3] * f1[1] * df1dc[3];
    s->dfdc[1][3][4] = dfk0[1][3][4] * f1[1] * f1[3];
    s->dfdc[1][4][1] = dfk0[1][4][1] * f1[1] * f1[4] + s->fk0[1][4] * df1dc[1] * f1[4];
    s->dfdc[1][4][2] = dfk0[1][4][2] * f1[1] * f1[4];
    s->dfdc[1][4][3] = dfk0[1][4][3] * f1[1] * f1[4];
    s->dfdc[1][4][4] = dfk0[1][4][4] * f1[1] * f1[4] + s->fk0[1][4] * f1[1] * df1dc[4];
    s->dfdc[2][3][1] = dfk0[2][3][1] * f1[2] * f1[3];
    s->dfdc[2][3][2] = dfk0[2][3][2] * f1[2] * f1[3] + s->fk0[2][3] * df1dc[2] * f1[3];
    s->dfdc[2][3][3] = dfk0[2][3][3] * f1[2] * f1[3] + s->fk0[2][3] * f1[2] * df1dc[3];
    s->dfdc[2][3][4] = dfk0[2][3][4] * f1[2] * f1[3];
    s->dfdc[2][4][1] = dfk0[2][4][1] * f1[2] * f1[4];
    s->dfdc[2][4][2] = dfk0[2][4][2] * f1[2] * f1[4] + s->fk0[2][4] * df1dc[2] * f1[4];
    s->dfdc[2][4][3] = dfk0[2][4][3] * f1[2] * f1[4];
    s->dfdc[2][4][4] = dfk0[2][4][4] * f1[2] * f1[4] + s->fk0[2][4] * f1[2] * df1dc[4];
    s->dfdc[3][4][1] = dfk0[3][4][1] * f1[3] * f1[4];
    s->dfdc[3][4][2] = dfk0[3][4][2] * f1[3] * f1[4];
    s->dfdc[3][4][3] = dfk0[3][4][3] * f1[3] * f1[4] + s->fk0[3][4] * df1dc[3] * f1[4];
    s->dfdc[3][4][4] = dfk0[3][4][4] * f1[3] * f1[4] + s->fk0[3][4] * f1[3] * df1dc[4];
    s->dfdh[1][2][1] = s->fk0[1][2] * df1dh[1] * f1[2];
    s->dfdh[1][2][2] = s->fk0[1][2] * f1[1] * df1dh[2];
    s->dfdh[1][2][3] = 0.0;
    s->dfdh[1][2][4] = 0.0;
    s->dfdh[1][3][1] = s->fk0[1][3] * df1dh[1] * f1[3];
    s->dfdh[1][3][2] = 0.0;
    s->dfdh[1][3][3] = s->fk0[1][3] * f1[1] * df1dh[3];
    s->dfdh[1][3][4] = 0.0;
    s->dfdh[1][4][1] = s->fk0[1][4] * df1dh[1] * f1[4];
    s->dfdh[1][4][2] = 0.0;
    s->dfdh[1][4][3] = 0.0;
    s->dfdh[1][4][4] = s->fk0[1][4] * f1[1] * df1dh[4];
    s->dfdh[2][3][1] = 0.0;
    s->dfdh[2][3][2] = s->fk0[2][3] * df1dh[2] * f1[3];
    s->dfdh[2][3][3] = s->fk0[2][3] * f1[2] * df1dh[3];
    s->dfdh[2][3][4] = 0.0;
    s->dfdh[2][4][1] = 0.0;
    s->dfdh[2][4][2] = s->fk0[2][4] * df1dh[2] * f1[4];
    s->dfdh[2][4][3] = 0.0;
    s->dfdh[2][4][4] = s->fk0[2][4] * f1[2] * df1dh[4];
    s->dfdh[3][4][1] = 0.0;
    s->dfdh[3][4][2] = 0.0;
    s->dfdh[3][4][3] = s->fk0[3][4] * df1dh[3] * f1[4];
    s->dfdh[3][4][4] = s->fk0[3][4] * f1[3] * df1dh[4];
}

/* ---- ipbend_ch4h (:865-984) ---- */
static void ch4h_ipbend(const ch4h_par *p, ch4h_state *s, real *vip_out)
{
    real costh[5][5], theta[5][5], dth[5][5];
    real vip, termth, dthi, dthj, dthc, dth0k, dth0c;
    real *pdot = s->pdot, *rch = s->rch, *rbh = s->rbh, *f1 = s->f1;
    int i, j, k, ind;
    vip = 0.0;
    ch4h_ipforce(p, s);
    for (i = 1; i <= 3; i++)
        for (j = i + 1; j <= 4; j++) {
            costh[i][j] = s->tch[i][1] * s->tch[j][1] + s->tch[i][2] * s->tch[j][2] +
                          s->tch[i][3] * s->tch[j][3];
            costh[i][j] = costh[i][j] / rch[i] / rch[j];
            theta[i][j] = acos(costh[i][j]);
            dth[i][j] = theta[i][j] - s->theta0[i][j];
            vip = vip + 0.5 * s->fk0[i][j] * f1[i] * f1[j] * ipow(dth[i][j], 2);
            termth = -1.0 / sqrt(1.0 - costh[i][j] * costh[i][j]);
            for (ind = 1; ind <= 3; ind++) {
                dthi = -s->tch[j][ind] / rch[i] / rch[j] +
                       costh[i][j] * s->tch[i][ind] / rch[i] / rch[i];
                dthi = dthi * termth;
                dthj = -s->tch[i][ind] / rch[i] / rch[j] +
                       costh[i][j] * s->tch[j][ind] / rch[j] / rch[j];
                dthj = dthj * termth;
                dthc = -(dthi + dthj);
                pdot[p->nh[i][ind]] =
                    pdot[p->nh[i][ind]] + s->fk0[i][j] * f1[i] * f1[j] * dthi * dth[i][j];
                pdot[p->nh[j][ind]] =
                    pdot[p->nh[j][ind]] + s->fk0[i][j] * f1[i] * f1[j] * dthj * dth[i][j];
                pdot[p->nc[ind]] =
                    pdot[p->nc[ind]] + s->fk0[i][j] * f1[i] * f1[j] * dthc * dth[i][j];
                for (k = 1; k <= 4; k++) {
                    dth0k = -s->dtheta0[i][j][k] * s->tch[k][ind] / rch[k];
                    dth0c = -dth0k;
                    pdot[p->nh[k][ind]] =
                        pdot[p->nh[k][ind]] -
                        0.5 * s->tch[k][ind] * s->dfdc[i][j][k] * ipow(dth[i][j], 2) / rch[k] -
                        0.5 * s->tbh[k][ind] * s->dfdh[i][j][k] * ipow(dth[i][j], 2) / rbh[k] -
                        s->fk0[i][j] * f1[i] * f1[j] * dth0k * dth[i][j];
                    pdot[p->nc[ind]] =
                        pdot[p->nc[ind]] +
                        0.5 * s->tch[k][ind] * s->dfdc[i][j][k] * ipow(dth[i][j], 2) / rch[k] -
                        s->fk0[i][j] * f1[i] * f1[j] * dth0c * dth[i][j];
                    pdot[p->nhb[ind]] =
                        pdot[p->nhb[ind]] +
                        0.5 * s->tbh[k][ind] * s->dfdh[i][j][k] * ipow(dth[i][j], 2) / rbh[k];
                }
            }
        }
#ifdef CBE_CH4OH
    {   /* H(i)-O-H(O) bends, egrad_ch4oh.f:1059-1171 */
        real angh2o[5], fkh2o[5], dkdr[5], dkdx[5][4], dedo[4], dedhi[4], dedho[4];
        real dot, cosine, arga, dang, deno, term1, dstda, xp, yp, zp, rp, terma, termc;
        const real *tno = s->tno, rno = s->rno;
        for (i = 1; i <= 4; i++) {
            dot = 0.0;
            for (j = 1; j <= 3; j++) dot = dot - tno[j] * s->tbh[i][j];
            cosine = dot / (rno * rbh[i]);
            cosine = (cosine > -1.0) ? cosine : (real)(-1.0);   /* min(1, max(-1, cosine)) */
            cosine = (cosine < 1.0) ? cosine : (real)1.0;
            angh2o[i] = acos(cosine);
        }
        for (i = 1; i <= 4; i++) {
            arga = p->alph2o * (rbh[i] - p->r0hh);
            if (arga < 19.0)
                fkh2o[i] = p->fkh2oeq * (1 - tanh(arga));
            else
                fkh2o[i] = 0.0;
        }
        for (i = 1; i <= 4; i++) {
            dang = (angh2o[i] - p->anh2oeq);
            vip = vip + 0.5 * fkh2o[i] * dang * dang;
        }
        for (i = 1; i <= 4; i++) {
            deno = cosh(p->alph2o * (rbh[i] - p->r0hh));
            dkdr[i] = -(p->fkh2oeq * p->alph2o) / (deno * deno);
        }
        for (i = 1; i <= 4; i++) {
            dkdr[i] = dkdr[i] / rbh[i];
            for (j = 1; j <= 3; j++) dkdx[i][j] = dkdr[i] * s->tbh[i][j];
        }
        for (i = 1; i <= 4; i++)
            for (j = 1; j <= 3; j++) {
                term1 = 0.5 * ((angh2o[i] - p->anh2oeq) * (angh2o[i] - p->anh2oeq));
                pdot[p->nhb[j]] = pdot[p->nhb[j]] + dkdx[i][j] * term1;
                pdot[p->nh[i][j]] = pdot[p->nh[i][j]] - dkdx[i][j] * term1;
            }
        for (i = 1; i <= 4; i++) {
            dstda = fkh2o[i] * (angh2o[i] - p->anh2oeq);
            xp = tno[2] * s->tbh[i][3] - tno[3] * s->tbh[i][2];
            yp = tno[3] * s->tbh[i][1] - tno[1] * s->tbh[i][3];
            zp = tno[1] * s->tbh[i][2] - tno[2] * s->tbh[i][1];
            rp = sqrt(xp * xp + yp * yp + zp * zp);
            if (rp < 1.0e-6) rp = 1.0e-6;
            terma = dstda / (rbh[i] * rbh[i] * rp);
            termc = dstda / (rno * rno * rp);
            dedhi[1] = -terma * (s->tbh[i][2] * zp - s->tbh[i][3] * yp);
            dedhi[2] = -terma * (s->tbh[i][3] * xp - s->tbh[i][1] * zp);
            dedhi[3] = -terma * (s->tbh[i][1] * yp - s->tbh[i][2] * xp);
            dedho[1] = -termc * (tno[2] * zp - tno[3] * yp);
            dedho[2] = -termc * (tno[3] * xp - tno[1] * zp);
            dedho[3] = -termc * (tno[1] * yp - tno[2] * xp);
            dedo[1] = -dedhi[1] - dedho[1];
            dedo[2] = -dedhi[2] - dedho[2];
            dedo[3] = -dedhi[3] - dedho[3];
            for (j = 1; j <= 3; j++) {
                pdot[p->nhb[j]] = pdot[p->nhb[j]] + dedo[j];
                pdot[p->nh[i][j]] = pdot[p->nh[i][j]] + dedhi[j];
                pdot[p->no[j]] = pdot[p->no[j]] + dedho[j];
            }
        }
    }
#endif
    *vip_out = vip;
}

/* ---- POT_ch4h (:161-285): R(1..18) cartesians in bohr -> energy (hartree), DEGSDR ---- */
#undef CBE_NC
#undef CBE_NAT
#undef CBE_EGRAD
#undef CBE_PARTS
#undef CBE_PARTS_GRAD
#if defined(CBE_CH4CN)
#define CBE_NC 21   /* POT_ch4cn :161-293 */
#define CBE_NAT 7
#define CBE_EGRAD oracle_egrad_ch4cn_real
#define CBE_PARTS oracle_ch4cn_parts_real
#define CBE_PARTS_GRAD oracle_ch4cn_parts_grad_real
#elif defined(CBE_GEH4OH)
#define CBE_NC 21   /* pot_geh4oh :84-217 */
#define CBE_NAT 7
#define CBE_EGRAD oracle_egrad_geh4oh_real
#define CBE_PARTS oracle_geh4oh_parts_real
#define CBE_PARTS_GRAD oracle_geh4oh_parts_grad_real
#elif defined(CBE_CH4OH)
#define CBE_NC 21   /* POT_ch4oh :157-286 */
#define CBE_NAT 7
#define CBE_EGRAD oracle_egrad_ch4oh_real
#define CBE_PARTS oracle_ch4oh_parts_real
#define CBE_PARTS_GRAD oracle_ch4oh_parts_grad_real
#else
#define CBE_NC 18
#define CBE_NAT 6
#define CBE_EGRAD oracle_egrad_ch4h_real
#define CBE_PARTS oracle_ch4h_parts_real
#define CBE_PARTS_GRAD oracle_ch4h_parts_grad_real
#endif
static void ch4h_pot_g(const ch4h_par *p, const real R[CBE_NC + 1], real *en_out, real DEGSDR[CBE_NC + 1],
                       real parts[3], real *gparts);
static void ch4h_pot(const ch4h_par *p, const real R[CBE_NC + 1], real *en_out, real DEGSDR[CBE_NC + 1],
                     real parts[3])
{
    ch4h_pot_g(p, R, en_out, DEGSDR, parts, (real *)0);
}
/* gparts (may be null): [3][CBE_NC] = pdot after stretch, the increment of opbend, the increment of ipbend
 * (all three accumulate into the same pdot(150), egrad_ch4h.f:250-262), in the routine's own units per Angstrom */
static void ch4h_pot_g(const ch4h_par *p, const real R[CBE_NC + 1], real *en_out, real DEGSDR[CBE_NC + 1],
                       real parts[3], real *gparts)
{
    ch4h_state s;
    real vstr, vop, vip, en;
    int i;
    for (i = 1; i <= CBE_NC; i++) {
        s.q[i] = R[i] * 0.52918;
        s.pdot[i] = 0.0;
    }
    ch4h_coorden(p, &s);
    ch4h_switchf(p, &s);
    ch4h_refangles(&s);
    ch4h_stretch(p, &s, &vstr);
    if (gparts)
        for (i = 1; i <= CBE_NC; i++) gparts[i - 1] = s.pdot[i];
    ch4h_opbend(p, &s, &vop);
    if (gparts)
        for (i = 1; i <= CBE_NC; i++) gparts[CBE_NC + i - 1] = s.pdot[i] - gparts[i - 1];
    ch4h_ipbend(p, &s, &vip);
    if (gparts)
        for (i = 1; i <= CBE_NC; i++) gparts[2 * CBE_NC + i - 1] = s.pdot[i] - gparts[i - 1] - gparts[CBE_NC + i - 1];
    en = vstr + vop + vip;
    en = en * 0.03812;
    *en_out = en;
    for (i = 1; i <= CBE_NC; i++) DEGSDR[i] = s.pdot[i] * 0.0201723;
    if (parts) {
        parts[0] = vstr;
        parts[1] = vop;
        parts[2] = vip;
    }
}

/* ---- egrad_ch4h (:74-132); atom order H,C,H,H,H,H(b) (nnc=2, nnb=6, nnh=3,4,5,1) ---- */
void CBE_EGRAD(const real *q, int natoms, int nbeads, real *V, real *dVdq, int *info)
{
    ch4h_par par;
    int k, i, j;
    ch4h_prepot(&par);
    *info = 0;
    for (k = 0; k < nbeads; k++) {
        const real *qk = q + (long)k * 3 * natoms;
        real *gk = dVdq + (long)k * 3 * natoms;
        real R[CBE_NC + 1], D[CBE_NC + 1];
        for (j = 0; j < CBE_NAT; j++)
            for (i = 0; i < 3; i++) R[3 * j + i + 1] = qk[3 * j + i];
        ch4h_pot(&par, R, &V[k], D, (real *)0);
        for (j = 0; j < natoms; j++)
            for (i = 0; i < 3; i++) gk[3 * j + i] = (j < CBE_NAT) ? D[3 * j + i + 1] : (real)0.0;
    }
}

/* energy split (vstr, vop, vip in 1e5 J/mol) for the per-term tests */
void CBE_PARTS(const real *q18, real parts[3], real *V)
{
    ch4h_par par;
    real R[CBE_NC + 1], D[CBE_NC + 1];
    int i;
    ch4h_prepot(&par);
    for (i = 0; i < CBE_NC; i++) R[i + 1] = q18[i];
    ch4h_pot(&par, R, V, D, parts);
}

/* the same with the gradient of every part, d(part)/d(q in Angstrom), [3][CBE_NC] */
void CBE_PARTS_GRAD(const real *q18, real parts[3], real *gparts)
{
    ch4h_par par;
    real R[CBE_NC + 1], D[CBE_NC + 1], V;
    int i;
    ch4h_prepot(&par);
    for (i = 0; i < CBE_NC; i++) R[i + 1] = q18[i];
    ch4h_pot_g(&par, R, &V, D, parts, gparts);
}
