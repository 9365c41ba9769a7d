/*
 * pes_clnh3.c -- CPU oracle: NH3 + Cl -> NH2 + HCl surface (Monge-Palacios, Rangel, Corchado, Espinosa-Garcia,
 * Int. J. Quantum Chem. 112, 1887 (2012); POTLIB form), /root/reference/src/egrad_clnh3.f, and -- with -DCBE3_NH3OH,
 * set by pes_nh3oh.c which includes this file -- NH3 + OH -> NH2 + H2O (Monge-Palacios, Rangel, Espinosa-Garcia,
 * J. Chem. Phys. 138, 084305 (2013)), /root/reference/src/egrad_nh3oh.f.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_real.h).  Parity UNPINNED by the reference (no golden vectors, cannot be
 * compiled here); pinned by finite differences per term with the reference lengths frozen, by the asymptotes and by the
 * known answers of tests/test_oracle_clnh3.py.
 *
 * The three-hydrogen members of the family of pes_ch4h.c: a LEPS-type stretch over the three N-H / H-b bond pairs and a
 * harmonic in-plane bend over the three H-N-H angles; the out-of-plane term is switched off in both sources
 * ("c      call opbend(vop)", "en=vstr+vip": egrad_clnh3.f:172-182, egrad_nh3oh.f:233-243), so opbend / opforce /
 * calcdelta and the tables of oprefangles have no effect on the result and are not restated.
 *   egrad_clnh3 :36-67 (copies bead 1 only; every call site passes one bead, gradient.f90:187 -- the oracle loops over
 *   the images it is given), POT_clnh3 :82-210 (CARTOU / CARTTOR / EUNITZERO / RTOCART / DEDCOU of util_clnh3.f are
 *   identity maps for NFLAG(1) = NFLAG(2) = 1, ICARTR = 1, ANUZERO = 0, as in pes_ch4h.c), coorden :216-306,
 *   refangles :312-424, stretch :580-783, ipbend :941-1058, ipforce :1416-1546, switchf :1552-1688,
 *   initialize_clnh3 :1690-1797, BLOCK DATA :1799-1901.  Atom order H, N, H, H, Cl (nnc = 2, nnb = 5, nnh = 3, 4, 1).
 * Two properties of the source that the oracle keeps:
 *   - the equilibrium N-H length r0ch is a function of the geometry (:287-299, "jcc-2010": a tanh blend between the
 *     reactant and product values), but no derivative of it enters pdot: the returned gradient is NOT the gradient of the
 *     returned energy (the difference is what d r0ch / d rch would contribute);
 *   - fk0 is one number (fkinf + ak exp(-bk sum (rch - r0ch)^2), :1467-1476).
 * With CBE3_NH3OH (egrad_nh3oh.f: egrad :73-131, POT :160-311, coorden :313-426, refangles :428-549, stretch :706-929,
 * ipbend :1088-1251, ipforce :1615-1748, switchf :1750-1891, PREPOT :1893-2012, BLOCK DATA :2014-2131; atom order
 * H, N, H, H, O, H(O): nnc = 2, nnb = 5, nnh = 1, 3, 4, nno = 6):
 *   - the O-H length of the forming water r0hh is blended on the spectator O-H distance rno (:414-424), d3cb is switched
 *     on the mean N-H distance with a REAL power of a negative base (:752-753), an O-H Morse term (:766-768, :789) and
 *     three H-O-H bends with a tanh-switched force constant (:1213-1248) are added, fk0(i,j) is built from s1, s2
 *     (:1638-1640);
 *   - the GRADIENT IS NUMERIC: POT_nh3oh overwrites the analytic DEGSDR by forward differences of the energy with
 *     PASO = 1e-5 A on the 18 coordinates in turn (:283-296), each coordinate restored by q(I) = q(I) - PASO, i.e. left
 *     at (q + h) - h for the evaluations that follow; the analytic pdot code still runs (19 times) on undefined fko /
 *     argfk0 (:1638-1693 never set them) and is dead.  The oracle restates the energy path and the difference loop;
 *     one ulp of the energy (1.1e-16 Eh at 0.8 Eh) is 5.9e-12 Eh/bohr in this gradient.
 */
#include "oracle_real.h"
#include "oracle.h"

#define N3_PI 3.141592653589793

#undef N3_NC
#undef N3_NAT
#undef N3_EGRAD
#undef N3_PARTS
#undef N3_PARTS_GRAD
#undef N3
#ifdef CBE3_NH3OH
#define N3_NC 18
#define N3_NAT 6
#define N3_EGRAD oracle_egrad_nh3oh_real
#define N3_PARTS oracle_nh3oh_parts_real
#define N3_PARTS_GRAD oracle_nh3oh_parts_grad_real
#define N3(x) nh3oh_##x
#else
#define N3_NC 15
#define N3_NAT 5
#define N3_EGRAD oracle_egrad_clnh3_real
#define N3_PARTS oracle_clnh3_parts_real
#define N3_PARTS_GRAD oracle_clnh3_parts_grad_real
#define N3(x) clnh3_##x
#endif

typedef struct {
    /* /POTCM/ after the unit scaling of initialize_clnh3 / PREPOT_nh3oh */
    double r0chr, r0chp, w1, w2, d1ch, d3ch, a1ch, b1ch, c1ch, d1hh, d3hh, ahh, r0cb, d1cb, acb, a3s, b3s, aphi, bphi,
        cphi, atheta, btheta, ctheta, fkinf, ak, bk, aa1, aa2, aa3, aa4, tau, taunh2;
#ifdef CBE3_NH3OH
    double r0hhr, r0hhp, w3, w4, d3cbi, a3cb, b3cb, rcbsp, fkh2oeq, alph2o, angh2oeq;
    int no[4];
#else
    double r0hh, d3cb;
#endif
    int nc[4], nhb[4], nh[4][4]; /* /ndx/ */
} N3(par);

typedef struct {
    real theta0[4][4], dtheta0[4][4][4];                 /* /angles/  */
    real rcb, rch[4], rbh[4], r0ch, r0hh;                /* /bonds/   */
    real tcb[4], tch[4][4], tbh[4][4];                   /* /coords/  */
    real fk0[4][4], f1[4], dfdc[4][4][4], dfdh[4][4][4]; /* /force1/  */
    real s1[4], ds1[4], s2[4], ds2[4], s3[4], ds3[4];    /* /ip1/, /op1/ */
    real sphi[4], dsphi[4], stheta[4], dstheta[4];       /* /switch1/ */
    real q[N3_NC + 1], pdot[N3_NC + 1];                  /* /qpdot_pl/ */
#ifdef CBE3_NH3OH
    real rno, tno[4], angh2o[4], fkh2o[4];
#endif
} N3(state);

/* BLOCK DATA followed by the scaling of initialize_clnh3 (:1780-1793) / PREPOT_nh3oh (:1993-2009) */
static void N3(prepot)(N3(par) * p)
{
    const double fact1 = 0.041840, fact2 = 6.022045;
    int ind, i;
#ifdef CBE3_NH3OH
    const int nnc = 2, nnb = 5, nnh[4] = {0, 1, 3, 4}, nno = 6;
    const double fact3 = 2.0 * 3.1415926 / 360.0;
    p->r0chr = 1.01417;
    p->r0chp = 1.02777;
    p->w1 = 3.00000;
    p->w2 = 1.01417;
    p->d1ch = 125.250;
    p->d3ch = 24.300;
    p->a1ch = 2.050000;
    p->b1ch = -0.200000;
    p->c1ch = 200.4000;
    p->d1hh = 135.250;
    p->d3hh = 32.800;
    p->ahh = 2.0500;
    p->r0hhr = 0.9710;
    p->r0hhp = 0.9595;
    p->w3 = 1.00;
    p->w4 = 0.973;
    p->r0cb = 1.83800;
    p->d1cb = 80.900;
    p->d3cbi = 26.700;
    p->acb = 1.4800000;
    p->a3s = 0.2419100;
    p->b3s = -0.4068400;
    p->aphi = 2.2287900;
    p->bphi = 0.0206600;
    p->cphi = 1.5209900;
    p->atheta = 1.1578700;
    p->btheta = 0.0358900;
    p->ctheta = 0.7115500;
    p->fkinf = 0.7100000;
    p->ak = -0.0900000;
    p->bk = 2.7132000;
    p->aa1 = 0.800000;
    p->aa2 = 2.509960;
    p->aa3 = 3.506600;
    p->aa4 = 1.500000;
    p->tau = 1.9022600;
    p->taunh2 = 1.8046700;
    p->fkh2oeq = 0.7100000;
    p->alph2o = 0.7250;
    p->angh2oeq = 103.5968;
    p->a3cb = 1.60;
    p->b3cb = 0.011;
    p->rcbsp = 1.63349;
#else
    const int nnc = 2, nnb = 5, nnh[4] = {0, 3, 4, 1};
    p->r0chr = 1.01410;
    p->r0chp = 1.02700;
    p->w1 = 1.00000;
    p->w2 = 1.01400;
    p->d1ch = 119.058;
    p->d3ch = 20.000;
    p->a1ch = 2.125000;
    p->b1ch = -0.090000;
    p->c1ch = 22.00000;
    p->r0hh = 1.27730;
    p->d1hh = 109.850;
    p->d3hh = 18.400;
    p->ahh = 1.8600;
    p->r0cb = 2.10400;
    p->d1cb = 65.100;
    p->d3cb = 16.530;
    p->acb = 0.7780000;
    p->a3s = 1.0897000;
    p->b3s = -0.8088000;
    p->aphi = 6.7730500;
    p->bphi = 6.8000000;
    p->cphi = 1.9226100;
    p->atheta = 6.7359700;
    p->btheta = 6.7000000;
    p->ctheta = 1.9505500;
    p->fkinf = 0.6950000;
    p->ak = -0.0100000;
    p->bk = 0.1000100;
    p->aa1 = 3.503370;
    p->aa2 = 6.130490;
    p->aa3 = 6.100000;
    p->aa4 = 3.232430;
    p->tau = 1.9022600;
    p->taunh2 = 1.8046700;
#endif
    for (ind = 1; ind <= 3; ind++) {
        const int icount = ind - 3;
        p->nc[ind] = 3 * nnc + icount;
        p->nhb[ind] = 3 * nnb + icount;
#ifdef CBE3_NH3OH
        p->no[ind] = 3 * nno + icount;
#endif
        for (i = 1; i <= 3; i++) p->nh[i][ind] = 3 * nnh[i] + icount;
    }
    p->d1ch = p->d1ch * fact1;
    p->d3ch = p->d3ch * fact1;
    p->d1cb = p->d1cb * fact1;
#ifdef CBE3_NH3OH
    p->d3cbi = p->d3cbi * fact1;
    p->a3cb = p->a3cb * fact1;
#else
    p->d3cb = p->d3cb * fact1;
#endif
    p->d1hh = p->d1hh * fact1;
    p->d3hh = p->d3hh * fact1;
    p->fkinf = p->fkinf * fact2;
    p->ak = p->ak * fact2;
#ifdef CBE3_NH3OH
    p->fkh2oeq = p->fkh2oeq * fact2;
    p->angh2oeq = p->angh2oeq * fact3;
#endif
}

/* test hook (not in the source): r0ch held at this value when it is > 0, for the finite differences of the per-term
 * tests -- the analytic gradient of the source treats r0ch as a constant */
static double N3(r0ch_frozen) = -1.0;

/* ---- coorden (egrad_clnh3.f:216-306, egrad_nh3oh.f:313-426) ---- */
static void N3(coorden)(const N3(par) * p, N3(state) * s)
{
    const real argmax = 19.0;
    real P1, argp1, t1tmp;
    int ind, i;
    for (ind = 1; ind <= 3; ind++) {
        s->tcb[ind] = s->q[p->nc[ind]] - s->q[p->nhb[ind]];
        for (i = 1; i <= 3; i++) {
            s->tch[i][ind] = s->q[p->nc[ind]] - s->q[p->nh[i][ind]];
            s->tbh[i][ind] = s->q[p->nhb[ind]] - s->q[p->nh[i][ind]];
        }
    }
#ifdef CBE3_NH3OH
    for (ind = 1; ind <= 3; ind++) s->tno[ind] = s->q[p->no[ind]] - s->q[p->nhb[ind]];
#endif
    s->rcb = sqrt(s->tcb[1] * s->tcb[1] + s->tcb[2] * s->tcb[2] + s->tcb[3] * s->tcb[3]);
#ifdef CBE3_NH3OH
    s->rno = sqrt(s->tno[1] * s->tno[1] + s->tno[2] * s->tno[2] + s->tno[3] * s->tno[3]);
#endif
    for (i = 1; i <= 3; i++) {
        s->rch[i] = sqrt(s->tch[i][1] * s->tch[i][1] + s->tch[i][2] * s->tch[i][2] + s->tch[i][3] * s->tch[i][3]);
        s->rbh[i] = sqrt(s->tbh[i][1] * s->tbh[i][1] + s->tbh[i][2] * s->tbh[i][2] + s->tbh[i][3] * s->tbh[i][3]);
    }
    /* jcc-2010: r0ch between its reactant and product values */
    P1 = 1.0;
    for (i = 1; i <= 3; i++) {
        argp1 = (p->w1 * (s->rch[i] - p->w2));
        if (argp1 < argmax)
            t1tmp = 1.0 - tanh(argp1);
        else
            t1tmp = 0.0;
        P1 = P1 * t1tmp;
    }
    s->r0ch = P1 * p->r0chr + (1.0 - P1) * p->r0chp;
    if (N3(r0ch_frozen) > 0.0) s->r0ch = N3(r0ch_frozen);
#ifdef CBE3_NH3OH
    {
        real P2 = 1.0, argp2 = (p->w3 * (s->rno - p->w4)), t2tmp;
        if (argp2 < argmax)
            t2tmp = 1.0 - tanh(argp2);
        else
            t2tmp = 0.0;
        P2 = P2 * t2tmp;
        s->r0hh = P2 * p->r0hhr + (1.0 - P2) * p->r0hhp;
    }
#else
    s->r0hh = p->r0hh;
#endif
}

static real N3(ipow)(real x, int n)
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
#define ipow N3(ipow)

/* ---- switchf (egrad_clnh3.f:1552-1688, egrad_nh3oh.f:1750-1891) ---- */
static void N3(switchf)(const N3(par) * p, N3(state) * s)
{
    const real argmax = 19.0;
    int i;
#ifdef CBE3_NH3OH
    const double a1s = 1.5313681e-9, b1s = -1.6696246, a2s = 1.0147402e-9, b2s = -1.363798;
#else
    const double a1s = 1.5313681e-7, b1s = -4.6696246, a2s = 1.0147402e-7, b2s = -12.363798;
#endif
    for (i = 1; i <= 3; i++) {
        const real rch = s->rch[i], r0ch = s->r0ch;
        real args1, args2, args3;
        args1 = a1s * (rch - r0ch) * ipow(rch - b1s, 8);
        if (args1 < argmax) {
            s->s1[i] = 1.0 - tanh(args1);
            s->ds1[i] = a1s * (ipow(rch - b1s, 8) + 8.0 * (rch - r0ch) * ipow(rch - b1s, 7));
            s->ds1[i] = -s->ds1[i] / ipow(cosh(args1), 2);
        } else {
            s->s1[i] = 0.0;
            s->ds1[i] = 0.0;
        }
        args2 = a2s * (rch - r0ch) * ipow(rch - b2s, 6);
        if (args2 < argmax) {
            s->s2[i] = 1.0 - tanh(args2);
            s->ds2[i] = a2s * (ipow(rch - b2s, 6) + 6.0 * (rch - r0ch) * ipow(rch - b2s, 5));
            s->ds2[i] = -s->ds2[i] / ipow(cosh(args2), 2);
        } else {
            s->s2[i] = 0.0;
            s->ds2[i] = 0.0;
        }
        args3 = p->a3s * (rch - r0ch) * ipow(rch - p->b3s, 2);
        if (args3 < argmax) {
            s->s3[i] = 1.0 - tanh(args3);
            s->ds3[i] = p->a3s * (3.0 * ipow(rch, 2) - 2.0 * rch * (r0ch + 2.0 * p->b3s) + p->b3s * (p->b3s + 2.0 * r0ch));
            s->ds3[i] = -s->ds3[i] / ipow(cosh(args3), 2);
        } else {
            s->s3[i] = 0.0;
            s->ds3[i] = 0.0;
        }
        if (rch < 3.8) {
            real argsphi = p->aphi * (rch - r0ch) * exp(p->bphi * ipow(rch - p->cphi, 3));
            s->sphi[i] = 1.0 - tanh(argsphi);
            s->dsphi[i] = p->aphi * (1.0 + 3.0 * p->bphi * (rch - r0ch) * ipow(rch - p->cphi, 2));
            s->dsphi[i] = s->dsphi[i] * exp(p->bphi * ipow(rch - p->cphi, 3));
            s->dsphi[i] = -s->dsphi[i] / ipow(cosh(argsphi), 2);
        } else {
            s->sphi[i] = 0.0;
            s->dsphi[i] = 0.0;
        }
        if (rch < 3.8) {
            real argstheta = p->atheta * (rch - r0ch) * exp(p->btheta * ipow(rch - p->ctheta, 3));
            s->stheta[i] = 1.0 - tanh(argstheta);
            s->dstheta[i] = p->atheta * (1.0 + 3.0 * p->btheta * (rch - r0ch) * ipow(rch - p->ctheta, 2));
            s->dstheta[i] = s->dstheta[i] * exp(p->btheta * ipow(rch - p->ctheta, 3));
            s->dstheta[i] = -s->dstheta[i] / ipow(cosh(argstheta), 2);
        } else {
            s->stheta[i] = 0.0;
            s->dstheta[i] = 0.0;
        }
    }
}

/* ---- refangles (egrad_clnh3.f:312-424, egrad_nh3oh.f:428-549) ---- */
static void N3(refangles)(const N3(par) * p, N3(state) * s)
{
    const real pi = N3_PI;
    const real twopi = 2.0 * pi;
    const real tau = p->tau, taunh2 = p->taunh2;
    const real ppito = (twopi - taunh2) / 2.0;
    real(*theta0)[4] = s->theta0, (*dtheta0)[4][4] = s->dtheta0;
    const real *sphi = s->sphi, *dsphi = s->dsphi, *stheta = s->stheta, *dstheta = s->dstheta;
    int i, j, k;
    for (i = 1; i <= 3; i++) {
        theta0[i][i] = 0.0;
        for (k = 1; k <= 3; k++) dtheta0[i][i][k] = 0.0; /* the source runs k to 4: dtheta0(4,4,4) */
    }
    theta0[1][2] = tau + (tau - ppito) * (sphi[1] * sphi[2] - 1.0) + (tau - taunh2) * (stheta[3] - 1.0);
    theta0[1][3] = tau + (tau - ppito) * (sphi[1] * sphi[3] - 1.0) + (tau - taunh2) * (stheta[2] - 1.0);
    theta0[2][3] = tau + (tau - ppito) * (sphi[2] * sphi[3] - 1.0) + (tau - taunh2) * (stheta[1] - 1.0);
    dtheta0[1][2][1] = (tau - ppito) * dsphi[1] * sphi[2];
    dtheta0[1][3][1] = (tau - ppito) * dsphi[1] * sphi[3];
    dtheta0[2][3][1] = (tau - taunh2) * dstheta[1];
    dtheta0[1][2][2] = (tau - ppito) * sphi[1] * dsphi[2];
    dtheta0[1][3][2] = (tau - taunh2) * dstheta[2];
    dtheta0[2][3][2] = (tau - ppito) * dsphi[2] * sphi[3];
    dtheta0[1][2][3] = (tau - taunh2) * dstheta[3];
    dtheta0[1][3][3] = (tau - ppito) * sphi[1] * dsphi[3];
    dtheta0[2][3][3] = (tau - ppito) * sphi[2] * dsphi[3];
    for (i = 1; i <= 2; i++)
        for (j = i + 1; j <= 3; j++) {
            theta0[j][i] = theta0[i][j];
            for (k = 1; k <= 3; k++) dtheta0[j][i][k] = dtheta0[i][j][k];
        }
}

/* ---- stretch (egrad_clnh3.f:580-783, egrad_nh3oh.f:706-929) ---- */
static void N3(stretch)(const N3(par) * p, N3(state) * s, real *vstr_out)
{
    real vqch[4], vjch[4], vqbh[4], vjbh[4], vq[4], vj[4], achdc[4], achdh[4][4];
    real rav, vstr, arga, ach, dumach, e1, e3, vqcb, vjcb, dumqcb;
    const double r0cb = p->r0cb, acb = p->acb, ahh = p->ahh, d1cb = p->d1cb, d1ch = p->d1ch, d3ch = p->d3ch,
                 d1hh = p->d1hh, d3hh = p->d3hh;
    const real r0ch = s->r0ch, r0hh = s->r0hh;
#ifdef CBE3_NH3OH
    real d3cb, texp, dt, expterm, vno;
#else
    const double d3cb = p->d3cb;
#endif
    real *rch = s->rch, *rbh = s->rbh, rcb = s->rcb, *pdot = s->pdot;
    real(*tch)[4] = s->tch, (*tbh)[4] = s->tbh, *tcb = s->tcb;
    const int *nc = p->nc, *nhb = p->nhb;
    int i, ind, j, k;
    rav = (rch[1] + rch[2] + rch[3]) / 3.0;
    vstr = 0.0;
#ifdef CBE3_NH3OH
    /* :752-753; x**4.d0 is a real power of a (usually negative) base: libm pow */
    texp = exp(-pow(4.0 * (rav - p->rcbsp) / p->b3cb, 4.0));
    d3cb = (p->d3cbi - p->a3cb) + p->a3cb * texp;
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
#ifdef CBE3_NH3OH
    /* O-H Morse term (:766-768); (..)**2.d0 is expanded to a product */
    dt = (s->rno - r0hh);
    expterm = exp(-ahh * dt);
    vno = d1hh * ((1.0 - expterm) * (1.0 - expterm));
#endif
    for (i = 1; i <= 3; i++) {
        e1 = d1ch * (exp(-2.0 * ach * (rch[i] - r0ch)) - 2.0 * exp(-ach * (rch[i] - r0ch)));
        e3 = d3ch * (exp(-2.0 * ach * (rch[i] - r0ch)) + 2.0 * exp(-ach * (rch[i] - r0ch)));
        vqch[i] = (e1 + e3) * 0.5;
        vjch[i] = (e1 - e3) * 0.5;
        e1 = d1hh * (exp(-2.0 * ahh * (rbh[i] - r0hh)) - 2.0 * exp(-ahh * (rbh[i] - r0hh)));
        e3 = d3hh * (exp(-2.0 * ahh * (rbh[i] - r0hh)) + 2.0 * exp(-ahh * (rbh[i] - r0hh)));
        vqbh[i] = (e1 + e3) * 0.5;
        vjbh[i] = (e1 - e3) * 0.5;
        vq[i] = vqch[i] + vqcb + vqbh[i];
        vj[i] = -sqrt((ipow(vjch[i] - vjcb, 2) + ipow(vjcb - vjbh[i], 2) + ipow(vjbh[i] - vjch[i], 2)) * 0.5);
        vstr = vstr + vq[i] + vj[i];
    }
#ifdef CBE3_NH3OH
    vstr = vstr + vno;
#endif
    *vstr_out = vstr;
#ifdef CBE3_NH3OH
    /* the partial derivatives that follow in the source (:793-927) go to pdot, which POT_nh3oh discards */
    (void)achdc, (void)achdh, (void)dumach, (void)dumqcb, (void)pdot, (void)tch, (void)tbh, (void)tcb, (void)nc,
        (void)nhb, (void)ind, (void)j, (void)k, (void)vq;
#else
    for (ind = 1; ind <= 3; ind++) {
        achdc[ind] = dumach * (tch[1][ind] / rch[1] + tch[2][ind] / rch[2] + tch[3][ind] / rch[3]) / 3.0;
        for (i = 1; i <= 3; i++) achdh[i][ind] = -dumach * tch[i][ind] / rch[i] / 3.0;
    }
    dumqcb = -acb * ((d1cb + d3cb) * exp(-2.0 * acb * (rcb - r0cb)) - (d1cb - d3cb) * exp(-acb * (rcb - r0cb))) / rcb;
    for (i = 1; i <= 3; i++) {
        real dumqbh, factj, dumjcb, dumjbh;
        dumqbh = -ahh * ((d1hh + d3hh) * exp(-2.0 * ahh * (rbh[i] - r0hh)) - (d1hh - d3hh) * exp(-ahh * (rbh[i] - r0hh))) /
                 rbh[i];
        factj = 0.5 / vj[i];
        dumjcb = -acb * ((d1cb - d3cb) * exp(-2.0 * acb * (rcb - r0cb)) - (d1cb + d3cb) * exp(-acb * (rcb - r0cb))) *
                 factj / rcb;
        dumjbh = -ahh * ((d1hh - d3hh) * exp(-2.0 * ahh * (rbh[i] - r0hh)) - (d1hh + d3hh) * exp(-ahh * (rbh[i] - r0hh))) *
                 factj / rbh[i];
        for (ind = 1; ind <= 3; ind++) {
            real dumqch, dumqhi, dumjch, dumjhi;
            const int nhi = p->nh[i][ind];
            /* deriv wrt hb */
            pdot[nhb[ind]] = pdot[nhb[ind]] - tcb[ind] * dumqcb + tbh[i][ind] * dumqbh +
                             (vjch[i] - vjcb) * (dumjcb * tcb[ind]) +
                             (vjcb - vjbh[i]) * (-dumjcb * tcb[ind] - dumjbh * tbh[i][ind]) +
                             (vjbh[i] - vjch[i]) * dumjbh * tbh[i][ind];
            /* dvqch(i)/dc */
            dumqch = -(ach * tch[i][ind] / rch[i] + achdc[ind] * (rch[i] - r0ch)) *
                     ((d1ch + d3ch) * exp(-2.0 * ach * (rch[i] - r0ch)) - (d1ch - d3ch) * exp(-ach * (rch[i] - r0ch)));
            pdot[nc[ind]] = pdot[nc[ind]] + dumqch + tcb[ind] * dumqcb;
            /* dvqch(i)/dh(i) */
            dumqhi = (ach * tch[i][ind] / rch[i] - achdh[i][ind] * (rch[i] - r0ch)) *
                     ((d1ch + d3ch) * exp(-2.0 * ach * (rch[i] - r0ch)) - (d1ch - d3ch) * exp(-ach * (rch[i] - r0ch)));
            pdot[nhi] = pdot[nhi] + dumqhi - tbh[i][ind] * dumqbh;
            /* dvjch(i)/dc */
            dumjch = -(ach * tch[i][ind] / rch[i] + achdc[ind] * (rch[i] - r0ch)) *
                     ((d1ch - d3ch) * exp(-2.0 * ach * (rch[i] - r0ch)) - (d1ch + d3ch) * exp(-ach * (rch[i] - r0ch))) *
                     factj;
            pdot[nc[ind]] = pdot[nc[ind]] + (vjch[i] - vjcb) * (dumjch - dumjcb * tcb[ind]) +
                            (vjcb - vjbh[i]) * dumjcb * tcb[ind] - (vjbh[i] - vjch[i]) * dumjch;
            /* dvjch(i)/dh(i) */
            dumjhi = (ach * tch[i][ind] / rch[i] - achdh[i][ind] * (rch[i] - r0ch)) *
                     ((d1ch - d3ch) * exp(-2.0 * ach * (rch[i] - r0ch)) - (d1ch + d3ch) * exp(-ach * (rch[i] - r0ch))) *
                     factj;
            pdot[nhi] = pdot[nhi] + (vjch[i] - vjcb) * dumjhi + (vjcb - vjbh[i]) * dumjbh * tbh[i][ind] +
                        (vjbh[i] - vjch[i]) * (-dumjbh * tbh[i][ind] - dumjhi);
            /* dv(i)/dh(j) */
            for (k = 1; k <= 2; k++) {
                real dumqhj, dumjhj;
                j = i + k;
                if (j > 3) j = j - 3;
                dumqhj = -achdh[j][ind] * (rch[i] - r0ch) *
                         ((d1ch + d3ch) * exp(-2.0 * ach * (rch[i] - r0ch)) - (d1ch - d3ch) * exp(-ach * (rch[i] - r0ch)));
                dumjhj = -achdh[j][ind] * (rch[i] - r0ch) *
                         ((d1ch - d3ch) * exp(-2.0 * ach * (rch[i] - r0ch)) - (d1ch + d3ch) * exp(-ach * (rch[i] - r0ch))) *
                         factj;
                pdot[p->nh[j][ind]] = pdot[p->nh[j][ind]] + dumqhj + (vjch[i] - vjcb) * dumjhj -
                                      (vjbh[i] - vjch[i]) * dumjhj;
            }
        }
    }
#endif
}

/* ---- ipforce (egrad_clnh3.f:1416-1546, egrad_nh3oh.f:1615-1748) ---- */
static void N3(ipforce)(const N3(par) * p, N3(state) * s)
{
    real dfko[4], df1dc[4], df1dh[4];
    const real r0ch = s->r0ch, r0hh = s->r0hh;
    const real *rch = s->rch, *rbh = s->rbh;
    real(*fk0)[4] = s->fk0, *f1 = s->f1, (*dfdc)[4][4] = s->dfdc, (*dfdh)[4][4] = s->dfdh;
    int i;
#ifdef CBE3_NH3OH
    const double f0 = p->fkinf + p->ak, f2 = p->fkinf;
    fk0[1][2] = f0 + f0 * (s->s1[1] * s->s1[2] - 1.0) + (f0 - f2) * (s->s2[3] - 1.0);
    fk0[1][3] = f0 + f0 * (s->s1[1] * s->s1[3] - 1.0) + (f0 - f2) * (s->s2[2] - 1.0);
    fk0[2][3] = f0 + f0 * (s->s1[2] * s->s1[3] - 1.0) + (f0 - f2) * (s->s2[1] - 1.0);
#else
    real argfk0, fko;
    argfk0 = p->bk * (ipow(rch[1] - r0ch, 2) + ipow(rch[2] - r0ch, 2) + ipow(rch[3] - r0ch, 2));
    fko = p->fkinf + p->ak * exp(-argfk0);
    fk0[1][2] = fko;
    fk0[1][3] = fko;
    fk0[2][3] = fko;
#endif
    for (i = 1; i <= 3; i++) {
        real arga1, arga2, a1, a2, duma1, duma2;
        arga1 = p->aa1 * rbh[i] * rbh[i];
        arga2 = p->aa4 * (rbh[i] - r0hh) * (rbh[i] - r0hh);
        a1 = 1.0 - exp(-arga1);
        a2 = p->aa2 + p->aa3 * exp(-arga2);
        f1[i] = a1 * exp(-a2 * ipow(rch[i] - r0ch, 2));
#ifndef CBE3_NH3OH
        dfko[i] = -2.0 * p->ak * p->bk * (rch[i] - r0ch) * exp(-argfk0);
        duma1 = 2.0 * p->aa1 * rbh[i] * exp(-arga1);
        duma2 = -2.0 * p->aa3 * p->aa4 * (rbh[i] - r0hh) * exp(-arga2);
        df1dc[i] = -2.0 * (rch[i] - r0ch) * a1 * a2 * exp(-a2 * ipow(rch[i] - r0ch, 2));
        df1dh[i] = duma1 * exp(-a2 * ipow(rch[i] - r0ch, 2)) -
                   duma2 * ipow(rch[i] - r0ch, 2) * a1 * exp(-a2 * ipow(rch[i] - r0ch, 2));
#else
        (void)duma1, (void)duma2;
#endif
    }
#ifndef CBE3_NH3OH
    dfdc[1][2][1] = dfko[1] * f1[1] * f1[2] + fko * df1dc[1] * f1[2];
    dfdc[1][2][2] = dfko[2] * f1[1] * f1[2] + fko * f1[1] * df1dc[2];
    dfdc[1][2][3] = dfko[3] * f1[1] * f1[2];
    dfdc[1][3][1] = dfko[1] * f1[1] * f1[3] + fko * df1dc[1] * f1[3];
    dfdc[1][3][2] = dfko[2] * f1[1] * f1[3];
    dfdc[1][3][3] = dfko[3] * f1[1] * f1[3] + fko * f1[1] * df1dc[3];
    dfdc[2][3][1] = dfko[1] * f1[2] * f1[3];
    dfdc[2][3][2] = dfko[2] * f1[2] * f1[3] + fko * df1dc[2] * f1[3];
    dfdc[2][3][3] = dfko[3] * f1[2] * f1[3] + fko * f1[2] * df1dc[3];
    dfdh[1][2][1] = fko * df1dh[1] * f1[2];
    dfdh[1][2][2] = fko * f1[1] * df1dh[2];
    dfdh[1][2][3] = 0.0;
    dfdh[1][3][1] = fko * df1dh[1] * f1[3];
    dfdh[1][3][2] = 0.0;
    dfdh[1][3][3] = fko * f1[1] * df1dh[3];
    dfdh[2][3][1] = 0.0;
    dfdh[2][3][2] = fko * df1dh[2] * f1[3];
    dfdh[2][3][3] = fko * f1[2] * df1dh[3];
#else
    (void)dfko, (void)df1dc, (void)df1dh, (void)dfdc, (void)dfdh;
#endif
}

/* ---- ipbend (egrad_clnh3.f:941-1058, egrad_nh3oh.f:1088-1251) ---- */
static void N3(ipbend)(const N3(par) * p, N3(state) * s, real *vip_out)
{
    real costh[4][4], theta[4][4], dth[4][4];
    real vip = 0.0;
    real *rch = s->rch, *rbh = s->rbh, *pdot = s->pdot, *f1 = s->f1;
    real(*tch)[4] = s->tch, (*tbh)[4] = s->tbh, (*fk0)[4] = s->fk0;
    const int *nc = p->nc, *nhb = p->nhb;
    int i, j, k, ind;
    N3(ipforce)(p, s);
    for (i = 1; i <= 2; i++)
        for (j = i + 1; j <= 3; j++) {
            real termth;
            costh[i][j] = tch[i][1] * tch[j][1] + tch[i][2] * tch[j][2] + tch[i][3] * tch[j][3];
            costh[i][j] = costh[i][j] / rch[i] / rch[j];
            theta[i][j] = acos(costh[i][j]);
            dth[i][j] = theta[i][j] - s->theta0[i][j];
            vip = vip + 0.5 * fk0[i][j] * f1[i] * f1[j] * ipow(dth[i][j], 2);
#ifndef CBE3_NH3OH
            termth = -1.0 / sqrt(1.0 - costh[i][j] * costh[i][j]);
            for (ind = 1; ind <= 3; ind++) {
                real dthi, dthj, dthc;
                dthi = -tch[j][ind] / rch[i] / rch[j] + costh[i][j] * tch[i][ind] / rch[i] / rch[i];
                dthi = dthi * termth;
                dthj = -tch[i][ind] / rch[i] / rch[j] + costh[i][j] * tch[j][ind] / rch[j] / rch[j];
                dthj = dthj * termth;
                dthc = -(dthi + dthj);
                pdot[p->nh[i][ind]] = pdot[p->nh[i][ind]] + fk0[i][j] * f1[i] * f1[j] * dthi * dth[i][j];
                pdot[p->nh[j][ind]] = pdot[p->nh[j][ind]] + fk0[i][j] * f1[i] * f1[j] * dthj * dth[i][j];
                pdot[nc[ind]] = pdot[nc[ind]] + fk0[i][j] * f1[i] * f1[j] * dthc * dth[i][j];
                for (k = 1; k <= 3; k++) {
                    real dth0k, dth0c;
                    dth0k = -s->dtheta0[i][j][k] * tch[k][ind] / rch[k];
                    dth0c = -dth0k;
                    pdot[p->nh[k][ind]] = pdot[p->nh[k][ind]] -
                                          0.5 * tch[k][ind] * s->dfdc[i][j][k] * ipow(dth[i][j], 2) / rch[k] -
                                          0.5 * tbh[k][ind] * s->dfdh[i][j][k] * ipow(dth[i][j], 2) / rbh[k] -
                                          fk0[i][j] * f1[i] * f1[j] * dth0k * dth[i][j];
                    pdot[nc[ind]] = pdot[nc[ind]] + 0.5 * tch[k][ind] * s->dfdc[i][j][k] * ipow(dth[i][j], 2) / rch[k] -
                                    fk0[i][j] * f1[i] * f1[j] * dth0c * dth[i][j];
                    pdot[nhb[ind]] = pdot[nhb[ind]] + 0.5 * tbh[k][ind] * s->dfdh[i][j][k] * ipow(dth[i][j], 2) / rbh[k];
                }
            }
#else
            (void)termth;
#endif
        }
#ifdef CBE3_NH3OH
    /* H-O-H bends of the forming water (:1213-1248) */
    for (i = 1; i <= 3; i++) {
        real dot = 0.0, cosine;
        for (j = 1; j <= 3; j++) dot = dot - s->tno[j] * tbh[i][j];
        cosine = dot / (s->rno * rbh[i]);
        cosine = (cosine > -1.0) ? cosine : (real)-1.0; /* min(1, max(-1, cosine)) */
        cosine = (cosine < 1.0) ? cosine : (real)1.0;
        s->angh2o[i] = acos(cosine);
    }
    for (i = 1; i <= 3; i++) {
        real arga = p->alph2o * (rbh[i] - s->r0hh);
        if (arga < 19.0)
            s->fkh2o[i] = p->fkh2oeq * (1.0 - tanh(arga));
        else
            s->fkh2o[i] = 0.0;
    }
    for (i = 1; i <= 3; i++) {
        real dang = (s->angh2o[i] - p->angh2oeq);
        vip = vip + 0.5 * s->fkh2o[i] * dang * dang;
    }
    (void)pdot, (void)nc, (void)nhb, (void)k, (void)ind, (void)rbh;
#endif
    *vip_out = vip;
}

/* energy in the routine's 1e5 J/mol from the state's q */
static void N3(energy)(const N3(par) * p, N3(state) * s, real *vstr, real *vip)
{
    N3(coorden)(p, s);
    N3(switchf)(p, s);
    N3(refangles)(p, s);
    N3(stretch)(p, s, vstr);
    N3(ipbend)(p, s, vip);
}

/* ---- POT: R(1..3 natoms) cartesians in bohr -> energy (hartree), DEGSDR (hartree/bohr) ----
 * gparts (may be null): [2][N3_NC] = d(vstr), d(vip) per Angstrom in the routine's units (clnh3 only) */
static void N3(pot)(const N3(par) * p, const real R[N3_NC + 1], real *en_out, real DEGSDR[N3_NC + 1], real parts[3],
                    real *gparts)
{
    N3(state) s;
    real vstr, vip, en, ENGYGS;
    int i;
    for (i = 1; i <= N3_NC; i++) {
        s.q[i] = R[i] * 0.52918;
        s.pdot[i] = 0.0;
    }
#ifdef CBE3_NH3OH
    N3(energy)(p, &s, &vstr, &vip);
    en = vstr + vip;
    en = en * 0.03812;
    ENGYGS = en;
    if (parts) {
        parts[0] = vstr;
        parts[1] = 0.0;
        parts[2] = vip;
    }
    {
        /* numeric derivatives (:283-296) */
        const real PASO = 1.0e-5;
        real v1, v2;
        for (i = 1; i <= N3_NC; i++) {
            s.q[i] = s.q[i] + PASO;
            N3(energy)(p, &s, &v1, &v2);
            en = v1 + v2;
            en = en * 0.03812;
            DEGSDR[i] = (en - ENGYGS) / PASO;
            DEGSDR[i] = DEGSDR[i] * 0.52918;
            s.q[i] = s.q[i] - PASO;
        }
    }
    (void)gparts;
#else
    N3(coorden)(p, &s);
    N3(switchf)(p, &s);
    N3(refangles)(p, &s);
    N3(stretch)(p, &s, &vstr);
    if (gparts)
        for (i = 1; i <= N3_NC; i++) gparts[i - 1] = s.pdot[i];
    N3(ipbend)(p, &s, &vip);
    if (gparts)
        for (i = 1; i <= N3_NC; i++) gparts[N3_NC + i - 1] = s.pdot[i] - gparts[i - 1];
    en = vstr + vip;
    en = en * 0.03812;
    ENGYGS = en;
    for (i = 1; i <= N3_NC; i++) DEGSDR[i] = s.pdot[i] * 0.0201723;
    if (parts) {
        parts[0] = vstr;
        parts[1] = s.r0ch; /* the geometry-dependent reference length, for the frozen-r0ch finite differences */
        parts[2] = vip;
    }
#endif
    *en_out = ENGYGS;
}

void N3_EGRAD(const real *q, int natoms, int nbeads, real *V, real *dVdq, int *info)
{
    N3(par) par;
    int k, i, j;
    N3(prepot)(&par);
    *info = 0;
    for (k = 0; k < nbeads; k++) {
        const real *qk = q + (long)k * 3 * natoms;
        real *gk = dVdq + (long)k * 3 * natoms;
        real R[N3_NC + 1], D[N3_NC + 1];
        for (j = 0; j < N3_NAT; j++)
            for (i = 0; i < 3; i++) R[3 * j + i + 1] = qk[3 * j + i];
        N3(pot)(&par, R, &V[k], D, (real *)0, (real *)0);
        for (j = 0; j < natoms; j++)
            for (i = 0; i < 3; i++) gk[3 * j + i] = (j < N3_NAT) ? D[3 * j + i + 1] : (real)0.0;
    }
}

/* energy split for the per-term tests: parts = (vstr, r0ch [clnh3] or 0 [nh3oh], vip) in 1e5 J/mol / Angstrom */
void N3_PARTS(const real *qin, real parts[3], real *V)
{
    N3(par) par;
    real R[N3_NC + 1], D[N3_NC + 1];
    int i;
    N3(prepot)(&par);
    for (i = 0; i < N3_NC; i++) R[i + 1] = qin[i];
    N3(pot)(&par, R, V, D, parts, (real *)0);
}

/* energy split with r0ch held at r0ch_frozen (<= 0: as in the source) */
void N3(parts_frozen_real)(const real *qin, double r0ch_frozen, real parts[3], real *V)
{
    N3(par) par;
    real R[N3_NC + 1], D[N3_NC + 1];
    int i;
    N3(prepot)(&par);
    for (i = 0; i < N3_NC; i++) R[i + 1] = qin[i];
    N3(r0ch_frozen) = r0ch_frozen;
    N3(pot)(&par, R, V, D, parts, (real *)0);
    N3(r0ch_frozen) = -1.0;
}

/* the same with the analytic gradient of the two parts (clnh3: [2][15], per Angstrom; nh3oh: untouched) */
void N3_PARTS_GRAD(const real *qin, real parts[3], real *gparts)
{
    N3(par) par;
    real R[N3_NC + 1], D[N3_NC + 1], V;
    int i;
    N3(prepot)(&par);
    for (i = 0; i < N3_NC; i++) R[i + 1] = qin[i];
    N3(pot)(&par, R, &V, D, parts, gparts);
}

#undef ipow
