/*
 * ewald.c -- CPU oracle of the SPME reciprocal-space sum.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Parity UNPINNED by the reference: ewald_recip has no
 * reachable caller there (ff_nonb.f90:337 hard-sets ewald=.false.), so this restatement
 * is pinned against the plain Ewald reciprocal sum (orc_ewald_direct_recip) instead.
 *
 * Literal restatement of
 *   set_periodic.f90:114-231   grid size from the `multi` table (x box length only), Ewald
 *                              coefficient by bisection on erfc(a r_c)/r_c = 1e-8 with r_c = 7 A,
 *                              bsorder = 5, B-spline moduli                       orc_ewald_setup
 *   bspline.f90:30-60, dftmod.f90:30-101                                          bspline, dftmod
 *   bsplgen.f90:30-98          B-spline values and first derivatives (level = 2)  bsplgen
 *   ewald_recip.f90:30-470     charge spreading (nchunk = 1), forward FFT, influence function and
 *                              energy, backward FFT, gradient                     orc_ewald_recip
 * The two dfftw_execute_dft calls are complex 3-D DFTs (sign -1 forward, +1 backward, unnormalised);
 * restated as three passes of dense 1-D DFTs, exact to rounding.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "ewald.h"

static const double PI = 3.1415926535897932384626433832795029;
static const int MULTI[63] = {2,   4,   6,   8,   10,  12,  16,  18,  20,  24,  30,  32,  36,  40,  48,  50,
                              54,  60,  64,  72,  80,  90,  96,  100, 108, 120, 128, 144, 150, 160, 162, 180,
                              192, 200, 216, 240, 250, 256, 270, 288, 300, 320, 324, 360, 384, 400, 432, 450,
                              480, 486, 500, 512, 540, 576, 600, 640, 648, 720, 750, 768, 800, 810, 864};

/* bspline.f90 */
static void bspline(double x, int n, double *c /* 1-based */)
{
    int i, k;
    c[1] = 1.0 - x;
    c[2] = x;
    for (k = 3; k <= n; k++) {
        const double denom = 1.0 / (double)(k - 1);
        c[k] = x * c[k - 1] * denom;
        for (i = 1; i <= k - 2; i++) c[k - i] = ((x + (double)i) * c[k - i - 1] + ((double)(k - i) - x) * c[k - i]) * denom;
        c[1] = (1.0 - x) * c[1] * denom;
    }
}

/* dftmod.f90 */
static void dftmod(double *bsmod /* 1-based */, const double *bsarray /* 1-based */, int nfft, int order)
{
    int i, j, k;
    double factor = 2.0 * PI / (double)nfft;
    const double eps = 1.0e-7;
    const int jcut = 50, order2 = 2 * order;
    for (i = 1; i <= nfft; i++) {
        double sum1 = 0.0, sum2 = 0.0;
        for (j = 1; j <= nfft; j++) {
            const double arg = factor * (double)((i - 1) * (j - 1));
            sum1 = sum1 + bsarray[j] * cos(arg);
            sum2 = sum2 + bsarray[j] * sin(arg);
        }
        bsmod[i] = sum1 * sum1 + sum2 * sum2;
    }
    if (bsmod[1] < eps) bsmod[1] = 0.5 * bsmod[2];
    for (i = 2; i <= nfft - 1; i++)
        if (bsmod[i] < eps) bsmod[i] = 0.5 * (bsmod[i - 1] + bsmod[i + 1]);
    if (bsmod[nfft] < eps) bsmod[nfft] = 0.5 * bsmod[nfft - 1];
    for (i = 1; i <= nfft; i++) {
        double zeta;
        k = i - 1;
        if (i > nfft / 2) k = k - nfft;
        if (k == 0) {
            zeta = 1.0;
        } else {
            double sum1 = 1.0, sum2 = 1.0;
            factor = PI * (double)k / (double)nfft;
            for (j = 1; j <= jcut; j++) {
                const double arg = factor / (factor + PI * (double)j);
                sum1 = sum1 + pow(arg, order);
                sum2 = sum2 + pow(arg, order2);
            }
            for (j = 1; j <= jcut; j++) {
                const double arg = factor / (factor - PI * (double)j);
                sum1 = sum1 + pow(arg, order);
                sum2 = sum2 + pow(arg, order2);
            }
            zeta = sum2 / sum1;
        }
        bsmod[i] = bsmod[i] * (zeta * zeta);
    }
}

/* set_periodic.f90:114-231 */
orc_ewald *orc_ewald_setup(const double box[3])
{
    orc_ewald *E = (orc_ewald *)calloc(1, sizeof(orc_ewald));
    const double r_ew_cut = 7 / 0.52917721092;
    const int maxfft = 864, minfft = 16;
    const double dens = 1.2, delta = 1e-8, eps = 1.0e-8;
    int i, k, ifft, nfft;
    double ratio, x_par, y_par, xlo, xhi, array[8], *bsarray;
    for (i = 0; i < 3; i++) E->box[i] = box[i];
    E->volbox = box[0] * box[1] * box[2];
    ifft = (int)(box[0] * 0.52917721092 * dens - delta) + 1;
    nfft = maxfft;
    for (i = 63; i >= 1; i--) {
        k = MULTI[i - 1];
        if (k <= maxfft)
            if (k >= ifft) nfft = k;
    }
    if (nfft < minfft) nfft = minfft;
    E->nfft = nfft;
    ratio = eps + 1.0;
    x_par = 0.5;
    i = 0;
    while (ratio >= eps) {
        i = i + 1;
        x_par = 2.0 * x_par;
        y_par = x_par * r_ew_cut;
        ratio = erfc(y_par) / r_ew_cut;
    }
    k = i + 60;
    xlo = 0.0;
    xhi = x_par;
    for (i = 1; i <= k; i++) {
        x_par = (xlo + xhi) / 2.0;
        y_par = x_par * r_ew_cut;
        ratio = erfc(y_par) / r_ew_cut;
        if (ratio >= eps)
            xlo = x_par;
        else
            xhi = x_par;
    }
    E->a_ewald = x_par;
    E->bsorder = 5;
    bsarray = (double *)calloc(nfft + 2, sizeof(double));
    bspline(0.0, E->bsorder, array);
    for (i = 1; i <= E->bsorder; i++) bsarray[i + 1] = array[i];
    E->bsmod1 = (double *)calloc(nfft + 1, sizeof(double));
    E->bsmod2 = (double *)calloc(nfft + 1, sizeof(double));
    E->bsmod3 = (double *)calloc(nfft + 1, sizeof(double));
    dftmod(E->bsmod1, bsarray, nfft, E->bsorder); /* 1-based: element 0 unused */
    dftmod(E->bsmod2, bsarray, nfft, E->bsorder);
    dftmod(E->bsmod3, bsarray, nfft, E->bsorder);
    free(bsarray);
    return E;
}
void orc_ewald_free(orc_ewald *E)
{
    if (!E) return;
    free(E->bsmod1);
    free(E->bsmod2);
    free(E->bsmod3);
    free(E);
}
int orc_ewald_nfft(const orc_ewald *E) { return E->nfft; }
double orc_ewald_alpha(const orc_ewald *E) { return E->a_ewald; }
void orc_ewald_bsmod(const orc_ewald *E, double *out)
{
    memcpy(out, E->bsmod1 + 1, sizeof(double) * E->nfft);
    memcpy(out + E->nfft, E->bsmod2 + 1, sizeof(double) * E->nfft);
    memcpy(out + 2 * E->nfft, E->bsmod3 + 1, sizeof(double) * E->nfft);
}

/* bsplgen.f90: thetai(4,bsorder) column-major -> TH(j,i) = thetai[(i-1)*4 + (j-1)], level = 2 */
static void bsplgen(double w, double *thetai, int bsorder)
{
    double b[8][8]; /* bsbuild(i,j), 1-based */
    int i, j, k;
    b[2][2] = w;
    b[2][1] = 1.0 - w;
    b[3][3] = 0.5 * w * b[2][2];
    b[3][2] = 0.5 * ((1.0 + w) * b[2][1] + (2.0 - w) * b[2][2]);
    b[3][1] = 0.5 * (1.0 - w) * b[2][1];
    for (i = 4; i <= bsorder; i++) {
        const double denom = 1.0 / (double)(i - 1);
        k = i - 1;
        b[i][i] = denom * w * b[k][k];
        for (j = 1; j <= i - 2; j++) b[i][i - j] = denom * ((w + (double)j) * b[k][i - j - 1] + ((double)(i - j) - w) * b[k][i - j]);
        b[i][1] = denom * (1.0 - w) * b[k][1];
    }
    k = bsorder - 1;
    b[k][bsorder] = b[k][bsorder - 1];
    for (i = bsorder - 1; i >= 2; i--) b[k][i] = b[k][i - 1] - b[k][i];
    b[k][1] = -b[k][1];
    for (i = 1; i <= bsorder; i++)
        for (j = 1; j <= 2; j++) thetai[(i - 1) * 4 + (j - 1)] = b[bsorder - j + 1][i];
}

/* unnormalised complex 3-D DFT, exponent sign sgn (FFTW_FORWARD = -1, FFTW_BACKWARD = +1); grid
 * qgrid(2,nfft,nfft,nfft) Fortran order: index ((k*nf + j)*nf + i)*2 + c, 0-based */
static void dft3(double *g, int nf, int sgn)
{
    double *cs = (double *)malloc(sizeof(double) * 2 * nf), *line = (double *)malloc(sizeof(double) * 2 * nf);
    int d, a, b2, t, u;
    for (t = 0; t < nf; t++) {
        cs[2 * t] = cos(2.0 * PI * t / nf);
        cs[2 * t + 1] = sgn * sin(2.0 * PI * t / nf);
    }
    for (d = 0; d < 3; d++) {
        const size_t stride = (d == 0) ? 1 : (d == 1 ? (size_t)nf : (size_t)nf * nf);
        for (a = 0; a < nf; a++)
            for (b2 = 0; b2 < nf; b2++) {
                const size_t base = (d == 0) ? ((size_t)a * nf + b2) * nf
                                             : (d == 1 ? (size_t)a * nf * nf + b2 : (size_t)a * nf + b2);
                for (t = 0; t < nf; t++) {
                    line[2 * t] = g[2 * (base + t * stride)];
                    line[2 * t + 1] = g[2 * (base + t * stride) + 1];
                }
                for (u = 0; u < nf; u++) {
                    double re = 0.0, im = 0.0;
                    for (t = 0; t < nf; t++) {
                        const int w = (int)(((long)u * t) % nf);
                        re += line[2 * t] * cs[2 * w] - line[2 * t + 1] * cs[2 * w + 1];
                        im += line[2 * t] * cs[2 * w + 1] + line[2 * t + 1] * cs[2 * w];
                    }
                    g[2 * (base + u * stride)] = re;
                    g[2 * (base + u * stride) + 1] = im;
                }
            }
    }
    free(cs);
    free(line);
}

#define QG(c, i, j, k) qgrid[((((size_t)(k) - 1) * nfft + ((j) - 1)) * nfft + ((i) - 1)) * 2 + (c)]
#define TH(t, j, i, a) (t)[((size_t)(a) * bsorder + ((i) - 1)) * 4 + ((j) - 1)]

/* ewald_recip.f90:30-470 (orthorhombic box, nchunk = 1) */
void orc_ewald_recip(const orc_ewald *E, int n, const double *xyz, const double *q, double *energy_out, double *grad)
{
    const int nfft = E->nfft, bsorder = E->bsorder;
    const double volbox = E->volbox, a_ewald = E->a_ewald;
    const double *bsmod1 = E->bsmod1, *bsmod2 = E->bsmod2, *bsmod3 = E->bsmod3;
    double recip[3][3]; /* recip(r,c) -> recip[r-1][c-1] */
    double *qgrid = (double *)calloc((size_t)2 * nfft * nfft * nfft, sizeof(double));
    int *igrid = (int *)malloc(sizeof(int) * 3 * n);
    double *th1 = (double *)calloc((size_t)4 * bsorder * n, sizeof(double)), *th2 = (double *)calloc((size_t)4 * bsorder * n, sizeof(double)),
           *th3 = (double *)calloc((size_t)4 * bsorder * n, sizeof(double));
    const double eps = 1.0e-8;
    const int nlpts = 2, nrpts = 2, grdoff = 4;
    int i, j, k, m, ii, jj, kk, iatm;
    double energy, f, pterm, volterm;
    int npoint, nff, nf;
    memset(recip, 0, sizeof(recip));
    {
        const double ar1 = E->box[0], br2 = E->box[1], cr3 = E->box[2];
        recip[0][0] = (br2 * cr3) / volbox;
        recip[1][1] = (cr3 * ar1) / volbox;
        recip[2][2] = (ar1 * br2) / volbox;
    }
    for (i = 0; i < n; i++) {
        const double xi = xyz[3 * i], yi = xyz[3 * i + 1], zi = xyz[3 * i + 2];
        double w, fr;
        int ifr, d;
        for (d = 0; d < 3; d++) {
            w = xi * recip[0][d] + yi * recip[1][d] + zi * recip[2][d];
            fr = (double)nfft * (w - round(w) + 0.5); /* anint: half away from zero = C round */
            ifr = (int)(fr - eps);
            w = fr - (double)ifr;
            igrid[3 * i + d] = ifr - bsorder;
            bsplgen(w, (d == 0 ? th1 : (d == 1 ? th2 : th3)) + (size_t)i * 4 * bsorder, bsorder);
        }
    }
    /* spread the charges (nchunk = 1: ewald_adjust gives offset = 1 - amin) */
    for (iatm = 0; iatm < n; iatm++) {
        int nearpt[3], abound[6], off[3];
        for (m = 0; m < 3; m++) {
            nearpt[m] = igrid[3 * iatm + m] + grdoff;
            abound[2 * m] = nearpt[m] - nlpts;
            abound[2 * m + 1] = nearpt[m] + nrpts;
            off[m] = 0 + 1 - abound[2 * m];
        }
        for (kk = abound[4]; kk <= abound[5]; kk++) {
            double v0;
            k = kk;
            m = k + off[2];
            if (k < 1) k = k + nfft;
            v0 = TH(th3, 1, m, iatm) * q[iatm];
            for (jj = abound[2]; jj <= abound[3]; jj++) {
                double u0, term;
                j = jj;
                m = j + off[1];
                if (j < 1) j = j + nfft;
                u0 = TH(th2, 1, m, iatm);
                term = v0 * u0;
                for (ii = abound[0]; ii <= abound[1]; ii++) {
                    double t0;
                    i = ii;
                    m = i + off[0];
                    if (i < 1) i = i + nfft;
                    t0 = TH(th1, 1, m, iatm);
                    QG(0, i, j, k) = QG(0, i, j, k) + term * t0;
                }
            }
        }
    }
    dft3(qgrid, nfft, -1);
    f = 0.5;
    npoint = nfft * nfft * nfft;
    pterm = (PI / a_ewald) * (PI / a_ewald);
    energy = 0.0;
    volterm = PI * volbox;
    nff = nfft * nfft;
    nf = (nfft + 1) / 2;
    for (i = 1; i <= npoint - 1; i++) {
        const int k3 = i / nff + 1;
        const int jr = i - (k3 - 1) * nff;
        const int k2 = jr / nfft + 1;
        const int k1 = jr - (k2 - 1) * nfft + 1;
        int m1 = k1 - 1, m2 = k2 - 1, m3 = k3 - 1;
        double r1, r2, r3, h1, h2, h3, hsq, term, expterm;
        if (k1 > nf) m1 = m1 - nfft;
        if (k2 > nf) m2 = m2 - nfft;
        if (k3 > nf) m3 = m3 - nfft;
        r1 = (double)m1;
        r2 = (double)m2;
        r3 = (double)m3;
        h1 = recip[0][0] * r1 + recip[0][1] * r2 + recip[0][2] * r3;
        h2 = recip[1][0] * r1 + recip[1][1] * r2 + recip[1][2] * r3;
        h3 = recip[2][0] * r1 + recip[2][1] * r2 + recip[2][2] * r3;
        hsq = h1 * h1 + h2 * h2 + h3 * h3;
        term = -pterm * hsq;
        expterm = 0.0;
        if (term > -50.0) {
            const double denom = volterm * hsq * bsmod1[k1] * bsmod2[k2] * bsmod3[k3];
            double struc2, e;
            expterm = exp(term) / denom;
            struc2 = QG(0, k1, k2, k3) * QG(0, k1, k2, k3) + QG(1, k1, k2, k3) * QG(1, k1, k2, k3);
            e = f * expterm * struc2;
            energy = energy + e;
        }
        QG(0, k1, k2, k3) = expterm * QG(0, k1, k2, k3);
        QG(1, k1, k2, k3) = expterm * QG(1, k1, k2, k3);
    }
    /* NB the reference leaves qgrid(:,1,1,1) untouched (the loop starts at the second point) */
    dft3(qgrid, nfft, +1);
    f = 1.0;
    {
        const double dn = (double)nfft;
        for (iatm = 0; iatm < n; iatm++) {
            const int igrd0 = igrid[3 * iatm], jgrd0 = igrid[3 * iatm + 1], kgrd0 = igrid[3 * iatm + 2];
            const double fi = f * q[iatm];
            double de1 = 0.0, de2 = 0.0, de3 = 0.0;
            int it1, it2, it3, i0, j0, k0;
            k0 = kgrd0;
            for (it3 = 1; it3 <= bsorder; it3++) {
                double t3, dt3;
                k0 = k0 + 1;
                k = k0 + 1 + (nfft - (k0 >= 0 ? nfft : -nfft)) / 2;
                t3 = TH(th3, 1, it3, iatm);
                dt3 = dn * TH(th3, 2, it3, iatm);
                j0 = jgrd0;
                for (it2 = 1; it2 <= bsorder; it2++) {
                    double t2, dt2;
                    j0 = j0 + 1;
                    j = j0 + 1 + (nfft - (j0 >= 0 ? nfft : -nfft)) / 2;
                    t2 = TH(th2, 1, it2, iatm);
                    dt2 = dn * TH(th2, 2, it2, iatm);
                    i0 = igrd0;
                    for (it1 = 1; it1 <= bsorder; it1++) {
                        double t1, dt1, term;
                        i0 = i0 + 1;
                        i = i0 + 1 + (nfft - (i0 >= 0 ? nfft : -nfft)) / 2;
                        t1 = TH(th1, 1, it1, iatm);
                        dt1 = dn * TH(th1, 2, it1, iatm);
                        term = QG(0, i, j, k);
                        de1 = de1 + term * dt1 * t2 * t3;
                        de2 = de2 + term * dt2 * t1 * t3;
                        de3 = de3 + term * dt3 * t1 * t2;
                    }
                }
            }
            grad[3 * iatm] = fi * (recip[0][0] * de1 + recip[0][1] * de2 + recip[0][2] * de3);
            grad[3 * iatm + 1] = fi * (recip[1][0] * de1 + recip[1][1] * de2 + recip[1][2] * de3);
            grad[3 * iatm + 2] = fi * (recip[2][0] * de1 + recip[2][1] * de2 + recip[2][2] * de3);
        }
    }
    *energy_out = energy;
    free(qgrid);
    free(igrid);
    free(th1);
    free(th2);
    free(th3);
}

/* E_rec = 1/(2 pi V) sum_{m != 0} exp(-pi^2 m^2 / a^2) / m^2 |S(m)|^2, S(m) = sum_j q_j exp(2 pi i m.r_j),
 * m = (m1/Lx, m2/Ly, m3/Lz); gradient analytically.  Validation reference only. */
void orc_ewald_direct_recip(const double box[3], double alpha, int mmax, int n, const double *xyz, const double *q,
                            double *energy, double *grad)
{
    const double V = box[0] * box[1] * box[2];
    int m1, m2, m3, j;
    double e = 0.0;
    double *cj = (double *)malloc(sizeof(double) * n), *sj = (double *)malloc(sizeof(double) * n);
    memset(grad, 0, sizeof(double) * 3 * n);
    for (m1 = -mmax; m1 <= mmax; m1++)
        for (m2 = -mmax; m2 <= mmax; m2++)
            for (m3 = -mmax; m3 <= mmax; m3++) {
                const double h1 = m1 / box[0], h2 = m2 / box[1], h3 = m3 / box[2], hsq = h1 * h1 + h2 * h2 + h3 * h3;
                double sr = 0.0, si = 0.0, pre;
                if (m1 == 0 && m2 == 0 && m3 == 0) continue;
                if (PI * PI * hsq / (alpha * alpha) > 50.0) continue;
                for (j = 0; j < n; j++) {
                    const double ph = 2.0 * PI * (h1 * xyz[3 * j] + h2 * xyz[3 * j + 1] + h3 * xyz[3 * j + 2]);
                    cj[j] = cos(ph);
                    sj[j] = sin(ph);
                    sr += q[j] * cj[j];
                    si += q[j] * sj[j];
                }
                pre = exp(-PI * PI * hsq / (alpha * alpha)) / (2.0 * PI * V * hsq);
                e += pre * (sr * sr + si * si);
                for (j = 0; j < n; j++) {
                    /* d|S|^2/dr_j = 2 q_j 2 pi h (-sin(ph) sr + cos(ph) si) */
                    const double t = pre * 2.0 * q[j] * 2.0 * PI * (-sj[j] * sr + cj[j] * si);
                    grad[3 * j] += t * h1;
                    grad[3 * j + 1] += t * h2;
                    grad[3 * j + 2] += t * h3;
                }
            }
    *energy = e;
    free(cj);
    free(sj);
}
