/*
 * pes_h2co.c -- CPU oracle: H2CO surface (permutationally invariant polynomial fit in Morse variables exp(-r/2)),
 * /root/reference/src/main_h2co.f90: egrad_h2co :3170-3309, edis :3311-3321, basis_h2co :3323-3335, tables
 * initialize_h2co :34-3168 (oracle/h2co_tables.h, generated from the source text), module h2co_mod.f90.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_real.h).  Parity UNPINNED by the reference (no golden vectors, cannot be
 * compiled here); pinned by the known answers of tests/test_oracle_h2co.py.
 *
 * What the source does, and the restatement keeps:
 *   - atoms C, O, H, H in bohr; six distances r1 = H-H, r2 = O-H3, r3 = C-H3, r4 = C-H4, r5 = O-H4, r6 = C-O, y = exp(-r/2);
 *   - `power` is declared real(kind=8) (h2co_mod.f90:49), so every y**power is a REAL power: ten libm pow calls per
 *     term, 1561 terms;
 *   - the 1561 coefficients and the two constants of `f=f+114.332958863-1.059892782251382E-004` carry no D exponent: read
 *     as REAL*4 (F(), oracle_real.h);
 *   - THE GRADIENT IS NUMERIC: central differences with step = 0.001 bohr on the 12 coordinates in turn (:3231-3304), the
 *     coordinate array modified in place -- x + h, then (x + h) - 2h, then ((x + h) - 2h) + h, which is what the later
 *     evaluations (and the caller, whose array it is) see;
 *   - for r(H-H) >= 8 bohr the surface is blended into an H + HCO potential (`hcopot`, util_h2co.f:34-119) that opens the
 *     file 'pot_h2co.para' with status='old' on every call (POT3IN :140-150).  The reference ships no such file: the
 *     run stops there.  The oracle returns the polynomial alone and sets info = 1.
 */
#include "oracle_real.h"
#include "oracle.h"
#define F32(x) F(x)
#include "h2co_tables.h"

static const double h2co_cof[H2CO_NCOF] = {H2CO_COF_LIST};
static const unsigned char h2co_pow[H2CO_NCOF][6] = {H2CO_POW_LIST};

/* edis :3311-3321; atoms as [xyz] triples */
static void h2co_edis(real *dis, real *edist, const real *ai, const real *aj)
{
    const real alfa = 2.0;
    *dis = sqrt((ai[0] - aj[0]) * (ai[0] - aj[0]) + (ai[1] - aj[1]) * (ai[1] - aj[1]) + (ai[2] - aj[2]) * (ai[2] - aj[2]));
    *edist = exp(-*dis / alfa);
}

/* the energy block of egrad_h2co (:3190-3229, repeated :3243-3290); cood[atom][xyz] */
static real h2co_getpot(real cood[4][3], int *far)
{
    real r[7], y[7], f, x, z, z1, z2, x2, x3, s;
    int k;
    h2co_edis(&r[1], &y[1], cood[2], cood[3]);
    h2co_edis(&r[2], &y[2], cood[1], cood[2]);
    h2co_edis(&r[3], &y[3], cood[0], cood[2]);
    h2co_edis(&r[4], &y[4], cood[0], cood[3]);
    h2co_edis(&r[5], &y[5], cood[1], cood[3]);
    h2co_edis(&r[6], &y[6], cood[0], cood[1]);
    x = r[1];
    z = x - 8.0;
    z1 = 10 - 8;
    z2 = z / z1;
    x2 = z2 * z2;
    x3 = x2 * z2;
    if (x >= 10.0)
        s = 1.0;
    else if (x <= 8.0)
        s = 0.0;
    else
        s = 10.0 * x3 - 15.0 * x3 * z2 + 6.0 * x3 * x2;
    (void)s;
    f = 0.0;
    for (k = 0; k < H2CO_NCOF; k++) {
        const unsigned char *p = h2co_pow[k];
        /* basis_h2co :3329-3332 */
        const real bas = (pow(y[1], (double)p[0]) * pow(y[6], (double)p[5])) *
                         (pow(y[2], (double)p[1]) * pow(y[3], (double)p[2]) * pow(y[4], (double)p[3]) * pow(y[5], (double)p[4]) +
                          pow(y[2], (double)p[4]) * pow(y[3], (double)p[3]) * pow(y[4], (double)p[2]) * pow(y[5], (double)p[1]));
        f = f + h2co_cof[k] * bas;
    }
    f = f + F(114.332958863) - F(1.059892782251382E-004);
    if (x >= 8.0) *far = 1; /* hcopot: the reference stops here (no pot_h2co.para); f stays the polynomial */
    return f;
}

void oracle_egrad_h2co_real(const real *q, int natoms, int nbeads, real *V, real *dVdq, int *info)
{
    const real step = 0.001;
    int b, i, j, k, far = 0;
    for (b = 0; b < nbeads; b++) {
        real cood[4][3], e_upper = 0.0, e_lower = 0.0;
        const real *qb = q + (long)b * 3 * natoms;
        real *gb = dVdq + (long)b * 3 * natoms;
        for (i = 0; i < 4; i++)
            for (j = 0; j < 3; j++) cood[i][j] = qb[3 * i + j];
        V[b] = h2co_getpot(cood, &far);
        for (i = 0; i < 4; i++)
            for (j = 0; j < 3; j++) {
                for (k = 1; k <= 2; k++) {
                    real e;
                    if (k == 1)
                        cood[i][j] = cood[i][j] + step;
                    else
                        cood[i][j] = cood[i][j] - 2.0 * step;
                    e = h2co_getpot(cood, &far);
                    if (k == 1)
                        e_upper = e;
                    else
                        e_lower = e;
                }
                cood[i][j] = cood[i][j] + step;
                gb[3 * i + j] = (e_upper - e_lower) / (2.0 * step);
            }
        for (i = 4; i < natoms; i++)
            for (j = 0; j < 3; j++) gb[3 * i + j] = 0.0;
    }
    *info = far;
}

/* the energy alone (known-answer tests, minimisations) */
void oracle_h2co_energy_real(const real *q12, real *V, int *far)
{
    real cood[4][3];
    int i, j;
    *far = 0;
    for (i = 0; i < 4; i++)
        for (j = 0; j < 3; j++) cood[i][j] = q12[3 * i + j];
    *V = h2co_getpot(cood, far);
}
