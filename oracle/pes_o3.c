/*
 * pes_o3.c -- CPU oracle: the O3 (ozone) ground-state surface of egrad_o3.f (permutationally invariant
 * polynomials in mixed exponential-Gaussian variables + fitted two-body term + D3(BJ) dispersion).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_real.h).  Parity UNPINNED by the reference (no vectors, cannot be compiled
 * here); pinned in tests/ by finite differences, permutation and rigid-motion invariance, the O + O2 asymptote and
 * the ozone minimum geometry, and -- where /root/reference is present -- by a test that re-reads the recurrence
 * tables and coefficients below out of the source text.
 *
 * Restatement of /root/reference/src/egrad_o3.f:
 *   egrad_o3 :33-58, pot_o3 :83-118 (Angstrom / kcal/mol inside; Cconv, Econv, Gconv, Eref)   oracle_egrad_o3_real
 *   o3pes :144-184, coord_convt_o3 :187-211                                                    o3_pes
 *   EvMorse_o3 :323-338, EvMono_o3 :343-361, EvPoly_o3 :366-445, evbas_o3 :448-500             o3_values
 *   ev2gm2_o3 :503-565 (the "modified parameters for D3(BJ)" set)                               o3_v2
 *   EvdVdR_o3 :567-612, EvdRdX_o3 :634-689, EvdMsdR_o3 :692-716, EvdMdR_o3 :730-757,
 *   EvdPdr_o3 :759-850, evdbdr_o3 :852-931                                                     o3_derivs
 *   d3disp_o3 :933-990, edisp_o3 :997-1090 (C6 fixed at 12.8, r2r4(8) a REAL*4 literal, F3)     o3_disp
 *   BLOCK DATA prmt_o3 :1096-1176                                                              O3_C, O3_A ...
 * The 67 polynomials are defined by the recurrences p(k) = p(a) p(b) - p(c1) - p(c2) - p(c3) (:375-441); the source
 * spells every one of them out twice, once for the values and once, by the product rule in the same operand order,
 * for the derivatives (:771-846: dpdr(i,k) = dpdr(i,a) p(b) + p(a) dpdr(i,b) - dpdr(i,c1) ...).  Here the recurrences
 * are ONE table (O3_REC) that both passes walk; the arithmetic, operation for operation, is the source's.
 */
#include "oracle_real.h"
#include "oracle.h"

/* p(k) = p(a) * p(b) - sum p(c): {a, b, nsub, c1, c2, c3}; a < 0 marks the four monomial sums set in o3_values */
static const signed char O3_REC[67][6] = {
    {-1, 0, 0, 0, 0, 0},   {-1, 0, 0, 0, 0, 0},   {-1, 0, 0, 0, 0, 0},   {1, 1, 2, 2, 2, 0},    {-1, 0, 0, 0, 0, 0},
    {1, 2, 3, 4, 4, 4},    {1, 3, 1, 5, 0, 0},    {1, 4, 0, 0, 0, 0},    {2, 2, 2, 7, 7, 0},    {2, 3, 1, 7, 0, 0},
    {1, 6, 1, 9, 0, 0},    {2, 4, 0, 0, 0, 0},    {3, 4, 0, 0, 0, 0},    {1, 8, 1, 11, 0, 0},   {2, 6, 1, 12, 0, 0},
    {1, 10, 1, 14, 0, 0},  {4, 4, 0, 0, 0, 0},    {4, 5, 0, 0, 0, 0},    {4, 6, 0, 0, 0, 0},    {2, 8, 1, 17, 0, 0},
    {1, 13, 3, 17, 19, 19}, {2, 10, 1, 18, 0, 0}, {1, 15, 1, 21, 0, 0},  {1, 16, 0, 0, 0, 0},   {4, 8, 0, 0, 0, 0},
    {3, 11, 1, 23, 0, 0},  {4, 10, 0, 0, 0, 0},   {1, 19, 1, 24, 0, 0},  {6, 8, 1, 23, 0, 0},   {2, 15, 1, 26, 0, 0},
    {1, 22, 1, 29, 0, 0},  {2, 16, 0, 0, 0, 0},   {3, 16, 0, 0, 0, 0},   {1, 24, 1, 31, 0, 0},  {4, 14, 0, 0, 0, 0},
    {4, 15, 0, 0, 0, 0},   {2, 19, 1, 33, 0, 0},  {3, 19, 1, 31, 0, 0},  {8, 10, 1, 32, 0, 0},  {2, 22, 1, 35, 0, 0},
    {1, 30, 1, 39, 0, 0},  {4, 16, 0, 0, 0, 0},   {4, 17, 0, 0, 0, 0},   {4, 19, 0, 0, 0, 0},   {6, 16, 0, 0, 0, 0},
    {4, 20, 0, 0, 0, 0},   {4, 21, 0, 0, 0, 0},   {4, 22, 0, 0, 0, 0},   {1, 36, 1, 43, 0, 0},  {1, 37, 2, 45, 48, 0},
    {8, 15, 1, 44, 0, 0},  {2, 30, 1, 47, 0, 0},  {1, 40, 1, 51, 0, 0},  {1, 41, 0, 0, 0, 0},   {4, 24, 0, 0, 0, 0},
    {3, 31, 1, 53, 0, 0},  {1, 43, 1, 54, 0, 0},  {10, 16, 0, 0, 0, 0},  {4, 28, 0, 0, 0, 0},   {4, 29, 0, 0, 0, 0},
    {4, 30, 0, 0, 0, 0},   {2, 36, 1, 56, 0, 0},  {3, 36, 1, 54, 0, 0},  {10, 19, 1, 53, 0, 0}, {8, 22, 1, 57, 0, 0},
    {2, 40, 1, 60, 0, 0},  {1, 52, 1, 65, 0, 0}};

/* evbas_o3 :459-497: b(j) = p(O3_BAS[j-1]); the pure powers of p(1) (two-body) are left out */
static const unsigned char O3_BAS[56] = {2,  4,  5,  7,  8,  9,  11, 12, 13, 14, 16, 17, 18, 19, 20, 21, 23, 24, 25,
                                         26, 27, 28, 29, 31, 32, 33, 34, 35, 36, 37, 38, 39, 41, 42, 43, 44, 45, 46,
                                         47, 48, 49, 50, 51, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63, 64, 65};

/* BLOCK DATA prmt_o3 :1107-1172 */
static const double O3_A = 0.83, O3_AB = 3.70, O3_RA = 1.25, O3_RB = 1.13;
static const double O3_C[56] = {
    -0.128814549305e+03, 0.104229418850e+04,  0.811983220935e+03,  -0.443324528752e+03, 0.904506805268e+03,
    -0.501026918125e+04, 0.197209669844e+05,  -0.251424247013e+05, -0.138013677810e+03, 0.169329202490e+05,
    0.352627837493e+05,  -0.334337897178e+05, 0.720500009412e+05,  0.116986232065e+05,  -0.521801943104e+04,
    -0.332486978745e+05, -0.177870892015e+05, 0.335720198273e+05,  0.268323174511e+05,  -0.933618945467e+05,
    -0.592242307973e+04, 0.287777488764e+04,  0.393607079595e+05,  -0.330171644074e+04, 0.200362806379e+05,
    -0.975981166385e+04, -0.266133829509e+05, 0.746650532707e+05,  0.120290055844e+05,  -0.464904653691e+04,
    0.244129022324e+04,  -0.273870502550e+05, -0.122995471301e+05, 0.722408057250e+04,  0.290562738593e+05,
    -0.212778140565e+05, -0.206522536997e+05, 0.263237776823e+05,  -0.370869661253e+05, -0.230814810543e+04,
    0.210942828415e+04,  -0.177122132932e+04, 0.102466183681e+05,  -0.556070962327e+03, 0.150041418580e+05,
    -0.117657568996e+05, -0.442916346835e+04, 0.140353851796e+05,  0.753137518090e+04,  -0.911889476033e+04,
    0.816658803687e+04,  0.316234339832e+04,  -0.202814305330e+04, 0.791358291948e+03,  0.115450049995e+03,
    -0.158182849802e+04};

/* ev2gm2_o3 :503-565 */
static void o3_v2(real r, real *v, real *grad)
{
    static const double alpha = 9.439784362354936e-1, beta = 1.262242998506810e0;
    static const double a[8] = {-1.488979427684798e3, 1.881435846488955e4,  -1.053475425838226e5, 2.755135591229064e5,
                                -4.277588997761775e5, 4.404104009614092e5, -2.946204062950765e5, 1.176861219078620e5};
    int k;
    *v = 0.0;
    for (k = 0; k < 8; k++) *v = *v + a[k] * exp(-alpha * pow(beta, (double)k) * (r * r));
    *v = *v * 627.509523475149e-3;
    *grad = 0.0;
    for (k = 0; k < 8; k++) *grad = *grad - 2.0 * a[k] * alpha * pow(beta, (double)k) * r * exp(-alpha * pow(beta, (double)k) * (r * r));
    *grad = *grad * 627.509523475149e-3;
}

/* d3disp_o3 :933-990 with edisp_o3 :997-1090: three O-O pairs, C6 = 12.8, BJ damping */
static void o3_disp(const real R[3], real *disp, real dispdr[3])
{
    const double autoang = 0.52917726, autokcal = 627.509541;
    const double s6 = 1.0, s8 = 2.0, a1 = 0.5299, a2 = 2.20;
    const double r2r4_O = F(2.59361680); /* r2r4(8), a REAL*4 literal in the array constructor (:1025) */
    int i;
    *disp = 0.0;
    for (i = 0; i < 3; i++) {
        const real r = R[i] / autoang;
        const real c6 = 12.8, c8 = 3.0 * c6 * r2r4_O * r2r4_O;
        const real tmp = sqrt(c8 / c6);
        const real d6 = pow(a1 * tmp + a2, 6.0), d8 = pow(a1 * tmp + a2, 8.0);
        const real r2 = r * r, r4 = r2 * r2, r5 = r4 * r, r6 = r4 * r2, r7 = r6 * r, r8 = r4 * r4;
        const real e6 = c6 / (r6 + d6), e8 = c8 / (r8 + d8);
        const real e6dr = c6 * (-6.0 * r5) / ((r6 + d6) * (r6 + d6)), e8dr = c8 * (-8.0 * r7) / ((r8 + d8) * (r8 + d8));
        *disp = *disp + (-s6 * e6 - s8 * e8) * autokcal;
        dispdr[i] = (-s6 * e6dr - s8 * e8dr) * autokcal / autoang;
    }
}

/* o3pes with igrad = 1: X(9) in Angstrom -> V (kcal/mol above the fit's zero) and dVdX(9) */
static void o3_pes(const real X[9], real *V, real dVdX[9])
{
    real R[3], rMs[3], rM[8], P[67], B[56], dMs[3], dM[3][8], dP[3][67], dVdR[3], dRdX[3][9];
    real v, v2, dv2, disp, dispdr[3];
    int i, j, k;
    R[0] = sqrt((X[3] - X[0]) * (X[3] - X[0]) + (X[4] - X[1]) * (X[4] - X[1]) + (X[5] - X[2]) * (X[5] - X[2]));
    R[1] = sqrt((X[6] - X[0]) * (X[6] - X[0]) + (X[7] - X[1]) * (X[7] - X[1]) + (X[8] - X[2]) * (X[8] - X[2]));
    R[2] = sqrt((X[3] - X[6]) * (X[3] - X[6]) + (X[4] - X[7]) * (X[4] - X[7]) + (X[5] - X[8]) * (X[5] - X[8]));
    /* EvMorse, EvMono */
    for (i = 0; i < 3; i++) rMs[i] = exp(-(R[i] - O3_RA) / O3_A - ((R[i] - O3_RB) * (R[i] - O3_RB)) / O3_AB);
    rM[0] = 1.0;
    rM[1] = rMs[2];
    rM[2] = rMs[1];
    rM[3] = rMs[0];
    rM[4] = rM[1] * rM[2];
    rM[5] = rM[1] * rM[3];
    rM[6] = rM[2] * rM[3];
    rM[7] = rM[1] * rM[6];
    /* EvPoly */
    for (k = 0; k < 67; k++) {
        const signed char *q = O3_REC[k];
        if (q[0] < 0) {
            P[k] = (k == 0) ? rM[0] : (k == 1) ? rM[1] + rM[2] + rM[3] : (k == 2) ? rM[4] + rM[5] + rM[6] : rM[7];
        } else {
            real t = P[q[0]] * P[q[1]];
            for (j = 0; j < q[2]; j++) t = t - P[q[3 + j]];
            P[k] = t;
        }
    }
    for (j = 0; j < 56; j++) B[j] = P[O3_BAS[j]];
    /* EvV */
    v = 240.486;
    for (i = 0; i < 3; i++) {
        o3_v2(R[i], &v2, &dv2);
        v = v + v2;
    }
    o3_disp(R, &disp, dispdr);
    v = v + disp;
    for (j = 0; j < 56; j++) v = v + O3_C[j] * B[j];
    *V = v;
    /* EvdVdR */
    for (i = 0; i < 3; i++) {
        o3_v2(R[i], &v2, &dv2);
        dVdR[i] = dv2;
    }
    for (i = 0; i < 3; i++) dVdR[i] = dVdR[i] + dispdr[i];
    /* EvdMsdR: only the diagonal is non-zero; EvdMdR, index i = distance */
    for (i = 0; i < 3; i++)
        dMs[i] = (-2.0 * (R[i] - O3_RB) / O3_AB - 1 / O3_A) * exp(-(R[i] - O3_RA) / O3_A - ((R[i] - O3_RB) * (R[i] - O3_RB)) / O3_AB);
    for (i = 0; i < 3; i++) {
        dM[i][0] = 0.0;
        dM[i][1] = (i == 2) ? dMs[2] : 0.0; /* dmsdr(i,3) */
        dM[i][2] = (i == 1) ? dMs[1] : 0.0; /* dmsdr(i,2) */
        dM[i][3] = (i == 0) ? dMs[0] : 0.0; /* dmsdr(i,1) */
        dM[i][4] = dM[i][1] * rM[2] + rM[1] * dM[i][2];
        dM[i][5] = dM[i][1] * rM[3] + rM[1] * dM[i][3];
        dM[i][6] = dM[i][2] * rM[3] + rM[2] * dM[i][3];
        dM[i][7] = dM[i][1] * rM[6] + rM[1] * dM[i][6];
        /* EvdPdr: the product rule on the same recurrences, operands in the source's order */
        for (k = 0; k < 67; k++) {
            const signed char *q = O3_REC[k];
            if (q[0] < 0) {
                dP[i][k] = (k == 0)   ? dM[i][0]
                           : (k == 1) ? dM[i][1] + dM[i][2] + dM[i][3]
                           : (k == 2) ? dM[i][4] + dM[i][5] + dM[i][6]
                                      : dM[i][7];
            } else {
                real t = dP[i][q[0]] * P[q[1]] + P[q[0]] * dP[i][q[1]];
                for (j = 0; j < q[2]; j++) t = t - dP[i][q[3 + j]];
                dP[i][k] = t;
            }
        }
        for (j = 0; j < 56; j++) dVdR[i] = dVdR[i] + O3_C[j] * dP[i][O3_BAS[j]];
    }
    /* EvdRdX */
    for (i = 0; i < 3; i++)
        for (j = 0; j < 9; j++) dRdX[i][j] = 0.0;
    for (k = 0; k < 3; k++) {
        dRdX[0][k] = (X[k] - X[3 + k]) / R[0];
        dRdX[0][3 + k] = -dRdX[0][k];
        dRdX[1][k] = (X[k] - X[6 + k]) / R[1];
        dRdX[1][6 + k] = -dRdX[1][k];
        dRdX[2][3 + k] = (X[3 + k] - X[6 + k]) / R[2];
        dRdX[2][6 + k] = -dRdX[2][3 + k];
    }
    for (i = 0; i < 9; i++) {
        dVdX[i] = 0.0;
        for (j = 0; j < 3; j++) dVdX[i] = dVdX[i] + dVdR[j] * dRdX[j][i];
    }
}

/* egrad_o3 + pot_o3: bohr / hartree at the interface */
void oracle_egrad_o3_real(const real *q, int natoms, int nbeads, real *V, real *dVdq, int *info)
{
    const double Cconv = 0.52917721092, Econv = 0.159360144e-2, Gconv = 0.843297564e-3, Eref = -0.19172848;
    int k, i;
    *info = 0;
    for (k = 0; k < nbeads; k++) {
        real X[9], v, dVdX[9];
        for (i = 0; i < 9; i++) X[i] = q[(long)k * 3 * natoms + i] * Cconv;
        o3_pes(X, &v, dVdX);
        V[k] = v * Econv + Eref;
        for (i = 0; i < 9; i++) dVdq[(long)k * 3 * natoms + i] = dVdX[i] * Gconv;
    }
}
