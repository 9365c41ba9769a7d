/*
 * pes_oh3.c -- CPU oracle: Schatz-Elgersma OH + H2 potential energy surface.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_real.h).  Parity UNPINNED by the reference
 * (no golden vectors, cannot be compiled here); pinned by finite differences, the
 * asymptote E(OH(re)+H2(re)) = 0 and the literature barrier in tests/.
 *
 * Literal restatement of /root/reference/src/egrad_oh3.f:
 *   egrad_oh3  :33-188   oracle_egrad_oh3_real (one bead per call, as gradient.f90:191
 *                        calls it with Nbeads=1; the reference zeroes the whole dVdq
 *                        inside its bead loop at :125)
 *   pot_oh3    :250-395  oh3_pot      (the POTLIB CARTOU/CARTTOR/RTOCART/DEDCOU calls work
 *                        on an all-zero CART and their results are unused: omitted)
 *   V3POT_oh3  :400-480  oh3_v3pot
 *   V4POT_oh3  :482-519  oh3_v4pot
 *   VH2O_oh3   :521-590  oh3_vh2o
 *   constants  :621-637  (BLOCK DATA PTPACM_oh3, all D0 -> clean doubles)
 * COMMON /POT2CM_oh3/ RSEND(4),ENERGY,DEDR(4) is the struct oh3_pot2.
 */
#include "oracle_real.h"
#include "oracle.h"

typedef struct {
    real RSEND[4];
    real ENERGY;
    real DEDR[4];
} oh3_pot2;

/* BLOCK DATA PTPACM_oh3 (egrad_oh3.f:621-637) */
static const double OH3_DE[3] = {0.148201, 0.0275690, 0.151548};
static const double OH3_BETA[3] = {1.260580, 0.924180, 1.068620};
static const double OH3_RE[3] = {1.863300, 2.907700, 1.428600};
static const double OH3_SATO = 0.10;
static const double OH3_GAM[3] = {2.399700, 1.058350, 2.399700};
static const double OH3_REOH = 1.808090;
static const double OH3_REHH = 2.861590;
static const double OH3_CON[7] = {-.0015920, 0.026963, 0.0014689, 0.080011,
                                  0.085816,  -0.063179, 0.101380};
static const double OH3_ALP[4] = {4.773, 7.14, 2.938, 5.28};
static const double OH3_CLAM[4] = {0.10, 0.10, 0.20, 0.03};
static const double OH3_ACON[2] = {0.10, 0.009};

/* statement functions VMOR / DVMOR (egrad_oh3.f:294-295) */
static real oh3_vmor(real D, real B, real T, real RR)
{
    real u = 1.0 - exp(-B * (RR - T));
    return D * (u * u);
}
static real oh3_dvmor(real D, real B, real T, real RR)
{
    return 2.0 * B * D * (1.0 - exp(-B * (RR - T))) * exp(-B * (RR - T));
}

/* ---- V3POT_oh3 (egrad_oh3.f:400-480): three-body LEPS ---- */
static void oh3_v3pot(oh3_pot2 *c)
{
    real DE[3], BETA[3], RE[3], Z[3], ZPO[3], OP3Z[3], ZP3[3], TZP3[3], TOP3Z[3], DO4Z[3],
        B[3], X[3], COUL[3], EXCH[3];
    real RAD, S;
    const real *R = c->RSEND;
    int i;
    DE[0] = OH3_DE[0];
    BETA[0] = OH3_BETA[0];
    RE[0] = OH3_RE[0];
    DE[1] = OH3_DE[0];
    BETA[1] = OH3_BETA[0];
    RE[1] = OH3_RE[0];
    DE[2] = OH3_DE[2];
    BETA[2] = OH3_BETA[2];
    RE[2] = OH3_RE[2];
    for (i = 0; i < 3; i++) {
        Z[i] = OH3_SATO;
        ZPO[i] = 1.0 + Z[i];
        OP3Z[i] = 1.0 + 3.0 * Z[i];
        TOP3Z[i] = 2.0 * OP3Z[i];
        ZP3[i] = Z[i] + 3.0;
        TZP3[i] = 2.0 * ZP3[i];
        DO4Z[i] = DE[i] / 4.0 / ZPO[i];
        B[i] = BETA[i] * DO4Z[i] * 2.0;
    }
    c->ENERGY = 0.0;
    for (i = 0; i < 3; i++) {
        X[i] = exp(-BETA[i] * (R[i] - RE[i]));
        COUL[i] = DO4Z[i] * (ZP3[i] * X[i] - TOP3Z[i]) * X[i];
        EXCH[i] = DO4Z[i] * (OP3Z[i] * X[i] - TZP3[i]) * X[i];
        c->ENERGY = c->ENERGY + COUL[i];
    }
    RAD = sqrt((EXCH[0] - EXCH[1]) * (EXCH[0] - EXCH[1]) +
               (EXCH[1] - EXCH[2]) * (EXCH[1] - EXCH[2]) +
               (EXCH[2] - EXCH[0]) * (EXCH[2] - EXCH[0]));
    c->ENERGY = c->ENERGY - RAD / sqrt((real)2.0);
    S = EXCH[0] + EXCH[1] + EXCH[2];
    for (i = 0; i < 3; i++) {
        c->DEDR[i] = B[i] * X[i] *
                     ((3.0 * EXCH[i] - S) / sqrt((real)2.0) * (OP3Z[i] * X[i] - ZP3[i]) / RAD -
                      ZP3[i] * X[i] + OP3Z[i]);
    }
}

/* ---- V4POT_oh3 (egrad_oh3.f:482-519): A=ALP, C=CLAM, COF=ACON by COMMON position ---- */
static void oh3_v4pot(oh3_pot2 *c)
{
    const real *R = c->RSEND;
    const double *A = OH3_ALP, *C = OH3_CLAM, *COF = OH3_ACON;
    real T1, T2;
    T1 = exp(-C[0] * ((R[0] - A[0]) * (R[0] - A[0])) - C[0] * ((R[1] - A[0]) * (R[1] - A[0])) -
             C[2] * ((R[2] - A[2]) * (R[2] - A[2])) - C[2] * ((R[3] - A[2]) * (R[3] - A[2]))) *
         COF[0];
    T2 = exp(-C[1] * ((R[0] - A[1]) * (R[0] - A[1])) - C[1] * ((R[1] - A[1]) * (R[1] - A[1])) -
             C[3] * ((R[2] - A[3]) * (R[2] - A[3])) - C[3] * ((R[3] - A[3]) * (R[3] - A[3]))) *
         COF[1];
    c->ENERGY = T1 + T2;
    c->DEDR[0] = -2.0 * (T1 * C[0] * (R[0] - A[0]) + T2 * C[1] * (R[0] - A[1]));
    c->DEDR[1] = -2.0 * (T1 * C[0] * (R[1] - A[0]) + T2 * C[1] * (R[1] - A[1]));
    c->DEDR[2] = -2.0 * (T1 * C[2] * (R[2] - A[2]) + T2 * C[3] * (R[2] - A[3]));
    c->DEDR[3] = -2.0 * (T1 * C[2] * (R[3] - A[2]) + T2 * C[3] * (R[3] - A[3]));
}

/* ---- VH2O_oh3 (egrad_oh3.f:521-590) ----
 * DEDR(I) is left untouched (stale COMMON value) when Q(I)==0, as in the reference. */
static void oh3_vh2o(oh3_pot2 *c)
{
    const real XMAX1 = 15.0, XMAX2 = 43.0;
    const double *C = OH3_CON;
    const real *R = c->RSEND;
    real S[3], Q[3], DQ[3], X[3], DP[3];
    real P, E, TEMP;
    int i;
    S[0] = R[0] - OH3_REOH;
    S[2] = R[1] - OH3_REOH;
    S[1] = R[2] - OH3_REHH;
    for (i = 0; i < 3; i++) {
        X[i] = 0.5 * OH3_GAM[i] * S[i];
        Q[i] = 1.0 - tanh(X[i]);
        if (!(X[i] < XMAX1)) {
            if (!(X[i] < XMAX2)) {
                Q[i] = 0.0;
                DQ[i] = 0.0;
                continue;
            }
            Q[i] = 2.0 / (1.0 + exp(2.0 * X[i])); /* TANLG */
        }
        DQ[i] = -0.5 * OH3_GAM[i] / (cosh(X[i]) * cosh(X[i]));
    }
    P = C[0] + C[1] * (S[0] + S[2]) + C[2] * S[1] + 0.5 * C[3] * (S[0] * S[0] + S[2] * S[2]) +
        0.5 * C[4] * S[1] * S[1] + C[5] * S[1] * (S[0] + S[2]) + C[6] * S[0] * S[2];
    E = Q[0] * Q[1] * Q[2] * P;
    c->ENERGY = E;
    DP[0] = C[1] + C[3] * S[0] + C[5] * S[1] + C[6] * S[2];
    DP[1] = C[2] + C[4] * S[1] + C[5] * (S[0] + S[2]);
    DP[2] = C[1] + C[3] * S[2] + C[5] * S[1] + C[6] * S[0];
    for (i = 0; i < 3; i++) {
        real TRM1;
        if (Q[i] == 0.0) continue;
        TRM1 = DQ[i] / Q[i];
        c->DEDR[i] = E * (TRM1 + (DP[i] / P));
    }
    TEMP = c->DEDR[1];
    c->DEDR[1] = c->DEDR[2];
    c->DEDR[2] = TEMP;
}

/* ---- pot_oh3 (egrad_oh3.f:250-395); R(1..6) = OH1,OH2,OH3,H1H2,H1H3,H2H3 ---- */
static void oh3_pot(const real R[6], real *VTOT, real DVDR[6])
{
    oh3_pot2 c;
    const double *DE = OH3_DE, *BETA = OH3_BETA, *RE = OH3_RE;
    int i;
    for (i = 0; i < 4; i++) {
        c.RSEND[i] = 0.0;
        c.DEDR[i] = 0.0;
    }
    c.ENERGY = 0.0;
    for (i = 0; i < 6; i++) DVDR[i] = 0.0;
    *VTOT = oh3_vmor(DE[0], BETA[0], RE[0], R[0]) + oh3_vmor(DE[1], BETA[1], RE[1], R[3]) +
            oh3_vmor(DE[1], BETA[1], RE[1], R[4]);
    DVDR[0] = DVDR[0] + oh3_dvmor(DE[0], BETA[0], RE[0], R[0]);
    DVDR[3] = DVDR[3] + oh3_dvmor(DE[1], BETA[1], RE[1], R[3]);
    DVDR[4] = DVDR[4] + oh3_dvmor(DE[1], BETA[1], RE[1], R[4]);
    /* three-body LEPS on (OH2, OH3, H2H3) */
    c.RSEND[0] = R[1];
    c.RSEND[1] = R[2];
    c.RSEND[2] = R[5];
    oh3_v3pot(&c);
    *VTOT = *VTOT + c.ENERGY;
    DVDR[1] = DVDR[1] + c.DEDR[0];
    DVDR[2] = DVDR[2] + c.DEDR[1];
    DVDR[5] = DVDR[5] + c.DEDR[2];
    /* H2O part for H1,H2 */
    c.RSEND[0] = R[0];
    c.RSEND[1] = R[1];
    c.RSEND[2] = R[3];
    oh3_vh2o(&c);
    *VTOT = *VTOT + c.ENERGY;
    DVDR[0] = DVDR[0] + c.DEDR[0];
    DVDR[1] = DVDR[1] + c.DEDR[1];
    DVDR[3] = DVDR[3] + c.DEDR[2];
    /* H2O part for H1,H3 */
    c.RSEND[0] = R[0];
    c.RSEND[1] = R[2];
    c.RSEND[2] = R[4];
    oh3_vh2o(&c);
    *VTOT = *VTOT + c.ENERGY;
    DVDR[0] = DVDR[0] + c.DEDR[0];
    DVDR[2] = DVDR[2] + c.DEDR[1];
    DVDR[4] = DVDR[4] + c.DEDR[2];
    /* four-body part */
    c.RSEND[0] = R[1];
    c.RSEND[1] = R[2];
    c.RSEND[2] = R[3];
    c.RSEND[3] = R[4];
    oh3_v4pot(&c);
    *VTOT = *VTOT + c.ENERGY;
    DVDR[1] = DVDR[1] + c.DEDR[0];
    DVDR[2] = DVDR[2] + c.DEDR[1];
    DVDR[3] = DVDR[3] + c.DEDR[2];
    DVDR[4] = DVDR[4] + c.DEDR[3];
    *VTOT = *VTOT - 2.0 * DE[1] + DE[2];
}

void oracle_oh3_pot_real(const real R[6], real *V, real dVdR[6]) { oh3_pot(R, V, dVdR); }

/* ---- egrad_oh3 (egrad_oh3.f:33-188), atom order O,H1,H2,H3.  Loops beads the way
 * gradient.f90:191 + verlet.f90:772-777 do (one call per bead). ---- */
void oracle_egrad_oh3_real(const real *q, int natoms, int nbeads, real *V, real *dVdq, int *info)
{
    static const int PA[6] = {0, 0, 0, 1, 1, 2}; /* first atom of distance R(m) */
    static const int PB[6] = {1, 2, 3, 2, 3, 3}; /* second atom: vector = q(PB)-q(PA) */
    int k, m, d;
    *info = 0;
    for (k = 0; k < nbeads; k++) {
        const real *qk = q + (long)k * 3 * natoms;
        real *gk = dVdq + (long)k * 3 * natoms;
        real vec[6][3], R[6], dVdr[6];
        for (m = 0; m < 6; m++) {
            for (d = 0; d < 3; d++) vec[m][d] = qk[3 * PB[m] + d] - qk[3 * PA[m] + d];
            R[m] = sqrt(vec[m][0] * vec[m][0] + vec[m][1] * vec[m][1] + vec[m][2] * vec[m][2]);
        }
        oh3_pot(R, &V[k], dVdr);
        for (d = 0; d < 3 * natoms; d++) gk[d] = 0.0;
        /* accumulation order of egrad_oh3.f:128-182 */
        for (d = 0; d < 3; d++) {
            gk[3 * 0 + d] = gk[3 * 0 + d] - dVdr[0] * vec[0][d] / R[0];
            gk[3 * 0 + d] = gk[3 * 0 + d] - dVdr[1] * vec[1][d] / R[1];
            gk[3 * 0 + d] = gk[3 * 0 + d] - dVdr[2] * vec[2][d] / R[2];
            gk[3 * 1 + d] = gk[3 * 1 + d] + dVdr[0] * vec[0][d] / R[0];
            gk[3 * 1 + d] = gk[3 * 1 + d] - dVdr[3] * vec[3][d] / R[3];
            gk[3 * 1 + d] = gk[3 * 1 + d] - dVdr[4] * vec[4][d] / R[4];
            gk[3 * 2 + d] = gk[3 * 2 + d] + dVdr[1] * vec[1][d] / R[1];
            gk[3 * 2 + d] = gk[3 * 2 + d] + dVdr[3] * vec[3][d] / R[3];
            gk[3 * 2 + d] = gk[3 * 2 + d] - dVdr[5] * vec[5][d] / R[5];
            gk[3 * 3 + d] = gk[3 * 3 + d] + dVdr[2] * vec[2][d] / R[2];
            gk[3 * 3 + d] = gk[3 * 3 + d] + dVdr[4] * vec[4][d] / R[4];
            gk[3 * 3 + d] = gk[3 * 3 + d] + dVdr[5] * vec[5][d] / R[5];
        }
    }
}
