// pes_oh3.cuh -- Schatz-Elgersma OH + H2 surface, one thread per image, FP64.
//
// Replaces /root/reference/src/egrad_oh3.f (egrad_oh3 :33-188, pot_oh3 :250-395,
// V3POT_oh3 :400-480, V4POT_oh3 :482-519, VH2O_oh3 :521-590; constants :621-637, all D0).
// Restructured for the GPU: one exp per Morse/LEPS leg, 1-tanh and sech^2 from one exp,
// the dead POTLIB wrapper calls on a zero CART are gone.  Atom order O,H1,H2,H3;
// R = (OH1, OH2, OH3, H1H2, H1H3, H2H3).
#pragma once
#include "crcl_common.cuh"

namespace crcl {
namespace oh3 {

// BLOCK DATA PTPACM_oh3 (egrad_oh3.f:621-637).  Scalars, because namespace-scope constexpr
// arrays are not addressable from device code.
constexpr double DE0 = 0.148201, DE1 = 0.0275690, DE2 = 0.151548;
constexpr double BETA0 = 1.260580, BETA1 = 0.924180, BETA2 = 1.068620;
constexpr double RE0 = 1.863300, RE1 = 2.907700, RE2 = 1.428600;
constexpr double SATO = 0.10;
constexpr double GAM0 = 2.399700, GAM1 = 1.058350, GAM2 = 2.399700;
constexpr double REOH = 1.808090, REHH = 2.861590;
constexpr double CON0 = -.0015920, CON1 = 0.026963, CON2 = 0.0014689, CON3 = 0.080011,
                 CON4 = 0.085816, CON5 = -0.063179, CON6 = 0.101380;
constexpr double ALP0 = 4.773, ALP1 = 7.14, ALP2 = 2.938, ALP3 = 5.28;
constexpr double CLAM0 = 0.10, CLAM1 = 0.10, CLAM2 = 0.20, CLAM3 = 0.03;
constexpr double ACON0 = 0.10, ACON1 = 0.009;

CRCL_HD __forceinline__ void morse(double D, double B, double T, double r, double& V, double& dV)
{
    const double x = CRCL_EXP(-B * (r - T));
    const double u = 1.0 - x;
    V += D * u * u;
    dV += 2.0 * B * D * u * x;
}

// H2O-like three-body term on (r_OHa, r_OHb, r_HaHb) (VH2O_oh3); adds to V and the three derivatives in that
// argument order (the reference's DEDR(2)<->DEDR(3) swap folded in).  D[3] mirrors COMMON /POT2CM_oh3/ DEDR(1:3) in the
// caller's argument order: where Q(I) = 0 (0.5 gamma (R - Re) >= 43, R_OH > 37 a0) the reference skips the assignment
// and the value of the PREVIOUS routine (V3POT_oh3, or the first VH2O_oh3 call) stays in DEDR(I), is swapped with the
// others and added to the gradient (egrad_oh3.f:570-583) -- reproduced, so parity holds there too.
CRCL_HD __forceinline__ void vh2o(double roha, double rohb, double rhh, double& V, double& da,
                                     double& db, double& dhh, double D[3])
{
    const double S1 = roha - REOH, S3 = rohb - REOH, S2 = rhh - REHH;
    const double S[3] = {S1, S2, S3};
    constexpr double GAM[3] = {GAM0, GAM1, GAM2};
    double Q[3], L[3];  // L = DQ/Q
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double X = 0.5 * GAM[i] * S[i];
        if (X < 43.0) {
            double omt, ms2;
            one_minus_tanh(X, omt, ms2);
            Q[i] = omt;
            L[i] = -0.5 * GAM[i] * (2.0 - omt);   // DQ/Q = -sech^2 / (1 - tanh) = -(1 + tanh)
        } else {  // reference sets Q=0 and leaves DEDR(I) stale; E is then exactly 0
            Q[i] = 0.0;
            L[i] = 0.0;
        }
    }
    const double P = CON0 + CON1 * (S1 + S3) + CON2 * S2 + 0.5 * CON3 * (S1 * S1 + S3 * S3) +
                     0.5 * CON4 * S2 * S2 + CON5 * S2 * (S1 + S3) + CON6 * S1 * S3;
    const double E = Q[0] * Q[1] * Q[2] * P;
    const double DP1 = CON1 + CON3 * S1 + CON5 * S2 + CON6 * S3;
    const double DP2 = CON2 + CON4 * S2 + CON5 * (S1 + S3);
    const double DP3 = CON1 + CON3 * S3 + CON5 * S2 + CON6 * S1;
    V += E;
    // DEDR(1:3) before the swap is indexed like S: (r_OHa, r_HH, r_OHb)
    const double iP = CRCL_RCP(P);
    const double p0 = (Q[0] == 0.0) ? D[0] : E * (L[0] + DP1 * iP);
    const double p1 = (Q[1] == 0.0) ? D[1] : E * (L[1] + DP2 * iP);
    const double p2 = (Q[2] == 0.0) ? D[2] : E * (L[2] + DP3 * iP);
    D[0] = p0;
    D[1] = p2;   // DEDR(2) <-> DEDR(3)
    D[2] = p1;
    da += D[0];
    db += D[1];
    dhh += D[2];
}

CRCL_HD __forceinline__ void pot(const double R[6], double& V, double dV[6])
{
    V = 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++) dV[i] = 0.0;
    morse(DE0, BETA0, RE0, R[0], V, dV[0]);
    morse(DE1, BETA1, RE1, R[3], V, dV[3]);
    morse(DE1, BETA1, RE1, R[4], V, dV[4]);
    double DEDR[3];   // COMMON /POT2CM_oh3/ DEDR(1:3) as the routines below leave it (see vh2o)
    // three-body LEPS on (OH2, OH3, H2H3)  (V3POT_oh3)
    {
        constexpr double Z = SATO, ZPO = 1.0 + Z, OP3Z = 1.0 + 3.0 * Z, TOP3Z = 2.0 * OP3Z,
                         ZP3 = Z + 3.0, TZP3 = 2.0 * ZP3;
        const double r[3] = {R[1], R[2], R[5]};
        constexpr double de[3] = {DE0, DE0, DE2};
        constexpr double be[3] = {BETA0, BETA0, BETA2};
        constexpr double re[3] = {RE0, RE0, RE2};
        double X[3], EX[3], S = 0.0, E = 0.0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double DO4Z = de[i] / 4.0 / ZPO;
            X[i] = CRCL_EXP(-be[i] * (r[i] - re[i]));
            E += DO4Z * (ZP3 * X[i] - TOP3Z) * X[i];
            EX[i] = DO4Z * (OP3Z * X[i] - TZP3) * X[i];
            S += EX[i];
        }
        double RAD, iRAD;
        sqrt_rsqrt(sqr(EX[0] - EX[1]) + sqr(EX[1] - EX[2]) + sqr(EX[2] - EX[0]), RAD, iRAD);
        constexpr double RS2 = 0.70710678118654752440;  // 1/sqrt(2)
        V += E - RAD * RS2;
        const int idx[3] = {1, 2, 5};
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double B = be[i] * (de[i] / 4.0 / ZPO) * 2.0;
            DEDR[i] = B * X[i] * ((3.0 * EX[i] - S) * RS2 * (OP3Z * X[i] - ZP3) * iRAD - ZP3 * X[i] + OP3Z);
            dV[idx[i]] += DEDR[i];
        }
    }
    vh2o(R[0], R[1], R[3], V, dV[0], dV[1], dV[3], DEDR);
    vh2o(R[0], R[2], R[4], V, dV[0], dV[2], dV[4], DEDR);
    // four-body term on (OH2, OH3, H1H2, H1H3)  (V4POT_oh3; A=ALP, C=CLAM, COF=ACON)
    {
        const double r[4] = {R[1], R[2], R[3], R[4]};
        const double T1 = ACON0 * CRCL_EXP(-CLAM0 * (sqr(r[0] - ALP0) + sqr(r[1] - ALP0)) -
                                        CLAM2 * (sqr(r[2] - ALP2) + sqr(r[3] - ALP2)));
        const double T2 = ACON1 * CRCL_EXP(-CLAM1 * (sqr(r[0] - ALP1) + sqr(r[1] - ALP1)) -
                                        CLAM3 * (sqr(r[2] - ALP3) + sqr(r[3] - ALP3)));
        V += T1 + T2;
        dV[1] += -2.0 * (T1 * CLAM0 * (r[0] - ALP0) + T2 * CLAM1 * (r[0] - ALP1));
        dV[2] += -2.0 * (T1 * CLAM0 * (r[1] - ALP0) + T2 * CLAM1 * (r[1] - ALP1));
        dV[3] += -2.0 * (T1 * CLAM2 * (r[2] - ALP2) + T2 * CLAM3 * (r[2] - ALP3));
        dV[4] += -2.0 * (T1 * CLAM2 * (r[3] - ALP2) + T2 * CLAM3 * (r[3] - ALP3));
    }
    V += -2.0 * DE1 + DE2;
}

}  // namespace oh3

struct PesOH3 {
    static constexpr int NATOMS = 4;
    static constexpr int ID = CRCL_PES_OH3;
    // one thread per image in the trajectory kernels (LANES cooperating threads per bead)
    static constexpr int LANES = 1;
    static constexpr int NOWN = 3 * NATOMS;
    CRCL_HD static __forceinline__ int owned(int, int k) { return k; }
    template <class QF>
    CRCL_HD static __forceinline__ int eval_coop(QF qf, int, unsigned, double& V, double* gown)
    {
        double x[NOWN];
#pragma unroll
        for (int c = 0; c < NOWN; c++) x[c] = qf(c);
        return eval(x, V, gown);
    }
    CRCL_HD static __forceinline__ int eval(const double* __restrict__ q, double& V,
                                               double* __restrict__ g)
    {
        constexpr int PA[6] = {0, 0, 0, 1, 1, 2};
        constexpr int PB[6] = {1, 2, 3, 2, 3, 3};
        double vec[6][3], R[6], iR[6], dV[6];
#pragma unroll
        for (int m = 0; m < 6; m++) {
#pragma unroll
            for (int d = 0; d < 3; d++) vec[m][d] = q[3 * PB[m] + d] - q[3 * PA[m] + d];
            sqrt_rsqrt(vec[m][0] * vec[m][0] + vec[m][1] * vec[m][1] + vec[m][2] * vec[m][2], R[m], iR[m]);
        }
        oh3::pot(R, V, dV);
#pragma unroll
        for (int d = 0; d < 12; d++) g[d] = 0.0;
#pragma unroll
        for (int m = 0; m < 6; m++) {
            const double f = dV[m] * iR[m];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                g[3 * PB[m] + d] += f * vec[m][d];
                g[3 * PA[m] + d] -= f * vec[m][d];
            }
        }
        return 0;
    }
};

}  // namespace crcl
