/*
 * pes_brh2.c -- CPU oracle: BrH2 DIM-3C potential energy surface (Clary 1982 / Last-Baer 1981).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_real.h).  Parity UNPINNED by the reference (no golden
 * vectors, cannot be compiled here); pinned in tests/ by finite differences, the zero of energy
 * (Br + H2(re) = 0), the H + HBr asymptote DHH - DHX, the energy-difference / integrated-gradient
 * identity, the H <-> H exchange symmetry, and the device kernel's independent Jacobi diagonalisation.
 *
 * Literal restatement of /root/reference/src/egrad_brh2.f:
 *   egrad_brh2       :33-80    oracle_egrad_brh2_real  (atom 2 is Br: R(1)=r12, R(2)=r13, R(3)=r23)
 *   initialize_brh2  :84-186   brh2_init   (G <- G/RHX, XIB, DENOM, RT34, RT38; SAVEd locals)
 *   POT_brh2 (entry) :188-440  brh2_pot
 *   VHX              :442-463  brh2_vhx
 *   RSP              :485-558  brh2_rsp    (EISPACK driver; MATZ = 1 branch only, as POT calls it)
 *   TRED3            :560-671  brh2_tred3  (packed Householder tridiagonalisation)
 *   TQL2_br          :808-978  brh2_tql2   (QL implicit; MACHEP = 2**(-37), 30 iterations)
 *   TRBAK3           :980-1059 brh2_trbak3 (back-transformation of the eigenvectors)
 *   BLOCK DATA       :465-483  constants (all D0 literals -> clean doubles)
 * All index arithmetic is kept 1-based through the A_()/Z_() accessors so that it reads like the
 * source.  The EISPACK iteration control stays in plain double (it is data-dependent control flow).
 */
#include "oracle_real.h"
#include "oracle.h"

/* BLOCK DATA PTPARM (egrad_brh2.f:465-483) */
static const double BR_EPS = 1.e-6, BR_ONE = 1.0;
static const double BR_RHH = 1.4016, BR_AHH = 1.0291, BR_DHH = 0.17447, BR_BHH = 0.018;
static const double BR_XIH = 1.0, BR_XIX = 1.6, BR_G0 = 0.22, BR_ALFW = 1.0;
static const double BR_RHX = 2.673, BR_AHX = 0.957, BR_DHX = 0.1439, BR_BHX = 0.012;
static const double BR_ETAHH = 0.393764, BR_ETA3S = 0.322, BR_ETA1P = 0.20286, BR_ETA3P = 0.1771;

typedef struct {
    double XIB, C3, DENOM, G, RT34, RT38;
} brh2_consts;

/* initialize_brh2 (egrad_brh2.f:149-159); read_pes calls it once, so G is divided by RHX once */
static brh2_consts brh2_init(void)
{
    brh2_consts c;
    double T;
    c.XIB = 0.5 * (BR_XIH + BR_XIX);
    T = c.XIB * BR_RHX;
    c.C3 = 1.0 / 3.0;
    c.DENOM = 2.0 * BR_RHX * (1.0 + T * (1.0 + c.C3 * T)) * exp(-T);
    c.G = BR_G0 / BR_RHX;
    c.RT34 = 0.25 * sqrt(3.0);
    c.RT38 = 0.5 * c.RT34;
    return c;
}

static real br_sign(real a, real b) { return (b >= 0.0) ? fabs(a) : -fabs(a); }

/* VHX (egrad_brh2.f:442-463): HBr singlet/triplet sigma and pi curves and derivatives */
static void brh2_vhx(real R, real *V1S, real *V3S, real *V1P, real *V3P, real *DV1S, real *DV3S, real *DV1P, real *DV3P)
{
    real RDIF = R - BR_RHX;
    real RDIF2 = RDIF * RDIF;
    real EX1 = exp(-BR_AHX * RDIF);
    real EX2 = exp(-BR_BHX * RDIF2 * RDIF);
    real T1 = BR_DHX * EX1 * EX2;
    real V1, DV1, T2;
    *V1S = T1 * (EX1 - 2.0);
    V1 = T1 * (EX1 + 2.0);
    *V3S = BR_ETA3S * V1;
    *V1P = BR_ETA1P * V1;
    *V3P = BR_ETA3P * V1;
    T1 = 2.0 * BR_AHX * T1;
    T2 = 3.0 * BR_BHX * RDIF2;
    *DV1S = T1 * (1.0 - EX1) - T2 * *V1S;
    DV1 = -T1 * (1.0 + EX1) - T2 * V1;
    *DV3S = BR_ETA3S * DV1;
    *DV1P = BR_ETA1P * DV1;
    *DV3P = BR_ETA3P * DV1;
}

#define A_(i) A[(i) - 1]
#define D_(i) D[(i) - 1]
#define E_(i) E[(i) - 1]
#define E2_(i) E2[(i) - 1]
#define Z_(k, j) Z[((j) - 1) * 4 + ((k) - 1)] /* Z(NM,N) column-major, NM = 4 */

/* TRED3 (egrad_brh2.f:560-671) */
static void brh2_tred3(int N, real *A, real *D, real *E, real *E2)
{
    int I, II, J, K, L, IZ, JK;
    real F, G, H, HH, SCALE;
    for (II = 1; II <= N; II++) {
        I = N + 1 - II;
        L = I - 1;
        IZ = (I * L) / 2;
        H = 0.0;
        SCALE = 0.0;
        if (L >= 1) {
            for (K = 1; K <= L; K++) {
                IZ = IZ + 1;
                D_(K) = A_(IZ);
                SCALE = SCALE + fabs(D_(K));
            }
        }
        if (L < 1 || SCALE == 0.0) {
            E_(I) = 0.0;
            E2_(I) = 0.0;
        } else {
            for (K = 1; K <= L; K++) {
                D_(K) = D_(K) / SCALE;
                H = H + D_(K) * D_(K);
            }
            E2_(I) = SCALE * SCALE * H;
            F = D_(L);
            G = -br_sign(sqrt(H), F);
            E_(I) = SCALE * G;
            H = H - F * G;
            D_(L) = F - G;
            A_(IZ) = SCALE * D_(L);
            if (L != 1) {
                F = 0.0;
                for (J = 1; J <= L; J++) {
                    G = 0.0;
                    JK = (J * (J - 1)) / 2;
                    for (K = 1; K <= L; K++) {
                        JK = JK + 1;
                        if (K > J) JK = JK + K - 2;
                        G = G + A_(JK) * D_(K);
                    }
                    E_(J) = G / H;
                    F = F + E_(J) * D_(J);
                }
                HH = F / (H + H);
                JK = 0;
                for (J = 1; J <= L; J++) {
                    F = D_(J);
                    G = E_(J) - HH * F;
                    E_(J) = G;
                    for (K = 1; K <= J; K++) {
                        JK = JK + 1;
                        A_(JK) = A_(JK) - F * E_(K) - G * D_(K);
                    }
                }
            }
        }
        D_(I) = A_(IZ + 1);
        A_(IZ + 1) = SCALE * sqrt(H);
    }
}

/* TQL2_br (egrad_brh2.f:808-978); returns IERR */
static int brh2_tql2(int N, real *D, real *E, real *Z)
{
    int I, II, J, K, L, L1, M, MML;
    real B, C, F, G, H, P, R, S;
    const double MACHEP = 7.275957614183426e-12; /* TWO**(-37) */
    if (N == 1) return 0;
    for (I = 2; I <= N; I++) E_(I - 1) = E_(I);
    F = 0.0;
    B = 0.0;
    E_(N) = 0.0;
    for (L = 1; L <= N; L++) {
        J = 0;
        H = MACHEP * (fabs(D_(L)) + fabs(E_(L)));
        if (B < H) B = H;
        for (M = L; M <= N; M++)
            if (fabs(E_(M)) <= B) break; /* E(N) = 0 always ends the search */
        if (M != L) {
            do {
                if (J == 30) return L;
                J = J + 1;
                L1 = L + 1;
                G = D_(L);
                P = (D_(L1) - G) / (2.0 * E_(L));
                R = sqrt(P * P + 1.0);
                D_(L) = E_(L) / (P + br_sign(R, P));
                H = G - D_(L);
                for (I = L1; I <= N; I++) D_(I) = D_(I) - H;
                F = F + H;
                P = D_(M);
                C = 1.0;
                S = 0.0;
                MML = M - L;
                for (II = 1; II <= MML; II++) {
                    I = M - II;
                    G = C * E_(I);
                    H = C * P;
                    if (!(fabs(P) < fabs(E_(I)))) {
                        C = E_(I) / P;
                        R = sqrt(C * C + 1.0);
                        E_(I + 1) = S * P * R;
                        S = C / R;
                        C = 1.0 / R;
                    } else {
                        C = P / E_(I);
                        R = sqrt(C * C + 1.0);
                        E_(I + 1) = S * E_(I) * R;
                        S = 1.0 / R;
                        C = C * S;
                    }
                    P = C * D_(I) - S * G;
                    D_(I + 1) = H + S * (C * G + S * D_(I));
                    for (K = 1; K <= N; K++) {
                        H = Z_(K, I + 1);
                        Z_(K, I + 1) = S * Z_(K, I) + C * H;
                        Z_(K, I) = C * Z_(K, I) - S * H;
                    }
                }
                E_(L) = S * P;
                D_(L) = C * P;
            } while (fabs(E_(L)) > B);
        }
        D_(L) = D_(L) + F;
    }
    /* order eigenvalues and eigenvectors (:953-971) */
    for (II = 2; II <= N; II++) {
        I = II - 1;
        K = I;
        P = D_(I);
        for (J = II; J <= N; J++) {
            if (D_(J) >= P) continue;
            K = J;
            P = D_(J);
        }
        if (K == I) continue;
        D_(K) = D_(I);
        D_(I) = P;
        for (J = 1; J <= N; J++) {
            P = Z_(J, I);
            Z_(J, I) = Z_(J, K);
            Z_(J, K) = P;
        }
    }
    return 0;
}

/* TRBAK3 (egrad_brh2.f:980-1059) */
static void brh2_trbak3(int N, const real *A, int M, real *Z)
{
    int I, J, K, L, IZ, IK;
    real H, S;
    if (M == 0 || N == 1) return;
    for (I = 2; I <= N; I++) {
        L = I - 1;
        IZ = (I * L) / 2;
        IK = IZ + I;
        H = A_(IK);
        if (H == 0.0) continue;
        for (J = 1; J <= M; J++) {
            S = 0.0;
            IK = IZ;
            for (K = 1; K <= L; K++) {
                IK = IK + 1;
                S = S + A_(IK) * Z_(K, J);
            }
            S = (S / H) / H;
            IK = IZ;
            for (K = 1; K <= L; K++) {
                IK = IK + 1;
                Z_(K, J) = Z_(K, J) - S * A_(IK);
            }
        }
    }
}

/* RSP (egrad_brh2.f:485-558), N = NM = 4, NV = 10, MATZ = 1 */
static int brh2_rsp(real *A, real *W, real *Z)
{
    real FV1[4], FV2[4];
    int I, J, ierr;
    brh2_tred3(4, A, W, FV1, FV2);
    for (I = 1; I <= 4; I++) {
        for (J = 1; J <= 4; J++) Z_(J, I) = 0.0;
        Z_(I, I) = 1.0;
    }
    ierr = brh2_tql2(4, W, FV1, Z);
    if (ierr != 0) return ierr;
    brh2_trbak3(4, A, 4, Z);
    return 0;
}

#undef A_
#undef D_
#undef E_
#undef E2_

/* POT_brh2 (egrad_brh2.f:188-440).  Rin = (r12, r13, r23) of egrad_brh2; returns IERR of RSP
 * (the reference STOPs on it, :379-385). */
static int brh2_pot(const brh2_consts *c, const real Rin[3], real *ENERGY, real DEDR[3])
{
    real H[10], DH1[10], DH2[10], DH3[10], U[16], E[4];
    real *Z = U;
    real R1, R2, R3, R1S, R2S, R3S, R12, R13, R23, T, T1, T2, T3, T11, T12, T21, T22, T31, T32;
    real CSA, CSA2, SNA2, SNA, SN2A, DCSA21, DSN2A1, DCSA22, DSN2A2, DCSA23, DSN2A3;
    real CSB, CSB2, SNB2, SNB, SN2B, DCSB21, DSN2B1, DCSB22, DSN2B2, DCSB23, DSN2B3;
    real CSG2, DCSG21, DCSG22, DCSG23;
    real RDIF, RDIF2, EX1, EX2, V1HH, V3HH, DV1HH, DV3HH;
    real V1S1, V3S1, V1P1, V3P1, DV1S1, DV3S1, DV1P1, DV3P1;
    real V1S2, V3S2, V1P2, V3P2, DV1S2, DV3S2, DV1P2, DV3P2;
    real S11, S12, S21, S22, S31, S32, P11, P12, P21, P22, P31, P32;
    real DS11, DS12, DS21, DS22, DS31, DS32, DP11, DP12, DP21, DP22, DP31, DP32;
    real D1, D2, D3, SHH, DSHH, SHX1, DSHX1, SHX2, DSHX2;
    int LCOL, K, L, LL, KK, INDEX = 0, ierr, i;
    /* R1 = R(Br-H), R2 = R(H-Br), R3 = R(H-H) (:197-199) */
    R1 = Rin[0];
    R2 = Rin[2];
    R3 = Rin[1];
    if (R1 > R2 + R3) R1 = R2 + R3;
    if (R2 > R1 + R3) R2 = R1 + R3;
    if (R3 > R1 + R2) R3 = R1 + R2;
    R1S = R1 * R1;
    R2S = R2 * R2;
    R3S = R3 * R3;
    R12 = R1 * R2;
    R13 = R1 * R3;
    R23 = R2 * R3;
    T3 = (R1S + R2S - R3S) * 0.5;
    T2 = (R1S + R3S - R2S) * 0.5;
    T1 = (R2S + R3S - R1S) * 0.5;
    CSA = T2 / R13;
    if (fabs(CSA) > BR_ONE) CSA = br_sign(BR_ONE, CSA);
    CSA2 = CSA * CSA;
    SNA2 = 1.0 - CSA2;
    LCOL = SNA2 < BR_EPS;
    SNA = sqrt(SNA2);
    T22 = 2.0 * CSA;
    T11 = 0.0;
    if (!LCOL) T11 = 2.0 * (SNA2 - CSA2) / SNA;
    SN2A = T22 * SNA;
    T = T3 / (R1 * R13);
    DCSA21 = T22 * T;
    DSN2A1 = T11 * T;
    T = -R2 / R13;
    DCSA22 = T22 * T;
    DSN2A2 = T11 * T;
    T = T1 / (R13 * R3);
    DCSA23 = T22 * T;
    DSN2A3 = T11 * T;
    CSB = T1 / R23;
    if (fabs(CSB) > BR_ONE) CSB = br_sign(BR_ONE, CSB);
    CSB2 = CSB * CSB;
    SNB2 = 1.0 - CSB2;
    SNB = sqrt(SNB2);
    T11 = 0.0;
    if (!LCOL) T11 = 2.0 * (SNB2 - CSB2) / SNB;
    T22 = 2.0 * CSB;
    SN2B = T22 * SNB;
    T = -R1 / R23;
    DCSB21 = T22 * T;
    DSN2B1 = T11 * T;
    T = T3 / (R2 * R23);
    DCSB22 = T22 * T;
    DSN2B2 = T11 * T;
    T = T2 / (R23 * R3);
    DCSB23 = T22 * T;
    DSN2B3 = T11 * T;
    T = T3 / R12;
    CSG2 = T * T;
    T = 2.0 * T;
    DCSG21 = T * T2 / (R1 * R12);
    DCSG22 = T * T1 / (R12 * R2);
    DCSG23 = -T * R3 / R12;
    /* diatomic curves: HH (:257-267) */
    RDIF = R3 - BR_RHH;
    EX1 = exp(-BR_AHH * RDIF);
    RDIF2 = RDIF * RDIF;
    EX2 = exp(-BR_BHH * RDIF * RDIF2);
    T1 = BR_DHH * EX1 * EX2;
    V1HH = T1 * (EX1 - 2.0);
    V3HH = BR_ETAHH * T1 * (EX1 + 2.0);
    T1 = 2.0 * BR_AHH * T1;
    T2 = 3.0 * BR_BHH * RDIF2;
    DV1HH = T1 * (1.0 - EX1) - T2 * V1HH;
    DV3HH = -T1 * BR_ETAHH * (1.0 + EX1) - T2 * V3HH;
    /* HX */
    brh2_vhx(R1, &V1S1, &V3S1, &V1P1, &V3P1, &DV1S1, &DV3S1, &DV1P1, &DV3P1);
    brh2_vhx(R2, &V1S2, &V3S2, &V1P2, &V3P2, &DV1S2, &DV3S2, &DV1P2, &DV3P2);
    S11 = V1S1 + 3.0 * V3S1;
    S12 = V1S2 + 3.0 * V3S2;
    S21 = 3.0 * V1S1 + V3S1;
    S22 = 3.0 * V1S2 + V3S2;
    S31 = V1S1 - V3S1;
    S32 = V1S2 - V3S2;
    P11 = V1P1 + 3.0 * V3P1;
    P12 = V1P2 + 3.0 * V3P2;
    P21 = 3.0 * V1P1 + V3P1;
    P22 = 3.0 * V1P2 + V3P2;
    P31 = V1P1 - V3P1;
    P32 = V1P2 - V3P2;
    DS11 = DV1S1 + 3.0 * DV3S1;
    DS12 = DV1S2 + 3.0 * DV3S2;
    DS21 = 3.0 * DV1S1 + DV3S1;
    DS22 = 3.0 * DV1S2 + DV3S2;
    DS31 = DV1S1 - DV3S1;
    DS32 = DV1S2 - DV3S2;
    DP11 = DV1P1 + 3.0 * DV3P1;
    DP12 = DV1P2 + 3.0 * DV3P2;
    DP21 = 3.0 * DV1P1 + DV3P1;
    DP22 = 3.0 * DV1P2 + DV3P2;
    DP31 = DV1P1 - DV3P1;
    DP32 = DV1P2 - DV3P2;
    /* 4x4 Hamiltonian, packed lower triangle (:300-320); index i here is H(i+1) of the source */
    for (i = 0; i < 10; i++) H[i] = DH1[i] = DH2[i] = DH3[i] = 0.0;
    H[0] = V1HH + 0.25 * (S11 * CSA2 + P11 * SNA2 + S12 * CSB2 + P12 * SNB2);
    H[2] = V1HH + 0.25 * (S11 * SNA2 + P11 * CSA2 + S12 * SNB2 + P12 * CSB2);
    H[5] = V3HH + 0.25 * (S21 * CSA2 + P21 * SNA2 + S22 * CSB2 + P22 * SNB2);
    H[9] = V3HH + 0.25 * (S21 * SNA2 + P21 * CSA2 + S22 * SNB2 + P22 * CSB2);
    T11 = S11 - P11;
    T12 = S12 - P12;
    H[1] = 0.0;
    if (!LCOL) H[1] = 0.125 * (T11 * SN2A - T12 * SN2B);
    H[3] = c->RT34 * (S31 * CSA2 + P31 * SNA2 - S32 * CSB2 - P32 * SNB2);
    T31 = S31 - P31;
    T32 = S32 - P32;
    H[4] = 0.0;
    if (!LCOL) H[4] = c->RT38 * (T31 * SN2A + T32 * SN2B);
    H[6] = H[4];
    H[7] = c->RT34 * (S31 * SNA2 + P31 * CSA2 - S32 * SNB2 - P32 * CSB2);
    T21 = S21 - P21;
    T22 = S22 - P22;
    H[8] = 0.0;
    if (!LCOL) H[8] = 0.125 * (T21 * SN2A - T22 * SN2B);
    /* derivative of the Hamiltonian (:323-362) */
    T = T11 * DCSA21 + T12 * DCSB21;
    DH1[0] = 0.25 * (DS11 * CSA2 + DP11 * SNA2 + T);
    DH1[1] = 0.125 * ((DS11 - DP11) * SN2A + T11 * DSN2A1 - T12 * DSN2B1);
    DH1[2] = 0.25 * (DS11 * SNA2 + DP11 * CSA2 - T);
    T = T11 * DCSA22 + T12 * DCSB22;
    DH2[0] = 0.25 * (DS12 * CSB2 + DP12 * SNB2 + T);
    DH2[1] = 0.125 * (-(DS12 - DP12) * SN2B + T11 * DSN2A2 - T12 * DSN2B2);
    DH2[2] = 0.25 * (DS12 * SNB2 + DP12 * CSB2 - T);
    T = 0.25 * (T11 * DCSA23 + T12 * DCSB23);
    DH3[0] = DV1HH + T;
    DH3[1] = 0.125 * (T11 * DSN2A3 - T12 * DSN2B3);
    DH3[2] = DV1HH - T;
    T = T21 * DCSA21 + T22 * DCSB21;
    DH1[5] = 0.25 * (DS21 * CSA2 + DP21 * SNA2 + T);
    DH1[8] = 0.125 * ((DS21 - DP21) * SN2A + T21 * DSN2A1 - T22 * DSN2B1);
    DH1[9] = 0.25 * (DS21 * SNA2 + DP21 * CSA2 - T);
    T = T21 * DCSA22 + T22 * DCSB22;
    DH2[5] = 0.25 * (DS22 * CSB2 + DP22 * SNB2 + T);
    DH2[8] = 0.125 * ((DP22 - DS22) * SN2B + T21 * DSN2A2 - T22 * DSN2B2);
    DH2[9] = 0.25 * (DS22 * SNB2 + DP22 * CSB2 - T);
    T = 0.25 * (T21 * DCSA23 + T22 * DCSB23);
    DH3[5] = DV3HH + T;
    DH3[8] = 0.125 * (T21 * DSN2A3 - T22 * DSN2B3);
    DH3[9] = DV3HH - T;
    T = T31 * DCSA21 - T32 * DCSB21;
    DH1[3] = c->RT34 * (DS31 * CSA2 + DP31 * SNA2 + T);
    DH1[4] = c->RT38 * ((DS31 - DP31) * SN2A + T31 * DSN2A1 + T32 * DSN2B1);
    DH1[6] = DH1[4];
    DH1[7] = c->RT34 * (DS31 * SNA2 + DP31 * CSA2 - T);
    T = T31 * DCSA22 - T32 * DCSB22;
    DH2[3] = c->RT34 * (-DS32 * CSB2 - DP32 * SNB2 + T);
    DH2[4] = c->RT38 * ((DS32 - DP32) * SN2B + T31 * DSN2A2 + T32 * DSN2B2);
    DH2[6] = DH2[4];
    DH2[7] = c->RT34 * (-DS32 * SNB2 - DP32 * CSB2 - T);
    T = T31 * DCSA23 - T32 * DCSB23;
    DH3[3] = c->RT34 * T;
    DH3[4] = c->RT38 * (T31 * DSN2A3 + T32 * DSN2B3);
    DH3[6] = DH3[4];
    DH3[7] = -DH3[3];
    /* diagonalise, lowest root (:367-377) */
    ierr = brh2_rsp(H, E, U);
    if (ierr != 0) return ierr;
    D1 = 0.0;
    D2 = 0.0;
    D3 = 0.0;
    for (L = 1; L <= 4; L++) {
        T1 = Z_(L, 1);
        LL = (L * (L - 1)) / 2;
        for (K = 1; K <= 4; K++) {
            T2 = T1 * Z_(K, 1);
            KK = (K * (K - 1)) / 2;
            if (L >= K) INDEX = LL + K;
            if (K > L) INDEX = KK + L;
            D1 = D1 + T2 * DH1[INDEX - 1];
            D2 = D2 + T2 * DH2[INDEX - 1];
            D3 = D3 + T2 * DH3[INDEX - 1];
        }
    }
    /* supplementary three-centre term (:406-427) */
    T = BR_XIH * R3;
    EX1 = exp(-T);
    SHH = (1.0 + T * (1.0 + c->C3 * T)) * EX1;
    DSHH = -BR_XIH * T * (1.0 + T) * c->C3 * EX1;
    T = c->XIB * R1;
    EX1 = exp(-T);
    SHX1 = R1 * (1.0 + T * (1.0 + c->C3 * T)) * EX1 / c->DENOM;
    DSHX1 = (1.0 + T * (1.0 - c->C3 * T * T)) * EX1 / c->DENOM;
    T = c->XIB * R2;
    EX1 = exp(-T);
    SHX2 = R2 * (1.0 + T * (1.0 + c->C3 * T)) * EX1 / c->DENOM;
    DSHX2 = (1.0 + T * (1.0 - c->C3 * T * T)) * EX1 / c->DENOM;
    RDIF = R1 - R2;
    T1 = c->G * exp(-BR_ALFW * RDIF * RDIF);
    T2 = SHH * (SHX1 + SHX2) + SHX1 * SHX2;
    *ENERGY = E[0] + T2 * T1 * CSG2 + BR_DHH;
    T3 = 2.0 * BR_ALFW * RDIF * CSG2;
    DEDR[0] = D1 + (DSHX1 * (SHH + SHX2) * CSG2 + T2 * (DCSG21 - T3)) * T1;
    DEDR[2] = D2 + (DSHX2 * (SHH + SHX1) * CSG2 + T2 * (DCSG22 + T3)) * T1;
    DEDR[1] = D3 + (DSHH * (SHX1 + SHX2) * CSG2 + T2 * DCSG23) * T1;
    return 0;
}
#undef Z_

void oracle_brh2_pot_real(const real R[3], real *V, real dVdR[3], int *ierr)
{
    const brh2_consts c = brh2_init();
    *ierr = brh2_pot(&c, R, V, dVdR);
}

/* egrad_brh2 (egrad_brh2.f:33-80) */
void oracle_egrad_brh2_real(const real *q, int natoms, int nbeads, real *V, real *dVdq, int *info)
{
    const brh2_consts c = brh2_init();
    int k, d;
    *info = 0;
    for (k = 0; k < nbeads; k++) {
        const real *qk = q + (long)k * 3 * natoms;
        real *gk = dVdq + (long)k * 3 * natoms;
        real AB[3], AC[3], BC[3], R[3], dVdr[3], rAB, rAC, rBC;
        for (d = 0; d < 3; d++) {
            AB[d] = qk[3 * 1 + d] - qk[3 * 0 + d];
            AC[d] = qk[3 * 0 + d] - qk[3 * 2 + d];
            BC[d] = qk[3 * 2 + d] - qk[3 * 1 + d];
        }
        rAB = sqrt(AB[0] * AB[0] + AB[1] * AB[1] + AB[2] * AB[2]);
        rAC = sqrt(AC[0] * AC[0] + AC[1] * AC[1] + AC[2] * AC[2]);
        rBC = sqrt(BC[0] * BC[0] + BC[1] * BC[1] + BC[2] * BC[2]);
        R[0] = rAB;
        R[1] = rAC;
        R[2] = rBC;
        if (brh2_pot(&c, R, &V[k], dVdr) != 0) *info = 1; /* the reference STOPs ('POT 3') */
        for (d = 0; d < 3; d++) {
            gk[3 * 0 + d] = dVdr[1] * AC[d] / rAC - dVdr[0] * AB[d] / rAB;
            gk[3 * 1 + d] = dVdr[0] * AB[d] / rAB - dVdr[2] * BC[d] / rBC;
            gk[3 * 2 + d] = dVdr[2] * BC[d] / rBC - dVdr[1] * AC[d] / rAC;
        }
    }
}
