/*
 * pes_h3.c -- CPU oracle: BKMP2 H+H2 potential energy surface.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_real.h).  Parity: UNPINNED by the
 * reference (it ships no golden vectors and cannot be compiled here);
 * pinned instead by literature known-answers and finite differences in tests/.
 *
 * Literal restatement, routine by routine, of /root/reference/src/egrad_h3.f:
 *   egrad_h3   :29-77     oracle_egrad_h3
 *   pote       :79-249    h3_pote
 *   TRIPLET95  :251-320   h3_triplet95
 *   VH2OPT95   :322-417   h3_vh2opt95
 *   H3LOND95   :419-480   h3_lond95
 *   VASCAL95   :482-545   h3_vascal95
 *   ACALC95    :547-574   h3_acalc95
 *   COMPAC95   :576-609   h3_compac95
 *   CASYM95    :611-705   h3_casym95
 *   CSYM95     :707-772   h3_csym95
 *   VBCB95     :774-1430  h3_vbcb95
 *   CHGEOM     :1432-1476 h3_chgeom (warnings -> info bits, never invalidates)
 * Real literals without a D exponent are REAL*4 in the reference as built
 * (Makefile:44, no -fdefault-real-8): those go through F() (oracle_real.h).
 */
#include "oracle_real.h"
#include "oracle.h"

/* ---- VH2OPT95 (egrad_h3.f:322-417): Schwenke H2 singlet curve, E, E', E'' ---- */
static void h3_vh2opt95(real R, real E[3], int ideriv)
{
    const real A0 = F(0.03537359271649620), A1 = F(2.013977588700072),
               A2 = F(-2.827452449964767), A3 = F(2.713257715593500),
               A4 = F(-2.792039234205731), A5 = F(2.166542078766724),
               A6 = F(-1.272679684173909), A7 = F(0.5630423099212294),
               A8 = F(-0.1879397372273814), A9 = F(0.04719891893374140),
               A10 = F(-0.008851622656489644), A11 = F(0.001224998776243630),
               A12 = -1.227820520228028e-04, A13 = 8.638783190083473e-06,
               A14 = -4.036967926499151e-07, A15 = 1.123286608335365e-08,
               A16 = -1.406619156782167e-10;
    const real R0 = 3.5284882, DD = 0.160979391, C6 = 6.499027, C8 = 124.3991,
               C10 = 3285.828;
    real R2, R3, R4, R5, R6, R7, R8, R9, R10, R11, R12, R13, R14, R15;
    real R02, R04, R06, RR2, RR4, RR6, RR25, RR26 = 0.0, RR43 = 0.0;
    real ALPHAR, EXALPH, VSR, VLR, DALPHR = 0.0;

    E[0] = -999.0;
    E[1] = -999.0;
    E[2] = -999.0;
    R2 = R * R;
    R3 = R2 * R;
    R4 = R3 * R;
    R5 = R4 * R;
    R6 = R5 * R;
    R7 = R6 * R;
    R8 = R7 * R;
    R9 = R8 * R;
    R10 = R9 * R;
    R11 = R10 * R;
    R12 = R11 * R;
    R13 = R12 * R;
    R14 = R13 * R;
    R15 = R14 * R;
    R02 = R0 * R0;
    R04 = R02 * R02;
    R06 = R04 * R02;
    RR2 = R2 + R02;
    RR4 = R4 + R04;
    RR6 = R6 + R06;
    RR25 = RR2 * RR2 * RR2 * RR2 * RR2;
    ALPHAR = A0 / R + A1 + A2 * R + A3 * R2 + A4 * R3 + A5 * R4 + A6 * R5 + A7 * R6 +
             A8 * R7 + A9 * R8 + A10 * R9 + A11 * R10 + A12 * R11 + A13 * R12 +
             A14 * R13 + A15 * R14 + A16 * R15;
    EXALPH = exp(ALPHAR);
    VSR = DD * (EXALPH - 1.0) * (EXALPH - 1.0) - DD;
    VLR = -C6 / RR6 - C8 / (RR4 * RR4) - C10 / RR25;
    E[0] = VSR + VLR;

    if (ideriv >= 1) {
        real DVSR, DVLR;
        R3 = R2 * R;
        R5 = R4 * R;
        RR26 = RR25 * RR2;
        RR43 = RR4 * RR4 * RR4;
        DALPHR = -A0 / R2 + A2 + 2.0 * A3 * R + 3.0 * A4 * R2 + 4.0 * A5 * R3 +
                 5.0 * A6 * R4 + 6.0 * A7 * R5 + 7.0 * A8 * R6 + 8.0 * A9 * R7 +
                 9.0 * A10 * R8 + 10.0 * A11 * R9 + 11.0 * A12 * R10 + 12.0 * A13 * R11 +
                 13.0 * A14 * R12 + 14.0 * A15 * R13 + 15.0 * A16 * R14;
        DVSR = 2.0 * DD * (EXALPH - 1.0) * EXALPH * DALPHR;
        DVLR = 6.0 * C6 * R5 / (RR6 * RR6) + 8.0 * C8 * R3 / RR43 + 10.0 * C10 * R / RR26;
        E[1] = DVSR + DVLR;
    }
    if (ideriv >= 2) {
        /* egrad_h3.f:399-415 -- always executed (IDERIV=2 from every caller) and its
           result E(3) is never consumed; kept so that the flop count is the reference's. */
        real RR27, RR44, RR62, RR63, DDALPH, DDVSR, DDVLR;
        R10 = R6 * R4;
        RR27 = RR26 * RR2;
        RR44 = RR43 * RR4;
        RR62 = RR6 * RR6;
        RR63 = RR62 * RR6;
        DDALPH = 2.0 * A0 / R3 + 2.0 * A3 + 6.0 * A4 * R + 12.0 * A5 * R2 + 20.0 * A6 * R3 +
                 30.0 * A7 * R4 + 42.0 * A8 * R5 + 56.0 * A9 * R6 + 72.0 * A10 * R7 +
                 90.0 * A11 * R8 + 110.0 * A12 * R9 + 132.0 * A13 * R10 + 156.0 * A14 * R11 +
                 182.0 * A15 * R12 + 210.0 * A16 * R13;
        DDVSR = 2.0 * DD * EXALPH *
                ((2.0 * EXALPH - 1.0) * DALPHR * DALPHR + (EXALPH - 1.0) * DDALPH);
        DDVLR = -72.0 * C6 * R10 / RR63 - 96.0 * C8 * R6 / RR44 - 120.0 * C10 * R2 / RR27 +
                30.0 * C6 * R4 / RR62 + 24.0 * C8 * R2 / RR43 + 10.0 * C10 / RR26;
        E[2] = DDVSR + DDVLR;
    }
}

/* ---- TRIPLET95 (egrad_h3.f:251-320) ---- */
static void h3_triplet95(real R, real E3[3])
{
    const real RL = 0.95, RR = 1.15;
    const real Z1 = 1.0, Z2 = 2.0;
    const real A1 = F(-0.0298546962), A2 = F(-23.9604445036), A3 = F(-42.5185569474),
               A4 = F(2.0382390988), A5 = F(-11.5214861455), A6 = F(1.5309487826),
               C1 = F(-0.4106358351531854), C2 = F(-0.0770355790707090),
               C3 = F(0.4303193846943223);
    real E1[3];
    E3[2] = 0.0;
    if (R >= RR) {
        real EXDR = exp(-A4 * R);
        real RSQ = R * R;
        real RA6 = pow(R, -A6);
        real RA61;
        E3[0] = A1 * (A2 + R + A3 * RSQ + A5 * RA6) * EXDR;
        RA61 = pow(R, -A6 - Z1);
        E3[1] = A1 * EXDR *
                (Z1 - A2 * A4 + (Z2 * A3 - A4) * R - A3 * A4 * RSQ - A5 * A6 * RA61 -
                 A4 * A5 * RA6);
    }
    if (R < RR) {
        real DR = R - RL;
        h3_vh2opt95(R, E1, 2);
        if (R <= RL) {
            E3[0] = E1[0] + C2 * DR + C3;
            E3[1] = E1[1] + C2;
        } else {
            E3[0] = E1[0] + C1 * DR * DR * DR + C2 * DR + C3;
            E3[1] = E1[1] + 3.0 * C1 * DR * DR + C2;
        }
    }
}

/* ---- H3LOND95 (egrad_h3.f:419-480) ---- */
static void h3_lond95(const real R[3], real *VLON, real DVLON[3])
{
    const real HALF = 0.5, TWO = 2.0, EPS2 = 1.0e-12;
    real Q[3], J[3], E1[3], E3[3], DE1[3], DE3[3], ESING[3], ETRIP[3];
    real SUMQ, SUMJ, JT, ROOTJT;
    int i;
    for (i = 0; i < 3; i++) {
        h3_vh2opt95(R[i], ESING, 2);
        E1[i] = ESING[0];
        DE1[i] = ESING[1];
        h3_triplet95(R[i], ETRIP);
        E3[i] = ETRIP[0];
        DE3[i] = ETRIP[1];
        Q[i] = HALF * (E1[i] + E3[i]);
        J[i] = HALF * (E1[i] - E3[i]);
    }
    SUMQ = Q[0] + Q[1] + Q[2];
    SUMJ = fabs(J[1] - J[0]) * fabs(J[1] - J[0]) + fabs(J[2] - J[1]) * fabs(J[2] - J[1]) +
           fabs(J[2] - J[0]) * fabs(J[2] - J[0]);
    JT = HALF * SUMJ + EPS2;
    ROOTJT = sqrt(JT);
    *VLON = SUMQ - ROOTJT;
    DVLON[0] = HALF * (DE1[0] + DE3[0]) -
               0.25 * (TWO * J[0] - J[1] - J[2]) * (DE1[0] - DE3[0]) / ROOTJT;
    DVLON[1] = HALF * (DE1[1] + DE3[1]) -
               0.25 * (TWO * J[1] - J[2] - J[0]) * (DE1[1] - DE3[1]) / ROOTJT;
    DVLON[2] = HALF * (DE1[2] + DE3[2]) -
               0.25 * (TWO * J[2] - J[0] - J[1]) * (DE1[2] - DE3[2]) / ROOTJT;
}

/* ---- ACALC95 (egrad_h3.f:547-574) ---- */
static void h3_acalc95(real R1, real R2, real R3, real *A, real DA[3])
{
    *A = (R1 - R2) * (R2 - R3) * (R3 - R1);
    DA[0] = (-2.0 * R1 + R2 + R3) * (R2 - R3);
    DA[1] = (-2.0 * R2 + R3 + R1) * (R3 - R1);
    DA[2] = (-2.0 * R3 + R1 + R2) * (R1 - R2);
    if (*A < 0.0) {
        *A = -*A;
        DA[0] = -DA[0];
        DA[1] = -DA[1];
        DA[2] = -DA[2];
    }
}

/* ---- VASCAL95 (egrad_h3.f:482-545) ---- */
static void h3_vascal95(const real RP[3], real *VAS, real DVAS[3])
{
    const real AA1 = F(0.3788951192E-02), AA2 = F(0.1478100901E-02),
               AA3 = F(-.1848513849E-03), AA4 = F(0.9230803609E-05),
               AA5 = F(-.1293180255E-06), AA6 = F(0.5237179303E+00),
               AA7 = F(-.1112326215E-02);
    real R1 = RP[0], R2 = RP[1], R3 = RP[2];
    real R = R1 + R2 + R3;
    real RSQ = R * R;
    real RCU = RSQ * R;
    real A, DA[3], DS[3];
    real A2, A3, A4, A5, EXP1, EXP6, S;
    int i;
    h3_acalc95(R1, R2, R3, &A, DA);
    A2 = A * A;
    A3 = A2 * A;
    A4 = A3 * A;
    A5 = A4 * A;
    EXP1 = exp(-AA1 * RCU);
    EXP6 = exp(-AA6 * R);
    S = AA2 * A2 + AA3 * A3 + AA4 * A4 + AA5 * A5;
    *VAS = S * EXP1 + AA7 * A2 * EXP6 / R;
    for (i = 0; i < 3; i++) {
        DS[i] = (2.0 * AA2 * A + 3.0 * AA3 * A2 + 4.0 * AA4 * A3 + 5.0 * AA5 * A4) * DA[i];
        DVAS[i] = -3.0 * AA1 * RSQ * S * EXP1 + DS[i] * EXP1 - AA7 * A2 * EXP6 / RSQ +
                  2.0 * AA7 * A * DA[i] * EXP6 / R - AA6 * AA7 * A2 * EXP6 / R;
    }
}

/* ---- COMPAC95 (egrad_h3.f:576-609) ---- */
static void h3_compac95(const real R[3], int *ICOMPC, real T[3], real DT[3])
{
    const real RR = 1.15, RP = 1.25;
    int i;
    for (i = 0; i < 3; i++) {
        T[i] = 0.0;
        DT[i] = 0.0;
    }
    *ICOMPC = 0;
    if (R[0] < RR) *ICOMPC += 1;
    if (R[1] < RR) *ICOMPC += 1;
    if (R[2] < RR) *ICOMPC += 1;
    if (*ICOMPC == 0) return;
    for (i = 0; i < 3; i++) {
        if (R[i] < RR) {
            real TOP = RR - R[i];
            real BOT = RP - R[i];
            real TOP2 = TOP * TOP;
            real TOP3 = TOP2 * TOP;
            real BOT2 = BOT * BOT;
            T[i] = TOP3 / BOT;
            DT[i] = -3.0 * TOP2 / BOT + TOP3 / BOT2;
        }
    }
}

/* ---- CASYM95 (egrad_h3.f:611-705) ---- */
static void h3_casym95(const real R[3], real *CAS, real DCAS[3], const real T[3],
                       const real DT[3])
{
    const real U1 = F(0.2210243144E+00), U2 = F(0.4367417579E+00), U3 = F(0.6994985432E-02),
               U4 = F(0.1491096501E+01), U5 = F(0.1602896673E+01), U6 = F(-.2821747323E+01),
               U7 = F(0.4948310833E+00), U8 = F(-.3540394679E-01), U9 = F(-.3305809954E+01),
               U10 = F(0.3644382172E+01), U11 = F(-.9997570970E+00),
               U12 = F(0.7989919534E-01), U13 = F(-.1075807322E-02);
    real A, DA[3], A2, SUMT, SR, PR, SR2, SR3, PR2, PR3, SERIES, TERM1, ETERM;
    real DPR[3], DSUMT[3], DTERM1[3], DETERM[3], DSERIES[3];
    int i;
    *CAS = 0.0;
    DCAS[0] = 0.0;
    DCAS[1] = 0.0;
    DCAS[2] = 0.0;
    h3_acalc95(R[0], R[1], R[2], &A, DA);
    A2 = A * A;
    SUMT = T[0] + T[1] + T[2];
    SR = R[0] + R[1] + R[2];
    PR = R[0] * R[1] * R[2];
    SR2 = SR * SR;
    SR3 = SR2 * SR;
    PR2 = PR * PR;
    PR3 = PR2 * PR;
    SERIES = 1.0 + U4 / PR2 + U5 / PR + U6 + U7 * PR + U8 * PR2 +
             A * (U9 / PR2 + U10 / PR + U11 + U12 * PR + U13 * PR2);
    TERM1 = U1 / pow(PR, U2);
    ETERM = exp(-U3 * SR3);
    *CAS = SUMT * A2 * TERM1 * SERIES * ETERM;
    DPR[0] = R[1] * R[2];
    DPR[1] = R[2] * R[0];
    DPR[2] = R[0] * R[1];
    for (i = 0; i < 3; i++) {
        DSUMT[i] = DT[i];
        DTERM1[i] = -1.0 * U1 * U2 * pow(PR, -U2 - 1.0) * DPR[i];
        DETERM[i] = ETERM * (-3.0 * U3 * SR2);
        DSERIES[i] = DPR[i] * (-2.0 * U4 / PR3 - U5 / PR2 + U7 + 2.0 * U8 * PR) +
                     DA[i] * (U9 / PR2 + U10 / PR + U11 + U12 * PR + U13 * PR2) +
                     A * DPR[i] * (-2.0 * U9 / PR3 - U10 / PR2 + U12 + 2.0 * U13 * PR);
        DCAS[i] = DSUMT[i] * A2 * TERM1 * SERIES * ETERM +
                  2.0 * A * DA[i] * SUMT * TERM1 * SERIES * ETERM +
                  DTERM1[i] * SUMT * A2 * SERIES * ETERM +
                  DSERIES[i] * SUMT * A2 * TERM1 * ETERM +
                  DETERM[i] * SUMT * A2 * TERM1 * SERIES;
    }
}

/* ---- CSYM95 (egrad_h3.f:707-772) ---- */
static void h3_csym95(const real R[3], real *CAL, real DCAL[3])
{
    const real RR = 1.15, RP = 1.25;
    const real V1 = F(-.2071708868E+00), V2 = F(-.5672350377E+00), V3 = F(0.9058780367E-02);
    real G[3], DG[3], SUMV[3];
    real SR = R[0] + R[1] + R[2];
    real SR2 = SR * SR;
    real SR3 = SR * SR2;
    real EXP3 = exp(-V3 * SR3);
    real DEXP3 = -3.0 * V3 * SR2 * EXP3;
    real SUMG;
    int i;
    for (i = 0; i < 3; i++) {
        real RI = R[i];
        real RRRI = RR - RI;
        real RRRI2 = RRRI * RRRI;
        real RRRI3 = RRRI * RRRI2;
        real RPRI = RP - RI;
        real RPRI2 = RPRI * RPRI;
        G[i] = 0.0;
        DG[i] = 0.0;
        SUMV[i] = V1 + V1 * V2 * RI;
        if (RI < RR) {
            G[i] = (RRRI3 / RPRI) * SUMV[i];
            DG[i] = (RRRI3 / RPRI2) * SUMV[i] - 3.0 * (RRRI2 / RPRI) * SUMV[i] +
                    (RRRI3 / RPRI) * V1 * V2;
        }
    }
    SUMG = G[0] + G[1] + G[2];
    *CAL = SUMG * EXP3;
    DCAL[0] = DG[0] * EXP3 + SUMG * DEXP3;
    DCAL[1] = DG[1] * EXP3 + SUMG * DEXP3;
    DCAL[2] = DG[2] * EXP3 + SUMG * DEXP3;
}

/* coefficient sets of VBCB95 (egrad_h3.f:809-866); index [0]=A/C set, [1]=G/D set */
typedef struct {
    real x11, x12, x13, x21, x22, x23, x24, x31, x32, x41, x42, x43, x44, x51, x52, x53;
} h3_vb_set;
typedef struct {
    real x11, x12, x13, x14, x15, x21, x22, x23, x24, x31, x32, x41, x42, x43, x44, x51, x52,
        x53, x61, x62, x63, x71, x72, x73, x74, x75, x81, x82, x83, x84;
} h3_cb_set;

/* ---- VBCB95 (egrad_h3.f:774-1430) ----
 * The reference writes the A/G (VBEND) and C/D (CBEND) blocks out twice with identical
 * structure; the two passes are restated here as one loop body executed for set 0 (B1A)
 * and set 1 (B1B) in the reference's order.  Expression order inside each statement is
 * the reference's.
 */
static void h3_vbcb95(const real RPASS[3], int ICOMPC, const real T[3], const real DT[3],
                      real *VBNDA, real *VBNDB, real DVBNDA[3], real DVBNDB[3], real *CBNDA,
                      real *CBNDB, real DCBNDA[3], real DCBNDB[3])
{
    const real Z58 = 0.625, Z38 = 0.375;
    const real BETA1 = 0.52, BETA2 = 0.052, BETA3 = 0.79;
    const h3_vb_set VBS[2] = {
        {F(-.1838073394E+03), F(0.1334593242E+02), F(-.2358129537E+00), F(-.4668193478E+01),
         F(0.7197506670E+01), F(0.2162004275E+02), F(0.2106294028E+02), F(0.4242962586E+01),
         F(0.4453505045E+01), F(-.1456918088E+00), F(-.1692657366E-01), F(0.1279520698E+01),
         F(-.4898940075E+00), F(0.1742295219E+03), F(0.3142175348E+02), F(0.5152903406E+01)},
        {F(-.4765732725E+02), F(0.3648933563E+01), F(-.7141145244E-01), F(0.1002349176E-01),
         F(0.9989856329E-02), F(-.4161953634E-02), F(0.9075807910E-03), F(-.2693628729E+00),
         F(-.1399065763E-01), F(-.1417634346E-01), F(-.4870024792E-03), F(0.1312231847E+00),
         F(-.4409850519E-01), F(0.5382970863E+02), F(0.4587102824E+01), F(0.1768550515E+01)}};
    const h3_cb_set CBS[2] = {
        {F(0.1860299931E+04), F(-.6134458037E+03), F(0.7337207161E+02), F(-.2676717625E+04),
         F(0.1344099415E+04), F(0.1538913137E+03), F(0.4348007369E+02), F(0.1719720677E+03),
         F(0.2115963042E+03), F(-.7026089414E+02), F(-.1300938992E+03), F(0.1310273564E+01),
         F(-.6175149574E+00), F(-.2679089358E+02), F(0.5577477171E+01), F(-.3543353539E+04),
         F(-.3740709591E+03), F(0.7979303144E+02), F(-.1104230585E+04), F(0.4603572025E+04),
         F(-.5593496634E+04), F(-.1069406434E+02), F(0.1021807153E+01), F(0.6669828341E-01),
         F(0.4168542348E+02), F(0.1751608567E+02), F(0.9486883238E+02), F(-.1519334221E+02),
         F(0.4024697252E+04), F(-.2225159395E+02)},
        {F(0.4203543357E+03), F(-.4922474096E+02), F(0.3362942544E+00), F(-.3827423082E+03),
         F(0.1746726001E+03), F(0.1699995737E-01), F(0.1513036778E-01), F(0.2659119354E-01),
         F(-.5760387483E-02), F(0.1020622621E+02), F(0.1050536271E-01), F(0.6836172780E+00),
         F(-.1627858240E+00), F(-.6925485045E+01), F(0.1632567385E+01), F(0.1083595009E+04),
         F(0.4641431791E+01), F(-.8233144461E+00), F(-.6157225942E+02), F(0.3094361471E+03),
         F(-.3299631143E+03), F(0.8866227120E+01), F(-.1382126854E+01), F(0.7620770145E-01),
         F(-.5145757859E+02), F(0.2046097265E+01), F(0.2540775558E+01), F(-.4889246569E+00),
         F(-.1127439280E+04), F(-.2269932295E+01)}};
    real R1 = RPASS[0], R2 = RPASS[1], R3 = RPASS[2];
    real T1, T2, T3, C1, C2, C3, SUM, B1A, B1B, C1CUBE, C2CUBE, C3CUBE;
    real COS3T1, COS3T2, COS3T3, SUMB;
    real DC1DR1, DC2DR2, DC3DR3, DC1DR2, DC1DR3, DC2DR1, DC2DR3, DC3DR1, DC3DR2;
    real DB1[2][3], B1S[2], D1, D2, D3;
    real R, RSQ, B2, B3, EPS2, B3B, DB2[3], DB3[6];
    real EXP1, EXP2, EXP7, DEXP1, DEXP2, DEXP7;
    real VBND[2], DVBND[2][3];
    int s, i;

    T1 = R1 * R1 - R2 * R2 - R3 * R3;
    T2 = R2 * R2 - R3 * R3 - R1 * R1;
    T3 = R3 * R3 - R1 * R1 - R2 * R2;
    C1 = T1 / (-2.0 * R2 * R3);
    C2 = T2 / (-2.0 * R3 * R1);
    C3 = T3 / (-2.0 * R1 * R2);
    SUM = C1 + C2 + C3;
    B1A = 1.0 - SUM;
    C1CUBE = C1 * C1 * C1;
    C2CUBE = C2 * C2 * C2;
    C3CUBE = C3 * C3 * C3;
    COS3T1 = 4.0 * C1CUBE - 3.0 * C1;
    COS3T2 = 4.0 * C2CUBE - 3.0 * C2;
    COS3T3 = 4.0 * C3CUBE - 3.0 * C3;
    SUMB = COS3T1 + COS3T2 + COS3T3;
    B1B = 1.0 - (Z58 * SUMB + Z38 * SUM);
    DC1DR1 = -R1 / (R2 * R3);
    DC2DR2 = -R2 / (R1 * R3);
    DC3DR3 = -R3 / (R1 * R2);
    DC1DR2 = (T1 / (R2 * R2) + 2.0) / (2.0 * R3);
    DC1DR3 = (T1 / (R3 * R3) + 2.0) / (2.0 * R2);
    DC2DR1 = (T2 / (R1 * R1) + 2.0) / (2.0 * R3);
    DC2DR3 = (T2 / (R3 * R3) + 2.0) / (2.0 * R1);
    DC3DR1 = (T3 / (R1 * R1) + 2.0) / (2.0 * R2);
    DC3DR2 = (T3 / (R2 * R2) + 2.0) / (2.0 * R1);
    DB1[0][0] = -1.0 * (DC1DR1 + DC2DR1 + DC3DR1);
    DB1[0][1] = -1.0 * (DC1DR2 + DC2DR2 + DC3DR2);
    DB1[0][2] = -1.0 * (DC1DR3 + DC2DR3 + DC3DR3);
    D1 = 12.0 * C1 * C1 - 3.0;
    D2 = 12.0 * C2 * C2 - 3.0;
    D3 = 12.0 * C3 * C3 - 3.0;
    DB1[1][0] = -Z58 * (D1 * DC1DR1 + D2 * DC2DR1 + D3 * DC3DR1) -
                Z38 * (DC1DR1 + DC2DR1 + DC3DR1);
    DB1[1][1] = -Z58 * (D1 * DC1DR2 + D2 * DC2DR2 + D3 * DC3DR2) -
                Z38 * (DC1DR2 + DC2DR2 + DC3DR2);
    DB1[1][2] = -Z58 * (D1 * DC1DR3 + D2 * DC2DR3 + D3 * DC3DR3) -
                Z38 * (DC1DR3 + DC2DR3 + DC3DR3);
    B1S[0] = B1A;
    B1S[1] = B1B;

    R = R1 + R2 + R3;
    RSQ = R * R;
    B2 = 1.0 / R1 + 1.0 / R2 + 1.0 / R3;
    B3 = (R2 - R1) * (R2 - R1) + (R3 - R2) * (R3 - R2) + (R1 - R3) * (R1 - R3);
    EPS2 = 1.0e-12;
    B3B = sqrt(B3 + EPS2);
    DB2[0] = -1.0 / (R1 * R1);
    DB2[1] = -1.0 / (R2 * R2);
    DB2[2] = -1.0 / (R3 * R3);
    DB3[0] = 4.0 * R1 - 2.0 * R2 - 2.0 * R3;
    DB3[1] = 4.0 * R2 - 2.0 * R3 - 2.0 * R1;
    DB3[2] = 4.0 * R3 - 2.0 * R1 - 2.0 * R2;
    DB3[3] = 0.5 * DB3[0] / B3B;
    DB3[4] = 0.5 * DB3[1] / B3B;
    DB3[5] = 0.5 * DB3[2] / B3B;
    EXP1 = exp(-BETA1 * R);
    EXP2 = exp(-BETA2 * RSQ);
    EXP7 = exp(-BETA3 * R);
    DEXP1 = -BETA1 * EXP1;
    DEXP2 = -2.0 * BETA2 * R * EXP2;
    DEXP7 = -BETA3 * EXP7;

    /* VBNDA then VBNDB (egrad_h3.f:1010-1143) */
    for (s = 0; s < 2; s++) {
        const h3_vb_set *a = &VBS[s];
        const real *DB1X = DB1[s];
        real B1 = B1S[s];
        real B12 = B1 * B1;
        real B13 = B12 * B1;
        real B14 = B13 * B1;
        real B15 = B14 * B1;
        real ASUM = a->x11 + a->x12 * R + a->x13 * RSQ;
        real BSUM = a->x21 * B12 + a->x22 * B13 + a->x23 * B14 + a->x24 * B15;
        real CSUM = a->x31 * B1 * EXP1 + a->x32 * B12 * EXP2;
        real DSUM1 = a->x41 * EXP1 + a->x42 * EXP2;
        real DSUM2 = a->x43 * EXP1 + a->x44 * EXP2;
        real FSUM = a->x51 + a->x52 * R + a->x53 * RSQ;
        real VB1 = B1 * ASUM * EXP1;
        real VB2 = BSUM * EXP2;
        real VB3 = B2 * CSUM;
        real VB4 = B1 * B3 * DSUM1 + B1 * B3B * DSUM2;
        real VB5 = B1 * FSUM * EXP7;
        real DASUM = a->x12 + 2.0 * a->x13 * R;
        real DBSUM = 2.0 * a->x21 * B1 + 3.0 * a->x22 * B12 + 4.0 * a->x23 * B13 +
                     5.0 * a->x24 * B14;
        real DDSUM1 = a->x41 * DEXP1 + a->x42 * DEXP2;
        real DDSUM2 = a->x43 * DEXP1 + a->x44 * DEXP2;
        real DFSUM = a->x52 + 2.0 * a->x53 * R;
        VBND[s] = VB1 + VB2 + VB3 + VB4 + VB5;
        for (i = 0; i < 3; i++) {
            real DVB1 = DB1X[i] * ASUM * EXP1 + B1 * DASUM * EXP1 + B1 * ASUM * DEXP1;
            real DVB2 = DBSUM * DB1X[i] * EXP2 + BSUM * DEXP2;
            real DVB3 = DB2[i] * CSUM +
                        B2 * (a->x31 * DB1X[i] * EXP1 + a->x31 * B1 * DEXP1 +
                              2.0 * a->x32 * B1 * DB1X[i] * EXP2 + a->x32 * B12 * DEXP2);
            real DVB4 = DB1X[i] * B3 * DSUM1 + B1 * DB3[i] * DSUM1 + B1 * B3 * DDSUM1 +
                        DB1X[i] * B3B * DSUM2 + B1 * DB3[3 + i] * DSUM2 + B1 * B3B * DDSUM2;
            real DVB5 = DB1X[i] * FSUM * EXP7 + B1 * DFSUM * EXP7 + B1 * FSUM * DEXP7;
            DVBND[s][i] = DVB1 + DVB2 + DVB3 + DVB4 + DVB5;
        }
    }
    *VBNDA = VBND[0];
    *VBNDB = VBND[1];
    for (i = 0; i < 3; i++) {
        DVBNDA[i] = DVBND[0][i];
        DVBNDB[i] = DVBND[1][i];
    }

    /* CBEND correction terms only for compact geometries (egrad_h3.f:1145) */
    if (ICOMPC == 0) return;
    {
        real SUMT = T[0] + T[1] + T[2];
        real RCU = RSQ * R;
        real P = R1 * R2 * R3;
        real PSQ = P * P;
        real PCU = PSQ * P;
        real DP[3], CBNDS[2], DCBNDS[2][3];
        DP[0] = R2 * R3;
        DP[1] = R3 * R1;
        DP[2] = R1 * R2;
        CBNDS[0] = 0.0;
        CBNDS[1] = 0.0;
        /* EXP7 is redefined with BETA2*R^3 (egrad_h3.f:1163-1164) */
        EXP7 = exp(-BETA2 * RCU);
        DEXP7 = -3.0 * BETA2 * RSQ * EXP7;
        for (s = 0; s < 2; s++) {
            const h3_cb_set *c = &CBS[s];
            const real *DB1X = DB1[s];
            real CX1 = c->x51 + c->x83; /* :910-911 */
            real B1 = B1S[s];
            real B12 = B1 * B1;
            real B13 = B12 * B1;
            real B14 = B13 * B1;
            real B15 = B14 * B1;
            real ASUM = c->x11 + c->x12 * R + c->x13 * RSQ + c->x14 / R + c->x15 / RSQ;
            real BSUM = c->x21 * B12 + c->x22 * B13 + c->x23 * B14 + c->x24 * B15;
            real CSUM = c->x31 * B1 * EXP1 + c->x32 * B12 * EXP2;
            real DSUM1 = c->x41 * EXP1 + c->x42 * EXP2;
            real DSUM2 = c->x43 * EXP1 + c->x44 * EXP2;
            real FSUM = CX1 + c->x52 * R + c->x53 * RSQ;
            real GSUM = c->x61 + c->x62 / R + c->x63 / RSQ;
            real AASUM = c->x71 + c->x72 * P + c->x73 * PSQ + c->x74 / P + c->x75 / PSQ;
            real FFSUM = c->x81 * P + c->x82 * PSQ + c->x84 / PSQ;
            real DASUM = c->x12 + 2.0 * c->x13 * R - c->x14 / RSQ - 2.0 * c->x15 / RCU;
            real DBSUM = 2.0 * c->x21 * B1 + 3.0 * c->x22 * B12 + 4.0 * c->x23 * B13 +
                         5.0 * c->x24 * B14;
            real DDSUM1 = c->x41 * DEXP1 + c->x42 * DEXP2;
            real DDSUM2 = c->x43 * DEXP1 + c->x44 * DEXP2;
            real DFSUM = c->x52 + 2.0 * c->x53 * R;
            real DGSUM = -c->x62 / RSQ - 2.0 * c->x63 / RCU;
            real DAASUM = c->x72 + 2.0 * c->x73 * P - c->x74 / PSQ - 2.0 * c->x75 / PCU;
            real DFFSUM = c->x81 + 2.0 * c->x82 * P - 2.0 * c->x84 / PCU;
            real CB1 = B1 * ASUM * EXP1 / P;
            real CB2 = BSUM * EXP2;
            real CB3 = B2 * CSUM;
            real CB4 = B1 * B3 * DSUM1 + B1 * B3B * DSUM2;
            real CB5 = B1 * FSUM * EXP7 / P;
            real CB6 = B1 * GSUM * EXP7;
            real CB7 = B1 * AASUM * EXP2;
            real CB8 = B1 * FFSUM * EXP7;
            CBNDS[s] = CB1 + CB2 + CB3 + CB4 + CB5 + CB6 + CB7 + CB8;
            for (i = 0; i < 3; i++) {
                real DCB1 = DB1X[i] * ASUM * EXP1 / P + B1 * DASUM * EXP1 / P +
                            B1 * ASUM * DEXP1 / P - B1 * ASUM * EXP1 * DP[i] / PSQ;
                real DCB2 = DBSUM * DB1X[i] * EXP2 + BSUM * DEXP2;
                real DCB3 = DB2[i] * CSUM +
                            B2 * (c->x31 * DB1X[i] * EXP1 + c->x31 * B1 * DEXP1 +
                                  c->x32 * 2.0 * B1 * DB1X[i] * EXP2 + c->x32 * B12 * DEXP2);
                real DCB4 = DB1X[i] * B3 * DSUM1 + B1 * DB3[i] * DSUM1 + B1 * B3 * DDSUM1 +
                            DB1X[i] * B3B * DSUM2 + B1 * DB3[3 + i] * DSUM2 +
                            B1 * B3B * DDSUM2;
                real DCB5 = DB1X[i] * FSUM * EXP7 / P + B1 * DFSUM * EXP7 / P +
                            B1 * FSUM * DEXP7 / P - B1 * FSUM * EXP7 * DP[i] / PSQ;
                real DCB6 = DB1X[i] * GSUM * EXP7 + B1 * DGSUM * EXP7 + B1 * GSUM * DEXP7;
                real DCB7 = DB1X[i] * AASUM * EXP2 + B1 * DAASUM * DP[i] * EXP2 +
                            B1 * AASUM * DEXP2;
                real DCB8 = DB1X[i] * FFSUM * EXP7 + B1 * DFFSUM * DP[i] * EXP7 +
                            B1 * FFSUM * DEXP7;
                DCBNDS[s][i] = DCB1 + DCB2 + DCB3 + DCB4 + DCB5 + DCB6 + DCB7 + DCB8;
            }
        }
        for (i = 0; i < 3; i++) {
            DCBNDA[i] = DT[i] * CBNDS[0] + SUMT * DCBNDS[0][i];
            DCBNDB[i] = DT[i] * CBNDS[1] + SUMT * DCBNDS[1][i];
        }
        *CBNDA = SUMT * CBNDS[0];
        *CBNDB = SUMT * CBNDS[1];
    }
}

/* ---- CHGEOM (egrad_h3.f:1432-1476): only prints warnings in the reference; the
 * oracle returns them as bits (1: triangle inequality, 2: R<0.2) and never invalidates. */
static int h3_chgeom(const real R[3])
{
    const real DERR = 1.0e-5;
    real RHI, RMID, RLO, t;
    int warn = 0;
    RHI = (R[0] > R[1]) ? R[0] : R[1];
    if (R[2] >= RHI) {
        RMID = RHI;
        RHI = R[2];
    } else {
        t = (R[0] < R[1]) ? R[0] : R[1];
        RMID = (R[2] > t) ? R[2] : t;
    }
    RLO = (R[0] < R[1]) ? R[0] : R[1];
    RLO = (RLO < R[2]) ? RLO : R[2];
    if (RLO + RMID + DERR < RHI) warn |= 1;
    if (RLO < 0.2) warn |= 2;
    return warn;
}

/* ---- pote (egrad_h3.f:79-249) ---- */
static int h3_pote(const real R[3], real *pe, real dpe[3])
{
    real VLON = 0.0, VAS = 0.0, VBNDA = 0.0, VBNDB = 0.0, CAL = 0.0, CAS = 0.0, CBNDA = 0.0,
         CBNDB = 0.0;
    real DVLON[3], DVAS[3], DVBNDA[3], DVBNDB[3], DCAL[3], DCAS[3], DCBNDA[3], DCBNDB[3];
    real T[3], DT[3];
    int ICOMPC, i, warn;
    warn = h3_chgeom(R);
    for (i = 0; i < 3; i++) {
        DVLON[i] = 0.0;
        DVAS[i] = 0.0;
        DVBNDA[i] = 0.0;
        DVBNDB[i] = 0.0;
        DCAL[i] = 0.0;
        DCAS[i] = 0.0;
        DCBNDA[i] = 0.0;
        DCBNDB[i] = 0.0;
    }
    h3_lond95(R, &VLON, DVLON);
    h3_vascal95(R, &VAS, DVAS);
    h3_compac95(R, &ICOMPC, T, DT);
    if (ICOMPC >= 1) {
        h3_csym95(R, &CAL, DCAL);
        h3_casym95(R, &CAS, DCAS, T, DT);
    }
    h3_vbcb95(R, ICOMPC, T, DT, &VBNDA, &VBNDB, DVBNDA, DVBNDB, &CBNDA, &CBNDB, DCBNDA, DCBNDB);
    *pe = VLON + VAS + VBNDA + VBNDB + CAL + CAS + CBNDA + CBNDB;
    for (i = 0; i < 3; i++)
        dpe[i] = DVLON[i] + DVAS[i] + DVBNDA[i] + DVBNDB[i] + DCAL[i] + DCAS[i] + DCBNDA[i] +
                 DCBNDB[i];
    return warn;
}

/* ---- egrad_h3 (egrad_h3.f:29-77), q(3,Natoms,Nbeads) column-major = C [bead][atom][xyz] ---- */
void oracle_egrad_h3_real(const real *q, int natoms, int nbeads, real *V, real *dVdq, int *info)
{
    int k;
    *info = 0;
    for (k = 0; k < nbeads; k++) {
        const real *qk = q + (long)k * 3 * natoms;
        real *gk = dVdq + (long)k * 3 * natoms;
        real R[3], dVdr[3];
        real xAB, yAB, zAB, rAB, xAC, yAC, zAC, rAC, xBC, yBC, zBC, rBC;
        xAB = qk[3 * 1 + 0] - qk[3 * 0 + 0];
        yAB = qk[3 * 1 + 1] - qk[3 * 0 + 1];
        zAB = qk[3 * 1 + 2] - qk[3 * 0 + 2];
        rAB = sqrt(xAB * xAB + yAB * yAB + zAB * zAB);
        R[0] = rAB;
        xAC = qk[3 * 0 + 0] - qk[3 * 2 + 0];
        yAC = qk[3 * 0 + 1] - qk[3 * 2 + 1];
        zAC = qk[3 * 0 + 2] - qk[3 * 2 + 2];
        rAC = sqrt(xAC * xAC + yAC * yAC + zAC * zAC);
        R[1] = rAC;
        xBC = qk[3 * 2 + 0] - qk[3 * 1 + 0];
        yBC = qk[3 * 2 + 1] - qk[3 * 1 + 1];
        zBC = qk[3 * 2 + 2] - qk[3 * 1 + 2];
        rBC = sqrt(xBC * xBC + yBC * yBC + zBC * zBC);
        R[2] = rBC;
        *info |= h3_pote(R, &V[k], dVdr);
        gk[3 * 0 + 0] = dVdr[1] * xAC / rAC - dVdr[0] * xAB / rAB;
        gk[3 * 0 + 1] = dVdr[1] * yAC / rAC - dVdr[0] * yAB / rAB;
        gk[3 * 0 + 2] = dVdr[1] * zAC / rAC - dVdr[0] * zAB / rAB;
        gk[3 * 1 + 0] = dVdr[0] * xAB / rAB - dVdr[2] * xBC / rBC;
        gk[3 * 1 + 1] = dVdr[0] * yAB / rAB - dVdr[2] * yBC / rBC;
        gk[3 * 1 + 2] = dVdr[0] * zAB / rAB - dVdr[2] * zBC / rBC;
        gk[3 * 2 + 0] = dVdr[2] * xBC / rBC - dVdr[1] * xAC / rAC;
        gk[3 * 2 + 1] = dVdr[2] * yBC / rBC - dVdr[1] * yAC / rAC;
        gk[3 * 2 + 2] = dVdr[2] * zBC / rBC - dVdr[1] * zAC / rAC;
    }
}

/* pote on internal distances, exported for the known-answer tests */
void oracle_h3_pote_real(const real R[3], real *pe, real dpe[3]) { (void)h3_pote(R, pe, dpe); }
