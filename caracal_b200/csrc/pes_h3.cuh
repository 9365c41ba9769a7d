// pes_h3.cuh -- BKMP2 H + H2 surface as a one-thread-per-image FP64 device function.
//
// Replaces /root/reference/src/egrad_h3.f (egrad_h3 :29-77, pote :79-249 and the routines
// below it).  Written for the GPU, not transcribed: the singlet polynomial is evaluated in
// Horner form, the never-used second derivative of VH2OPT95 (:399-415) is dropped, the two
// structurally identical VBEND blocks (A/G sets, :1010-1143) and CBEND blocks (C/D sets,
// :1166-1346) are one templated routine each, and the shared products are hoisted.
// Compact-geometry corrections (any R < 1.15 a0) stay a rare divergent branch.
//
// Literals: the reference is compiled by gfortran without -fdefault-real-8, so every real
// literal without a D exponent is REAL*4.  FL(x) reproduces that rounding
// at compile time; build with -DCRCL_LITERALS_EXACT for the unrounded decimal values.
#pragma once
#include "crcl_common.cuh"

namespace crcl {
namespace h3 {

struct VbSet {
    double x11, x12, x13, x21, x22, x23, x24, x31, x32, x41, x42, x43, x44, x51, x52, x53;
};
struct CbSet {
    double x11, x12, x13, x14, x15, x21, x22, x23, x24, x31, x32, x41, x42, x43, x44, x51, x52,
        x53, x61, x62, x63, x71, x72, x73, x74, x75, x81, x82, x83, x84;
};

// egrad_h3.f:811-825
CRCL_HD constexpr VbSet VBA_()
{
    return VbSet{
    FL(-.1838073394E+03), FL(0.1334593242E+02), FL(-.2358129537E+00), FL(-.4668193478E+01),
    FL(0.7197506670E+01), FL(0.2162004275E+02), FL(0.2106294028E+02), FL(0.4242962586E+01),
    FL(0.4453505045E+01), FL(-.1456918088E+00), FL(-.1692657366E-01), FL(0.1279520698E+01),
    FL(-.4898940075E+00), FL(0.1742295219E+03), FL(0.3142175348E+02), FL(0.5152903406E+01)};
}
CRCL_HD constexpr VbSet VBG_()
{
    return VbSet{
    FL(-.4765732725E+02), FL(0.3648933563E+01), FL(-.7141145244E-01), FL(0.1002349176E-01),
    FL(0.9989856329E-02), FL(-.4161953634E-02), FL(0.9075807910E-03), FL(-.2693628729E+00),
    FL(-.1399065763E-01), FL(-.1417634346E-01), FL(-.4870024792E-03), FL(0.1312231847E+00),
    FL(-.4409850519E-01), FL(0.5382970863E+02), FL(0.4587102824E+01), FL(0.1768550515E+01)};
}
// egrad_h3.f:842-866
CRCL_HD constexpr CbSet CBC_()
{
    return CbSet{
    FL(0.1860299931E+04), FL(-.6134458037E+03), FL(0.7337207161E+02), FL(-.2676717625E+04),
    FL(0.1344099415E+04), FL(0.1538913137E+03), FL(0.4348007369E+02), FL(0.1719720677E+03),
    FL(0.2115963042E+03), FL(-.7026089414E+02), FL(-.1300938992E+03), FL(0.1310273564E+01),
    FL(-.6175149574E+00), FL(-.2679089358E+02), FL(0.5577477171E+01), FL(-.3543353539E+04),
    FL(-.3740709591E+03), FL(0.7979303144E+02), FL(-.1104230585E+04), FL(0.4603572025E+04),
    FL(-.5593496634E+04), FL(-.1069406434E+02), FL(0.1021807153E+01), FL(0.6669828341E-01),
    FL(0.4168542348E+02), FL(0.1751608567E+02), FL(0.9486883238E+02), FL(-.1519334221E+02),
    FL(0.4024697252E+04), FL(-.2225159395E+02)};
}
CRCL_HD constexpr CbSet CBD_()
{
    return CbSet{
    FL(0.4203543357E+03), FL(-.4922474096E+02), FL(0.3362942544E+00), FL(-.3827423082E+03),
    FL(0.1746726001E+03), FL(0.1699995737E-01), FL(0.1513036778E-01), FL(0.2659119354E-01),
    FL(-.5760387483E-02), FL(0.1020622621E+02), FL(0.1050536271E-01), FL(0.6836172780E+00),
    FL(-.1627858240E+00), FL(-.6925485045E+01), FL(0.1632567385E+01), FL(0.1083595009E+04),
    FL(0.4641431791E+01), FL(-.8233144461E+00), FL(-.6157225942E+02), FL(0.3094361471E+03),
    FL(-.3299631143E+03), FL(0.8866227120E+01), FL(-.1382126854E+01), FL(0.7620770145E-01),
    FL(-.5145757859E+02), FL(0.2046097265E+01), FL(0.2540775558E+01), FL(-.4889246569E+00),
    FL(-.1127439280E+04), FL(-.2269932295E+01)};
}

// Schwenke H2 singlet curve (VH2OPT95, egrad_h3.f:322-397): E and dE/dR.
CRCL_HD __forceinline__ void singlet(double R, double rinv, double& E, double& dE)
{
    constexpr double A0 = FL(0.03537359271649620), A1 = FL(2.013977588700072),
                     A2 = FL(-2.827452449964767), A3 = FL(2.713257715593500),
                     A4 = FL(-2.792039234205731), A5 = FL(2.166542078766724),
                     A6 = FL(-1.272679684173909), A7 = FL(0.5630423099212294),
                     A8 = FL(-0.1879397372273814), A9 = FL(0.04719891893374140),
                     A10 = FL(-0.008851622656489644), A11 = FL(0.001224998776243630),
                     A12 = -1.227820520228028e-04, A13 = 8.638783190083473e-06,
                     A14 = -4.036967926499151e-07, A15 = 1.123286608335365e-08,
                     A16 = -1.406619156782167e-10;
    constexpr double R0 = 3.5284882, DD = 0.160979391, C6 = 6.499027, C8 = 124.3991,
                     C10 = 3285.828;
    constexpr double R02 = R0 * R0, R04 = R02 * R02, R06 = R04 * R02;
    // alpha(R) = A0/R + sum_{n=0}^{15} A_{n+1} R^n  and its derivative, Horner
    double al = A16, dal = 15.0 * A16;
    al = fma(al, R, A15);  dal = fma(dal, R, 14.0 * A15);
    al = fma(al, R, A14);  dal = fma(dal, R, 13.0 * A14);
    al = fma(al, R, A13);  dal = fma(dal, R, 12.0 * A13);
    al = fma(al, R, A12);  dal = fma(dal, R, 11.0 * A12);
    al = fma(al, R, A11);  dal = fma(dal, R, 10.0 * A11);
    al = fma(al, R, A10);  dal = fma(dal, R, 9.0 * A10);
    al = fma(al, R, A9);   dal = fma(dal, R, 8.0 * A9);
    al = fma(al, R, A8);   dal = fma(dal, R, 7.0 * A8);
    al = fma(al, R, A7);   dal = fma(dal, R, 6.0 * A7);
    al = fma(al, R, A6);   dal = fma(dal, R, 5.0 * A6);
    al = fma(al, R, A5);   dal = fma(dal, R, 4.0 * A5);
    al = fma(al, R, A4);   dal = fma(dal, R, 3.0 * A4);
    al = fma(al, R, A3);   dal = fma(dal, R, 2.0 * A3);
    al = fma(al, R, A2);   dal = dal * R + A2;
    al = fma(al, R, A1);
    al = fma(A0, rinv, al);
    dal = fma(-A0 * rinv, rinv, dal);
    const double ex = CRCL_EXP(al);
    const double em1 = ex - 1.0;
    const double R2 = R * R, R4 = R2 * R2, R6 = R4 * R2;
    const double i2 = CRCL_RCP(R2 + R02), i4 = CRCL_RCP(R4 + R04), i6 = CRCL_RCP(R6 + R06);
    const double i25 = i2 * i2 * i2 * i2 * i2;
    E = DD * em1 * em1 - DD - C6 * i6 - C8 * i4 * i4 - C10 * i25;
    dE = 2.0 * DD * em1 * ex * dal +
         R * (6.0 * C6 * R4 * i6 * i6 + 8.0 * C8 * R2 * i4 * i4 * i4 + 10.0 * C10 * i25 * i2);
}

// H2 triplet curve (TRIPLET95, egrad_h3.f:251-320) given the singlet values.
CRCL_HD __forceinline__ void triplet(double R, double rinv, double E1, double dE1, double& E3, double& dE3)
{
    constexpr double RL = 0.95, RR = 1.15;
    constexpr double A1 = FL(-0.0298546962), A2 = FL(-23.9604445036), A3 = FL(-42.5185569474),
                     A4 = FL(2.0382390988), A5 = FL(-11.5214861455), A6 = FL(1.5309487826),
                     C1 = FL(-0.4106358351531854), C2 = FL(-0.0770355790707090),
                     C3 = FL(0.4303193846943223);
    // the three branches of the reference as selects: the outer form costs one exp and one log + exp now, and a
    // warp whose images straddle R = 1.15 a0 no longer executes both sides one after the other
    const double ex = CRCL_EXP(-A4 * R);
    const double ra6 = CRCL_POW(R, -A6);
    const double RSQ = R * R;
    const double Eo = A1 * (A2 + R + A3 * RSQ + A5 * ra6) * ex;
    const double dEo = A1 * ex *
                       (1.0 - A2 * A4 + (2.0 * A3 - A4) * R - A3 * A4 * RSQ - A5 * A6 * (ra6 * rinv) - A4 * A5 * ra6);
    const double DR = R - RL;
    const double cub = (R <= RL) ? 0.0 : C1;
    const double Ei = E1 + cub * DR * DR * DR + C2 * DR + C3;
    const double dEi = dE1 + 3.0 * cub * DR * DR + C2;
    E3 = (R >= RR) ? Eo : Ei;
    dE3 = (R >= RR) ? dEo : dEi;
}

// quantities shared by the VBEND / CBEND blocks
struct Bend {
    double R, RSQ, B2, B3, B3B, EXP1, EXP2, DEXP1, DEXP2;
    double DB2[3], DB3[3], DB3B[3];
};

// one VBEND block (egrad_h3.f:1010-1075 for the A set, :1078-1143 for the G set)
CRCL_HD __forceinline__ void vbend(const VbSet& a, const Bend& c, double B1, const double DB1[3],
                                      double EXP7, double DEXP7, double& V, double dV[3])
{
    const double B12 = B1 * B1, B13 = B12 * B1, B14 = B13 * B1, B15 = B14 * B1;
    const double ASUM = a.x11 + a.x12 * c.R + a.x13 * c.RSQ;
    const double BSUM = a.x21 * B12 + a.x22 * B13 + a.x23 * B14 + a.x24 * B15;
    const double CSUM = a.x31 * B1 * c.EXP1 + a.x32 * B12 * c.EXP2;
    const double DSUM1 = a.x41 * c.EXP1 + a.x42 * c.EXP2;
    const double DSUM2 = a.x43 * c.EXP1 + a.x44 * c.EXP2;
    const double FSUM = a.x51 + a.x52 * c.R + a.x53 * c.RSQ;
    const double DS = c.B3 * DSUM1 + c.B3B * DSUM2;
    V = B1 * ASUM * c.EXP1 + BSUM * c.EXP2 + c.B2 * CSUM + B1 * DS + B1 * FSUM * EXP7;
    const double DASUM = a.x12 + 2.0 * a.x13 * c.R;
    const double DBSUM = 2.0 * a.x21 * B1 + 3.0 * a.x22 * B12 + 4.0 * a.x23 * B13 + 5.0 * a.x24 * B14;
    const double DDSUM1 = a.x41 * c.DEXP1 + a.x42 * c.DEXP2;
    const double DDSUM2 = a.x43 * c.DEXP1 + a.x44 * c.DEXP2;
    const double DFSUM = a.x52 + 2.0 * a.x53 * c.R;
    // coefficient of dB1/dR_i and the part common to all three i (R = R1+R2+R3 only)
    const double cB1 = ASUM * c.EXP1 + DBSUM * c.EXP2 +
                       c.B2 * (a.x31 * c.EXP1 + 2.0 * a.x32 * B1 * c.EXP2) + DS + FSUM * EXP7;
    const double com = B1 * (DASUM * c.EXP1 + ASUM * c.DEXP1) + BSUM * c.DEXP2 +
                       c.B2 * (a.x31 * B1 * c.DEXP1 + a.x32 * B12 * c.DEXP2) +
                       B1 * (c.B3 * DDSUM1 + c.B3B * DDSUM2) + B1 * (DFSUM * EXP7 + FSUM * DEXP7);
#pragma unroll
    for (int i = 0; i < 3; i++)
        dV[i] = cB1 * DB1[i] + com + c.DB2[i] * CSUM + B1 * (c.DB3[i] * DSUM1 + c.DB3B[i] * DSUM2);
}

// one CBEND block without the SUMT factor (egrad_h3.f:1166-1253 / :1256-1345)
CRCL_HD __noinline__ inline void cbend(const CbSet& a, const Bend& c, double B1, const double DB1[3],
                                   double EXP7, double DEXP7, double P, const double DP[3],
                                   double& V, double dV[3])
{
    const double CX1 = a.x51 + a.x83;
    const double RCU = c.RSQ * c.R, PSQ = P * P, PCU = PSQ * P;
    const double B12 = B1 * B1, B13 = B12 * B1, B14 = B13 * B1, B15 = B14 * B1;
    const double ASUM = a.x11 + a.x12 * c.R + a.x13 * c.RSQ + a.x14 / c.R + a.x15 / c.RSQ;
    const double BSUM = a.x21 * B12 + a.x22 * B13 + a.x23 * B14 + a.x24 * B15;
    const double CSUM = a.x31 * B1 * c.EXP1 + a.x32 * B12 * c.EXP2;
    const double DSUM1 = a.x41 * c.EXP1 + a.x42 * c.EXP2;
    const double DSUM2 = a.x43 * c.EXP1 + a.x44 * c.EXP2;
    const double FSUM = CX1 + a.x52 * c.R + a.x53 * c.RSQ;
    const double GSUM = a.x61 + a.x62 / c.R + a.x63 / c.RSQ;
    const double AASUM = a.x71 + a.x72 * P + a.x73 * PSQ + a.x74 / P + a.x75 / PSQ;
    const double FFSUM = a.x81 * P + a.x82 * PSQ + a.x84 / PSQ;
    const double DASUM = a.x12 + 2.0 * a.x13 * c.R - a.x14 / c.RSQ - 2.0 * a.x15 / RCU;
    const double DBSUM = 2.0 * a.x21 * B1 + 3.0 * a.x22 * B12 + 4.0 * a.x23 * B13 + 5.0 * a.x24 * B14;
    const double DDSUM1 = a.x41 * c.DEXP1 + a.x42 * c.DEXP2;
    const double DDSUM2 = a.x43 * c.DEXP1 + a.x44 * c.DEXP2;
    const double DFSUM = a.x52 + 2.0 * a.x53 * c.R;
    const double DGSUM = -a.x62 / c.RSQ - 2.0 * a.x63 / RCU;
    const double DAASUM = a.x72 + 2.0 * a.x73 * P - a.x74 / PSQ - 2.0 * a.x75 / PCU;
    const double DFFSUM = a.x81 + 2.0 * a.x82 * P - 2.0 * a.x84 / PCU;
    const double DS = c.B3 * DSUM1 + c.B3B * DSUM2;
    V = B1 * ASUM * c.EXP1 / P + BSUM * c.EXP2 + c.B2 * CSUM + B1 * DS + B1 * FSUM * EXP7 / P +
        B1 * GSUM * EXP7 + B1 * AASUM * c.EXP2 + B1 * FFSUM * EXP7;
    const double cB1 = ASUM * c.EXP1 / P + DBSUM * c.EXP2 +
                       c.B2 * (a.x31 * c.EXP1 + 2.0 * a.x32 * B1 * c.EXP2) + DS + FSUM * EXP7 / P +
                       GSUM * EXP7 + AASUM * c.EXP2 + FFSUM * EXP7;
    const double com = B1 * (DASUM * c.EXP1 + ASUM * c.DEXP1) / P + BSUM * c.DEXP2 +
                       c.B2 * (a.x31 * B1 * c.DEXP1 + a.x32 * B12 * c.DEXP2) +
                       B1 * (c.B3 * DDSUM1 + c.B3B * DDSUM2) +
                       B1 * (DFSUM * EXP7 + FSUM * DEXP7) / P +
                       B1 * (DGSUM * EXP7 + GSUM * DEXP7) + B1 * AASUM * c.DEXP2 +
                       B1 * FFSUM * DEXP7;
    // coefficient of dP/dR_i
    const double cP = -B1 * ASUM * c.EXP1 / PSQ - B1 * FSUM * EXP7 / PSQ + B1 * DAASUM * c.EXP2 +
                      B1 * DFFSUM * EXP7;
#pragma unroll
    for (int i = 0; i < 3; i++)
        dV[i] = cB1 * DB1[i] + com + cP * DP[i] + c.DB2[i] * CSUM +
                B1 * (c.DB3[i] * DSUM1 + c.DB3B[i] * DSUM2);
}

// |(R1-R2)(R2-R3)(R3-R1)| and its derivatives (ACALC95, egrad_h3.f:547-574)
CRCL_HD __forceinline__ void acalc(const double R[3], double& A, double DA[3])
{
    A = (R[0] - R[1]) * (R[1] - R[2]) * (R[2] - R[0]);
    DA[0] = (-2.0 * R[0] + R[1] + R[2]) * (R[1] - R[2]);
    DA[1] = (-2.0 * R[1] + R[2] + R[0]) * (R[2] - R[0]);
    DA[2] = (-2.0 * R[2] + R[0] + R[1]) * (R[0] - R[1]);
    if (A < 0.0) {
        A = -A;
        DA[0] = -DA[0];
        DA[1] = -DA[1];
        DA[2] = -DA[2];
    }
}

// compact-geometry corrections CSYM95 + CASYM95 + CBEND (egrad_h3.f:611-772, 1145-1360);
// executed only when some R < 1.15 a0.
CRCL_HD __noinline__ inline void compact_terms(const double R[3], const Bend& c, double B1A, double B1B,
                                           const double DB1A[3], const double DB1B[3], double A,
                                           const double DA[3], double& V, double dV[3])
{
    constexpr double RR = 1.15, RP = 1.25, BETA2 = 0.052;
    double T[3], DT[3];
    double SUMT = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        T[i] = 0.0;
        DT[i] = 0.0;
        if (R[i] < RR) {
            const double TOP = RR - R[i], BOT = RP - R[i];
            const double TOP2 = TOP * TOP, TOP3 = TOP2 * TOP;
            T[i] = TOP3 / BOT;
            DT[i] = -3.0 * TOP2 / BOT + TOP3 / (BOT * BOT);
        }
        SUMT += T[i];
    }
    const double SR = c.R, SR2 = c.RSQ, SR3 = SR2 * SR;
    // ---- CSYM95 ----
    {
        constexpr double V1 = FL(-.2071708868E+00), V2 = FL(-.5672350377E+00),
                         V3 = FL(0.9058780367E-02);
        const double EXP3 = exp(-V3 * SR3);
        const double DEXP3 = -3.0 * V3 * SR2 * EXP3;
        double G[3], DG[3], SUMG = 0.0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double RI = R[i];
            const double a = RR - RI, a2 = a * a, a3 = a * a2;
            const double b = RP - RI;
            const double SUMV = V1 + V1 * V2 * RI;
            G[i] = 0.0;
            DG[i] = 0.0;
            if (RI < RR) {
                G[i] = (a3 / b) * SUMV;
                DG[i] = (a3 / (b * b)) * SUMV - 3.0 * (a2 / b) * SUMV + (a3 / b) * V1 * V2;
            }
            SUMG += G[i];
        }
        V += SUMG * EXP3;
#pragma unroll
        for (int i = 0; i < 3; i++) dV[i] += DG[i] * EXP3 + SUMG * DEXP3;
    }
    const double PR = R[0] * R[1] * R[2];
    double DPR[3] = {R[1] * R[2], R[2] * R[0], R[0] * R[1]};
    // ---- CASYM95 ----
    {
        constexpr double U1 = FL(0.2210243144E+00), U2 = FL(0.4367417579E+00),
                         U3 = FL(0.6994985432E-02), U4 = FL(0.1491096501E+01),
                         U5 = FL(0.1602896673E+01), U6 = FL(-.2821747323E+01),
                         U7 = FL(0.4948310833E+00), U8 = FL(-.3540394679E-01),
                         U9 = FL(-.3305809954E+01), U10 = FL(0.3644382172E+01),
                         U11 = FL(-.9997570970E+00), U12 = FL(0.7989919534E-01),
                         U13 = FL(-.1075807322E-02);
        const double A2 = A * A;
        const double PR2 = PR * PR, PR3 = PR2 * PR;
        const double S2 = U9 / PR2 + U10 / PR + U11 + U12 * PR + U13 * PR2;
        const double SERIES = 1.0 + U4 / PR2 + U5 / PR + U6 + U7 * PR + U8 * PR2 + A * S2;
        const double TERM1 = U1 * pow(PR, -U2);
        const double ETERM = exp(-U3 * SR3);
        const double DTERM1P = -U2 * TERM1 / PR;           // d TERM1 / d PR
        const double DETERM = ETERM * (-3.0 * U3 * SR2);   // d ETERM / d R_i
        const double DSERP = (-2.0 * U4 / PR3 - U5 / PR2 + U7 + 2.0 * U8 * PR) +
                             A * (-2.0 * U9 / PR3 - U10 / PR2 + U12 + 2.0 * U13 * PR);
        V += SUMT * A2 * TERM1 * SERIES * ETERM;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double DSER = DPR[i] * DSERP + DA[i] * S2;
            dV[i] += DT[i] * A2 * TERM1 * SERIES * ETERM +
                     2.0 * A * DA[i] * SUMT * TERM1 * SERIES * ETERM +
                     DTERM1P * DPR[i] * SUMT * A2 * SERIES * ETERM +
                     DSER * SUMT * A2 * TERM1 * ETERM + DETERM * SUMT * A2 * TERM1 * SERIES;
        }
    }
    // ---- CBEND (A then B) ----
    {
        const double EXP7 = exp(-BETA2 * SR3);
        const double DEXP7 = -3.0 * BETA2 * SR2 * EXP7;
        double Va, Vb, dVa[3], dVb[3];
        cbend(CBC_(), c, B1A, DB1A, EXP7, DEXP7, PR, DPR, Va, dVa);
        cbend(CBD_(), c, B1B, DB1B, EXP7, DEXP7, PR, DPR, Vb, dVb);
        V += SUMT * (Va + Vb);
#pragma unroll
        for (int i = 0; i < 3; i++) dV[i] += DT[i] * (Va + Vb) + SUMT * (dVa[i] + dVb[i]);
    }
}

// pote (egrad_h3.f:79-249): potential and dV/dR on the three distances
// R = (r12, r13, r23).  warn gets CHGEOM's two conditions as bits (the reference prints).
// SPLIT > 0 (device only): the SPLIT lanes of a trajectory segment evaluate the surface together on the same structure
// (PesSpread, traj_inst.cuh) -- the three pair curves of the London term, 40 % of the instructions of an evaluation, are
// then ONE pass with lane % 3 choosing the distance, the four numbers per distance handed round by shuffles from the
// first three lanes of the segment; Q is summed in the order of the loop (the same bits).
template <int SPLIT = 0>
CRCL_HD __forceinline__ void pote(const double R[3], const double iR[3], double& V, double dV[3], int& warn, int lane = 0,
                                  unsigned mask = 0u)
{
    // CHGEOM (egrad_h3.f:1432-1476)
    {
        const double hi = fmax(R[0], fmax(R[1], R[2]));
        const double lo = fmin(R[0], fmin(R[1], R[2]));
        const double mid = R[0] + R[1] + R[2] - hi - lo;
        warn = 0;
        if (lo + mid + 1.0e-5 < hi) warn |= 1;
        if (lo < 0.2) warn |= 2;
    }
    // ---- London term (H3LOND95, :419-480) ----
    double Q = 0.0, J[3], dQ[3], dJ[3];
#ifdef __CUDA_ARCH__
    if constexpr (SPLIT > 0) {
        const int i = lane % 3;
        const double Ri = (i == 0) ? R[0] : ((i == 1) ? R[1] : R[2]);
        const double iRi = (i == 0) ? iR[0] : ((i == 1) ? iR[1] : iR[2]);
        double E1, dE1, E3, dE3;
        singlet(Ri, iRi, E1, dE1);
        triplet(Ri, iRi, E1, dE1, E3, dE3);
        const double q = 0.5 * (E1 + E3), j = 0.5 * (E1 - E3), dq = 0.5 * (dE1 + dE3), dj = 0.5 * (dE1 - dE3);
        const int first = (int)(threadIdx.x & 31u) - lane;   // lane 0 of this trajectory's segment of the warp
#pragma unroll
        for (int k = 0; k < 3; k++) {
            Q += __shfl_sync(mask, q, first + k);
            J[k] = __shfl_sync(mask, j, first + k);
            dQ[k] = __shfl_sync(mask, dq, first + k);
            dJ[k] = __shfl_sync(mask, dj, first + k);
        }
    } else
#endif
    {
#pragma unroll
        for (int i = 0; i < 3; i++) {
            double E1, dE1, E3, dE3;
            singlet(R[i], iR[i], E1, dE1);
            triplet(R[i], iR[i], E1, dE1, E3, dE3);
            Q += 0.5 * (E1 + E3);
            J[i] = 0.5 * (E1 - E3);
            dQ[i] = 0.5 * (dE1 + dE3);
            dJ[i] = 0.5 * (dE1 - dE3);
        }
    }
    const double d10 = J[1] - J[0], d21 = J[2] - J[1], d20 = J[2] - J[0];
    double rootjt, irootjt;
    sqrt_rsqrt(0.5 * (d10 * d10 + d21 * d21 + d20 * d20) + 1.0e-12, rootjt, irootjt);
    V = Q - rootjt;
    const double hr = 0.5 * irootjt;
    dV[0] = dQ[0] - hr * (2.0 * J[0] - J[1] - J[2]) * dJ[0];
    dV[1] = dQ[1] - hr * (2.0 * J[1] - J[2] - J[0]) * dJ[1];
    dV[2] = dQ[2] - hr * (2.0 * J[2] - J[0] - J[1]) * dJ[2];

    Bend c;
    c.R = R[0] + R[1] + R[2];
    c.RSQ = c.R * c.R;
    // ---- asymmetric correction (VASCAL95, :482-545) ----
    double A, DA[3];
    acalc(R, A, DA);
    {
        constexpr double AA1 = FL(0.3788951192E-02), AA2 = FL(0.1478100901E-02),
                         AA3 = FL(-.1848513849E-03), AA4 = FL(0.9230803609E-05),
                         AA5 = FL(-.1293180255E-06), AA6 = FL(0.5237179303E+00),
                         AA7 = FL(-.1112326215E-02);
        const double A2 = A * A, A3 = A2 * A, A4 = A3 * A, A5 = A4 * A;
        const double EXP1 = CRCL_EXP(-AA1 * c.RSQ * c.R);
        const double EXP6 = CRCL_EXP(-AA6 * c.R);
        const double iSR = CRCL_RCP(c.R);
        const double S = AA2 * A2 + AA3 * A3 + AA4 * A4 + AA5 * A5;
        const double e6r = AA7 * EXP6 * iSR;
        V += S * EXP1 + A2 * e6r;
        const double dSdA = 2.0 * AA2 * A + 3.0 * AA3 * A2 + 4.0 * AA4 * A3 + 5.0 * AA5 * A4;
        const double com = -3.0 * AA1 * c.RSQ * S * EXP1 - A2 * e6r * iSR - AA6 * A2 * e6r;
        const double cA = dSdA * EXP1 + 2.0 * A * e6r;
#pragma unroll
        for (int i = 0; i < 3; i++) dV[i] += com + cA * DA[i];
    }
    // ---- bending terms (VBCB95, :774-1143) ----
    constexpr double Z58 = 0.625, Z38 = 0.375, BETA1 = 0.52, BETA2 = 0.052, BETA3 = 0.79;
    const double R1 = R[0], R2 = R[1], R3 = R[2];
    const double s1 = R1 * R1, s2 = R2 * R2, s3 = R3 * R3;
    const double T1 = s1 - s2 - s3, T2 = s2 - s3 - s1, T3 = s3 - s1 - s2;
    const double i1 = iR[0], i2 = iR[1], i3 = iR[2];   // the reference divides; here one reciprocal per distance serves all quotients
    const double C1 = -0.5 * T1 * (i2 * i3), C2 = -0.5 * T2 * (i3 * i1), C3 = -0.5 * T3 * (i1 * i2);
    const double SUM = C1 + C2 + C3;
    const double B1A = 1.0 - SUM;
    const double SUMB = (4.0 * C1 * C1 * C1 - 3.0 * C1) + (4.0 * C2 * C2 * C2 - 3.0 * C2) +
                        (4.0 * C3 * C3 * C3 - 3.0 * C3);
    const double B1B = 1.0 - (Z58 * SUMB + Z38 * SUM);
    // dC_a/dR_b
    const double DC[3][3] = {
        {-R1 * (i2 * i3), (T1 * (i2 * i2) + 2.0) * (0.5 * i3), (T1 * (i3 * i3) + 2.0) * (0.5 * i2)},
        {(T2 * (i1 * i1) + 2.0) * (0.5 * i3), -R2 * (i1 * i3), (T2 * (i3 * i3) + 2.0) * (0.5 * i1)},
        {(T3 * (i1 * i1) + 2.0) * (0.5 * i2), (T3 * (i2 * i2) + 2.0) * (0.5 * i1), -R3 * (i1 * i2)}};
    const double D1 = 12.0 * C1 * C1 - 3.0, D2 = 12.0 * C2 * C2 - 3.0, D3 = 12.0 * C3 * C3 - 3.0;
    double DB1A[3], DB1B[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double sd = DC[0][i] + DC[1][i] + DC[2][i];
        DB1A[i] = -sd;
        DB1B[i] = -Z58 * (D1 * DC[0][i] + D2 * DC[1][i] + D3 * DC[2][i]) - Z38 * sd;
    }
    c.B2 = i1 + i2 + i3;
    c.B3 = (R2 - R1) * (R2 - R1) + (R3 - R2) * (R3 - R2) + (R1 - R3) * (R1 - R3);
    double iB3B;
    sqrt_rsqrt(c.B3 + 1.0e-12, c.B3B, iB3B);
    c.DB2[0] = -(i1 * i1);
    c.DB2[1] = -(i2 * i2);
    c.DB2[2] = -(i3 * i3);
    c.DB3[0] = 4.0 * R1 - 2.0 * R2 - 2.0 * R3;
    c.DB3[1] = 4.0 * R2 - 2.0 * R3 - 2.0 * R1;
    c.DB3[2] = 4.0 * R3 - 2.0 * R1 - 2.0 * R2;
    const double hb = 0.5 * iB3B;
    c.DB3B[0] = hb * c.DB3[0];
    c.DB3B[1] = hb * c.DB3[1];
    c.DB3B[2] = hb * c.DB3[2];
    c.EXP1 = CRCL_EXP(-BETA1 * c.R);
    c.EXP2 = CRCL_EXP(-BETA2 * c.RSQ);
    const double EXP7 = CRCL_EXP(-BETA3 * c.R);
    c.DEXP1 = -BETA1 * c.EXP1;
    c.DEXP2 = -2.0 * BETA2 * c.R * c.EXP2;
    const double DEXP7 = -BETA3 * EXP7;
    {
        double Va, Vb, dVa[3], dVb[3];
        vbend(VBA_(), c, B1A, DB1A, EXP7, DEXP7, Va, dVa);
        vbend(VBG_(), c, B1B, DB1B, EXP7, DEXP7, Vb, dVb);
        V += Va + Vb;
#pragma unroll
        for (int i = 0; i < 3; i++) dV[i] += dVa[i] + dVb[i];
    }
    // ---- compact-geometry corrections (COMPAC95 :576-609 decides) ----
    // The rare branch works on COPIES: compact_terms is not inlined, so whatever it receives by address lives on the
    // stack -- handed the originals, every evaluation stored c, R, DB1A, DB1B, DA, V and dV there (74 local stores and
    // 20 reloads per step of a one-bead trajectory, profiles/r2ak_chain_h3_source.txt) whether or not the branch ran.
    if (R1 < 1.15 || R2 < 1.15 || R3 < 1.15) {
        const Bend cc = c;
        const double Rc[3] = {R[0], R[1], R[2]}, DAc[3] = {DA[0], DA[1], DA[2]};
        const double DB1Ac[3] = {DB1A[0], DB1A[1], DB1A[2]}, DB1Bc[3] = {DB1B[0], DB1B[1], DB1B[2]};
        double Vc = V, dVc[3] = {dV[0], dV[1], dV[2]};
        compact_terms(Rc, cc, B1A, B1B, DB1Ac, DB1Bc, A, DAc, Vc, dVc);
        V = Vc;
#pragma unroll
        for (int i = 0; i < 3; i++) dV[i] = dVc[i];
    }
}

}  // namespace h3

// PES policy used by the egrad kernel and the fused trajectory kernel.
struct PesH3 {
    static constexpr int NATOMS = 3;
    static constexpr int ID = CRCL_PES_H3;
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
    // q, g: [atom][xyz] of one image.  Returns warning bits (0 = clean).
    CRCL_HD static __forceinline__ int eval(const double* __restrict__ q, double& V,
                                               double* __restrict__ g)
    {
        return eval_split<0>(q, 0, 0u, V, g);
    }
    // the same with the SPLIT lanes of a trajectory segment sharing the London term (pote<SPLIT>; PesSpread)
    static constexpr bool SPLIT_OK = true;
    template <int SPLIT>
    CRCL_HD static __forceinline__ int eval_split(const double* __restrict__ q, int lane, unsigned mask, double& V,
                                                  double* __restrict__ g)
    {
        // R(1)=|q2-q1|, R(2)=|q1-q3|, R(3)=|q3-q2|  (egrad_h3.f:44-63)
        const double ab[3] = {q[3] - q[0], q[4] - q[1], q[5] - q[2]};
        const double ac[3] = {q[0] - q[6], q[1] - q[7], q[2] - q[8]};
        const double bc[3] = {q[6] - q[3], q[7] - q[4], q[8] - q[5]};
        double R[3], iR[3], dV[3];
        sqrt_rsqrt(ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2], R[0], iR[0]);
        sqrt_rsqrt(ac[0] * ac[0] + ac[1] * ac[1] + ac[2] * ac[2], R[1], iR[1]);
        sqrt_rsqrt(bc[0] * bc[0] + bc[1] * bc[1] + bc[2] * bc[2], R[2], iR[2]);
        int warn;
        h3::pote<SPLIT>(R, iR, V, dV, warn, lane, mask);
        const double f0 = dV[0] * iR[0], f1 = dV[1] * iR[1], f2 = dV[2] * iR[2];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            g[d] = f1 * ac[d] - f0 * ab[d];
            g[3 + d] = f0 * ab[d] - f2 * bc[d];
            g[6 + d] = f2 * bc[d] - f1 * ac[d];
        }
        return warn;
    }
};

}  // namespace crcl
