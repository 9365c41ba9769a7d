// pes_nh3x.cuh -- the three-hydrogen members of the POTLIB family of pes_ch4h.cuh, FP64:
//   NH3 + Cl -> NH2 + HCl   (Monge-Palacios, Rangel, Corchado, Espinosa-Garcia, Int. J. Quantum Chem. 112, 1887 (2012))
//   NH3 + OH -> NH2 + H2O   (Monge-Palacios, Rangel, Espinosa-Garcia, J. Chem. Phys. 138, 084305 (2013))
//
// Replaces /root/reference/src/egrad_clnh3.f (egrad_clnh3 :36-67, POT_clnh3 :82-210, coorden :216-306, refangles
// :312-424, stretch :580-783, ipbend :941-1058, ipforce :1416-1546, switchf :1552-1688, constants :1780-1898) and
// /root/reference/src/egrad_nh3oh.f (egrad_nh3oh :73-131, POT_nh3oh :160-311 and the routines below it, constants
// :1993-2131).  Both sources switch the out-of-plane term off (energy = stretch + in-plane bend).
//
// The evaluation is written in internal coordinates: the energy is a function of the seven distances r(N-b),
// r(N-H_i), r(b-H_i) and the three H-N-H angles (plus, for NH3 + OH, the spectator O-H length and three H-O-H
// angles); its partial derivatives are accumulated per distance / angle and mapped to the atoms once through the
// bond unit vectors -- the source differentiates every term straight to Cartesian components.
//
// Two properties of the sources that are reproduced because they are part of the reference's answer:
//   * NH3 + Cl: the equilibrium N-H length r0ch is a tanh blend of its reactant and product values on the three
//     N-H distances (coorden :287-299), but the analytic gradient treats it as a constant, so the gradient returned
//     is not exactly the derivative of the energy returned.  surface<K, true> does the same: no d r0ch term.
//   * NH3 + OH: POT_nh3oh discards its analytic gradient and returns FORWARD DIFFERENCES of the energy, step
//     PASO = 1e-5 A, one coordinate after the other, each left at (q + h) - h afterwards (:283-296).  The device does
//     the same 1 + 18 energy evaluations (surface<K, false>: no derivative code at all); a four-lane form for the
//     trajectory kernels gives every lane the base energy and a quarter of the displaced ones (PesNH3OH4).
//     The quotient amplifies one ulp of the energy (1.1e-16 Eh) to 5.9e-12 Eh/bohr, so device and CPU agree to
//     ~1e-10 in this gradient, not to the last bits.
#pragma once
#include "crcl_common.cuh"

namespace crcl {
namespace nh3x {

// NH3 + Cl: BLOCK DATA of egrad_clnh3.f:1846-1898 after the scaling of initialize_clnh3 :1780-1793
// (fact1 = 0.041840: kcal/mol -> 1e5 J/mol, fact2 = 6.022045: mdyn A -> 1e5 J/mol).  Atoms H, N, H, H, Cl.
struct KCl {
    static constexpr int NATOMS = 5, ID = CRCL_PES_CLNH3;
    static constexpr bool HAS_OH = false;
    static constexpr int AN = 1, AB = 4, AH0 = 2, AH1 = 3, AH2 = 0, AO = 0;   // nnc = 2, nnb = 5, nnh = 3, 4, 1
    static constexpr double R0CHR = 1.01410, R0CHP = 1.02700, W1 = 1.00000, W2 = 1.01400;
    static constexpr double D1CH = 119.058 * 0.041840, D3CH = 20.000 * 0.041840;
    static constexpr double A1CH = 2.125000, B1CH = -0.090000, C1CH = 22.00000;
    static constexpr double R0HHR = 1.27730, R0HHP = 1.27730, W3 = 0.0, W4 = 0.0;   // r0hh is a constant here
    static constexpr double D1HH = 109.850 * 0.041840, D3HH = 18.400 * 0.041840, AHH = 1.8600;
    static constexpr double R0CB = 2.10400, D1CB = 65.100 * 0.041840, D3CBI = 16.530 * 0.041840, ACB = 0.7780000;
    static constexpr double A3CB = 0.0, B3CB = 1.0, RCBSP = 0.0;
    static constexpr double APHI = 6.7730500, BPHI = 6.8000000, CPHI = 1.9226100;
    static constexpr double ATHETA = 6.7359700, BTHETA = 6.7000000, CTHETA = 1.9505500;
    static constexpr double FKINF = 0.6950000 * 6.022045, AK = -0.0100000 * 6.022045, BK = 0.1000100;
    static constexpr double AA1 = 3.503370, AA2 = 6.130490, AA3 = 6.100000, AA4 = 3.232430;
    static constexpr double TAU = 1.9022600, TAUNH2 = 1.8046700;
    static constexpr double A1S = 0.0, B1S = 0.0, A2S = 0.0, B2S = 0.0;              // s1, s2 feed only the dead terms
    static constexpr double FKH2OEQ = 0.0, ALPH2O = 0.0, ANGH2OEQ = 0.0;
};

// NH3 + OH: BLOCK DATA of egrad_nh3oh.f:2081-2131 after PREPOT_nh3oh :1993-2009 (fact3 = 2 * 3.1415926 / 360).
// Atoms H, N, H, H, O, H(O).
struct KOH {
    static constexpr int NATOMS = 6, ID = CRCL_PES_NH3OH;
    static constexpr bool HAS_OH = true;
    static constexpr int AN = 1, AB = 4, AH0 = 0, AH1 = 2, AH2 = 3, AO = 5;   // nnc = 2, nnb = 5, nnh = 1, 3, 4, nno = 6
    static constexpr double R0CHR = 1.01417, R0CHP = 1.02777, W1 = 3.00000, W2 = 1.01417;
    static constexpr double D1CH = 125.250 * 0.041840, D3CH = 24.300 * 0.041840;
    static constexpr double A1CH = 2.050000, B1CH = -0.200000, C1CH = 200.4000;
    static constexpr double R0HHR = 0.9710, R0HHP = 0.9595, W3 = 1.00, W4 = 0.973;
    static constexpr double D1HH = 135.250 * 0.041840, D3HH = 32.800 * 0.041840, AHH = 2.0500;
    static constexpr double R0CB = 1.83800, D1CB = 80.900 * 0.041840, D3CBI = 26.700 * 0.041840, ACB = 1.4800000;
    static constexpr double A3CB = 1.60 * 0.041840, B3CB = 0.011, RCBSP = 1.63349;
    static constexpr double APHI = 2.2287900, BPHI = 0.0206600, CPHI = 1.5209900;
    static constexpr double ATHETA = 1.1578700, BTHETA = 0.0358900, CTHETA = 0.7115500;
    static constexpr double FKINF = 0.7100000 * 6.022045, AK = -0.0900000 * 6.022045, BK = 2.7132000;
    static constexpr double AA1 = 0.800000, AA2 = 2.509960, AA3 = 3.506600, AA4 = 1.500000;
    static constexpr double TAU = 1.9022600, TAUNH2 = 1.8046700;
    static constexpr double A1S = 1.5313681e-9, B1S = -1.6696246, A2S = 1.0147402e-9, B2S = -1.363798;   // switchf :1766-1769
    static constexpr double FKH2OEQ = 0.7100000 * 6.022045, ALPH2O = 0.7250;
    static constexpr double ANGH2OEQ = 103.5968 * (2.0 * 3.1415926 / 360.0);
};

CRCL_HD __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// 1 - tanh(x) with the sources' "if (x .lt. 19) ... else 0" and -sech^2(x) (0 beyond the test as well)
CRCL_HD __forceinline__ void sw19(double x, double& omt, double& ms2)
{
    double a, b;
    one_minus_tanh(x, a, b);
    const bool on = x < 19.0;
    omt = on ? a : 0.0;
    ms2 = on ? b : 0.0;
}

// Energy in the routines' 1e5 J/mol from Cartesians x in Angstrom; with GRAD the derivative per Angstrom as the source
// forms it (r0ch a constant).  g may be null without GRAD.
template <class K, bool GRAD>
CRCL_HD __forceinline__ double surface(const double* __restrict__ x, double* __restrict__ g)
{
    constexpr int AH[3] = {K::AH0, K::AH1, K::AH2};
    // ---- distances and bond unit vectors (coorden) ----
    double uc[3][3], ub[3][3], ucb[3], rch[3], rbh[3], irch[3], rcb;
    {
        double t[3], ircb;
#pragma unroll
        for (int d = 0; d < 3; d++) t[d] = x[3 * K::AN + d] - x[3 * K::AB + d];
        sqrt_rsqrt(dot3(t, t), rcb, ircb);
#pragma unroll
        for (int d = 0; d < 3; d++) ucb[d] = t[d] * ircb;
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double tc[3], tb[3], irb;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            tc[d] = x[3 * K::AN + d] - x[3 * AH[i] + d];
            tb[d] = x[3 * K::AB + d] - x[3 * AH[i] + d];
        }
        sqrt_rsqrt(dot3(tc, tc), rch[i], irch[i]);
        sqrt_rsqrt(dot3(tb, tb), rbh[i], irb);
#pragma unroll
        for (int d = 0; d < 3; d++) {
            uc[i][d] = tc[d] * irch[i];
            ub[i][d] = tb[d] * irb;
        }
    }
    // r0ch between reactant and product value ("jcc-2010"); no derivative of it anywhere below, as in the source
    double r0ch;
    {
        double P1 = 1.0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            double omt, ms2;
            sw19(K::W1 * (rch[i] - K::W2), omt, ms2);
            P1 = P1 * omt;
        }
        r0ch = P1 * K::R0CHR + (1.0 - P1) * K::R0CHP;
    }
    double r0hh = K::R0HHR, rno = 0.0, uno[3] = {0.0, 0.0, 0.0};
    if (K::HAS_OH) {
        double t[3], irno, omt, ms2;
#pragma unroll
        for (int d = 0; d < 3; d++) t[d] = x[3 * K::AO + d] - x[3 * K::AB + d];
        sqrt_rsqrt(dot3(t, t), rno, irno);
#pragma unroll
        for (int d = 0; d < 3; d++) uno[d] = t[d] * irno;
        sw19(K::W3 * (rno - K::W4), omt, ms2);
        r0hh = omt * K::R0HHR + (1.0 - omt) * K::R0HHP;
    }
    // ---- switching functions of the three N-H bonds (switchf): sphi, stheta (+ s1, s2 for NH3 + OH) ----
    double sphi[3], sth[3], dsphi[3], dsth[3], s1[3], s2[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double r = rch[i], dr = r - r0ch;
        const bool in = r < 3.8;
        double omt, ms2;
        {
            const double u = r - K::CPHI, ex = CRCL_EXP(K::BPHI * (u * u * u));
            one_minus_tanh(K::APHI * dr * ex, omt, ms2);
            sphi[i] = in ? omt : 0.0;
            if (GRAD) dsphi[i] = in ? K::APHI * (1.0 + 3.0 * K::BPHI * dr * (u * u)) * ex * ms2 : 0.0;
        }
        {
            const double u = r - K::CTHETA, ex = CRCL_EXP(K::BTHETA * (u * u * u));
            one_minus_tanh(K::ATHETA * dr * ex, omt, ms2);
            sth[i] = in ? omt : 0.0;
            if (GRAD) dsth[i] = in ? K::ATHETA * (1.0 + 3.0 * K::BTHETA * dr * (u * u)) * ex * ms2 : 0.0;
        }
        if (K::HAS_OH) {
            const double u = r - K::B1S, u2 = u * u, u4 = u2 * u2;
            sw19(K::A1S * dr * (u4 * u4), s1[i], ms2);
            const double v = r - K::B2S, v2 = v * v;
            sw19(K::A2S * dr * (v2 * v2 * v2), s2[i], ms2);
        }
    }
    // partial derivatives of the energy with respect to the distances
    double dEc[3] = {0.0, 0.0, 0.0}, dEb[3] = {0.0, 0.0, 0.0}, dEcb = 0.0;

    // ---- LEPS-type stretch (stretch) ----
    double vstr = 0.0;
    {
        const double rav = (rch[0] + rch[1] + rch[2]) / 3.0;
        double d3cb = K::D3CBI;
        if (K::HAS_OH) {
            // d3cb switched on the mean N-H length (:752-753): (4 (rav - rcbsp) / b3cb) ** 4.d0
            const double y = 4.0 * (rav - K::RCBSP) / K::B3CB, y2 = y * y;
            d3cb = (K::D3CBI - K::A3CB) + K::A3CB * CRCL_EXP(-(y2 * y2));
        }
        double omt, ms2;
        const double arga = K::C1CH * (rav - r0ch);
        one_minus_tanh(arga, omt, ms2);
        const bool on = arga < 19.0;
        const double ach = on ? K::A1CH + K::B1CH * (2.0 - omt) * 0.5 : K::A1CH + K::B1CH;
        const double dach3 = on ? K::B1CH * K::C1CH * (-ms2) * (0.5 / 3.0) : 0.0;   // d ach / d rch_k
        const double xcb = CRCL_EXP(-K::ACB * (rcb - K::R0CB)), xcb2 = xcb * xcb;
        const double qcb = 0.5 * ((K::D1CB + d3cb) * xcb2 - 2.0 * (K::D1CB - d3cb) * xcb);
        const double jcb = 0.5 * ((K::D1CB - d3cb) * xcb2 - 2.0 * (K::D1CB + d3cb) * xcb);
        double common = 0.0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double xc = CRCL_EXP(-ach * (rch[i] - r0ch)), xc2 = xc * xc;
            const double xb = CRCL_EXP(-K::AHH * (rbh[i] - r0hh)), xb2 = xb * xb;
            const double qch = 0.5 * ((K::D1CH + K::D3CH) * xc2 - 2.0 * (K::D1CH - K::D3CH) * xc);
            const double jch = 0.5 * ((K::D1CH - K::D3CH) * xc2 - 2.0 * (K::D1CH + K::D3CH) * xc);
            const double qbh = 0.5 * ((K::D1HH + K::D3HH) * xb2 - 2.0 * (K::D1HH - K::D3HH) * xb);
            const double jbh = 0.5 * ((K::D1HH - K::D3HH) * xb2 - 2.0 * (K::D1HH + K::D3HH) * xb);
            const double a = jch - jcb, b = jcb - jbh, c = jbh - jch;
            const double s2v = (a * a + b * b + c * c) * 0.5;
            double vj, ivj;
            sqrt_rsqrt(s2v, vj, ivj);
            vstr += (qch + qcb + qbh) - vj;
            if (GRAD) {
                // d(-sqrt(S/2)) = -(1 / (2 sqrt(S/2))) [(a - c) dA + (b - a) dB + (c - b) dC]
                const double f = -0.5 * ivj;
                const double wA = f * (a - c), wB = f * (b - a), wC = f * (c - b);
                const double cq = (K::D1CH + K::D3CH) * xc2 - (K::D1CH - K::D3CH) * xc;
                const double cj = (K::D1CH - K::D3CH) * xc2 - (K::D1CH + K::D3CH) * xc;
                const double w = cq + wA * cj;
                dEc[i] -= ach * w;
                common -= dach3 * (rch[i] - r0ch) * w;
                dEb[i] -= K::AHH * (((K::D1HH + K::D3HH) * xb2 - (K::D1HH - K::D3HH) * xb) +
                                    wC * ((K::D1HH - K::D3HH) * xb2 - (K::D1HH + K::D3HH) * xb));
                dEcb -= K::ACB * (((K::D1CB + d3cb) * xcb2 - (K::D1CB - d3cb) * xcb) +
                                  wB * ((K::D1CB - d3cb) * xcb2 - (K::D1CB + d3cb) * xcb));
            }
        }
        if (GRAD) {
#pragma unroll
            for (int k = 0; k < 3; k++) dEc[k] += common;
        }
        if (K::HAS_OH) {
            const double om = 1.0 - CRCL_EXP(-K::AHH * (rno - r0hh));      // spectator O-H Morse bond (:766-768)
            vstr += K::D1HH * (om * om);
        }
    }

    // ---- harmonic H-N-H bends (ipbend, ipforce, refangles) ----
    double vip = 0.0;
    {
        constexpr double PPITO = (2.0 * 3.141592653589793 - K::TAUNH2) / 2.0;
        constexpr double TP = K::TAU - PPITO, TN = K::TAU - K::TAUNH2;
        double f1[3], df1c[3], df1b[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double e1 = CRCL_EXP(-K::AA1 * rbh[i] * rbh[i]);
            const double db = rbh[i] - r0hh, e2 = CRCL_EXP(-K::AA4 * db * db);
            const double a1 = 1.0 - e1, a2 = K::AA2 + K::AA3 * e2;
            const double dc = rch[i] - r0ch, eg = CRCL_EXP(-a2 * (dc * dc));
            f1[i] = a1 * eg;
            if (GRAD) {
                df1c[i] = -2.0 * dc * a2 * f1[i];
                df1b[i] = (2.0 * K::AA1 * rbh[i] * e1 + 2.0 * K::AA3 * K::AA4 * db * e2 * (dc * dc) * a1) * eg;
            }
        }
        double fko = 0.0, dfko[3] = {0.0, 0.0, 0.0};
        if (!K::HAS_OH) {
            const double d0 = rch[0] - r0ch, d1 = rch[1] - r0ch, d2 = rch[2] - r0ch;
            const double ex = CRCL_EXP(-K::BK * (d0 * d0 + d1 * d1 + d2 * d2));
            fko = K::FKINF + K::AK * ex;
            if (GRAD) {
                dfko[0] = -2.0 * K::AK * K::BK * d0 * ex;
                dfko[1] = -2.0 * K::AK * K::BK * d1 * ex;
                dfko[2] = -2.0 * K::AK * K::BK * d2 * ex;
            }
        }
        double gH[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};     // angle part, straight to the hydrogens
#pragma unroll
        for (int p = 0; p < 3; p++) {
            const int i = (p == 2) ? 1 : 0, j = (p == 0) ? 1 : 2, k = 3 - i - j;     // pairs (0,1), (0,2), (1,2)
            const double c = dot3(uc[i], uc[j]);
            const double th = CRCL_ACOS(c);
            const double th0 = K::TAU + TP * (sphi[i] * sphi[j] - 1.0) + TN * (sth[k] - 1.0);
            const double dth = th - th0;
            double fk = fko;
            if (K::HAS_OH) {
                constexpr double F0 = K::FKINF + K::AK, F2 = K::FKINF;
                fk = F0 + F0 * (s1[i] * s1[j] - 1.0) + (F0 - F2) * (s2[k] - 1.0);
            }
            const double F = fk * f1[i] * f1[j];
            vip += 0.5 * F * (dth * dth);
            if (GRAD) {
                const double hd2 = 0.5 * (dth * dth), Fd = F * dth;
                // through F = fko f1_i f1_j and through the reference angle
                dEc[i] += hd2 * (dfko[i] * f1[i] * f1[j] + fko * df1c[i] * f1[j]) - Fd * TP * dsphi[i] * sphi[j];
                dEc[j] += hd2 * (dfko[j] * f1[i] * f1[j] + fko * f1[i] * df1c[j]) - Fd * TP * sphi[i] * dsphi[j];
                dEc[k] += hd2 * (dfko[k] * f1[i] * f1[j]) - Fd * TN * dsth[k];
                dEb[i] += hd2 * fko * df1b[i] * f1[j];
                dEb[j] += hd2 * fko * f1[i] * df1b[j];
                // through the angle: d theta / d H_i = (uc_j - c uc_i) / (r_i sin theta)   (uc points from H to N)
                const double is = CRCL_RSQRT(1.0 - c * c);
                const double wi = Fd * is * irch[i], wj = Fd * is * irch[j];
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    gH[i][d] += wi * (uc[j][d] - c * uc[i][d]);
                    gH[j][d] += wj * (uc[i][d] - c * uc[j][d]);
                }
            }
        }
        if (K::HAS_OH) {
            // H_i-O-H(O) bends of the forming water, force constant switched off with r(O-H_i) (:1213-1248)
#pragma unroll
            for (int i = 0; i < 3; i++) {
                double cs = -dot3(uno, ub[i]);
                cs = fmin(1.0, fmax(-1.0, cs));
                const double ang = CRCL_ACOS(cs);
                double omt, ms2;
                sw19(K::ALPH2O * (rbh[i] - r0hh), omt, ms2);
                const double dang = ang - K::ANGH2OEQ;
                vip += 0.5 * (K::FKH2OEQ * omt) * dang * dang;
            }
        }
        if (GRAD) {
            // chain rule to the atoms: r(N-H_i) along uc_i, r(b-H_i) along ub_i, r(N-b) along ucb
#pragma unroll
            for (int d = 0; d < 3; d++) {
                double gn = dEcb * ucb[d], gb = -dEcb * ucb[d];
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    const double a = dEc[i] * uc[i][d], b = dEb[i] * ub[i][d];
                    g[3 * AH[i] + d] = gH[i][d] - a - b;
                    gn += a - gH[i][d];
                    gb += b;
                }
                g[3 * K::AN + d] = gn;
                g[3 * K::AB + d] = gb;
            }
        }
    }
    return vstr + vip;
}

}  // namespace nh3x

// NH3 + Cl: one thread per image, analytic gradient
struct PesClNH3 {
    using K = nh3x::KCl;
    static constexpr int NATOMS = K::NATOMS;
    static constexpr int ID = K::ID;
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
    // q, g: [NATOMS][3] bohr, hartree/bohr; unit factors of POT_clnh3 :155, :192, :201
    CRCL_HD static __forceinline__ int eval(const double* __restrict__ q, double& V, double* __restrict__ g)
    {
        double x[NOWN], ga[NOWN];
#pragma unroll
        for (int c = 0; c < NOWN; c++) x[c] = q[c] * 0.52918;
        const double en = nh3x::surface<K, true>(x, ga);
        V = en * 0.03812;
#pragma unroll
        for (int c = 0; c < NOWN; c++) g[c] = ga[c] * 0.0201723;
        return 0;
    }
};

// NH3 + OH: one thread per image, the reference's forward-difference gradient (POT_nh3oh :283-296)
struct PesNH3OH {
    using K = nh3x::KOH;
    static constexpr int NATOMS = K::NATOMS;
    static constexpr int ID = K::ID;
    static constexpr int LANES = 1;
    static constexpr int NOWN = 3 * NATOMS;
    static constexpr double PASO = 1.0e-5;
    CRCL_HD static __forceinline__ int owned(int, int k) { return k; }
    template <class QF>
    CRCL_HD static __forceinline__ int eval_coop(QF qf, int, unsigned, double& V, double* gown)
    {
        double x[NOWN];
#pragma unroll
        for (int c = 0; c < NOWN; c++) x[c] = qf(c);
        return eval(x, V, gown);
    }
    // energy (hartree) at the Angstrom coordinates xa with coordinate I displaced and every J < I left at (x + h) - h
    CRCL_HD static __forceinline__ double displaced(const double* __restrict__ xa, int I)
    {
        double y[NOWN];
#pragma unroll
        for (int c = 0; c < NOWN; c++) {
            const double up = xa[c] + PASO;
            y[c] = (c < I) ? up - PASO : ((c == I) ? up : xa[c]);
        }
        double en = nh3x::surface<K, false>(y, nullptr);
        en = en * 0.03812;
        return en;
    }
    CRCL_HD static __forceinline__ int eval(const double* __restrict__ q, double& V, double* __restrict__ g)
    {
        double xa[NOWN];
#pragma unroll
        for (int c = 0; c < NOWN; c++) xa[c] = q[c] * 0.52918;
        const double e0 = displaced(xa, -1);
        V = e0;
#pragma unroll 1
        for (int I = 0; I < NOWN; I++) {
            const double en = displaced(xa, I);
            double d = (en - e0) / PASO;
            d = d * 0.52918;
            g[I] = d;
        }
        return 0;
    }
};

#ifdef __CUDACC__
// NH3 + OH in the trajectory kernels: the four lanes of a bead share the 18 displaced energies (lane x takes the
// coordinates x, x + 4, ...: 5, 5, 4, 4) and each evaluates the base energy itself, so nothing travels between the
// lanes: 6 energy evaluations per lane instead of 19 per thread.
struct PesNH3OH4 {
    using K = nh3x::KOH;
    static constexpr int NATOMS = K::NATOMS;
    static constexpr int ID = K::ID;
    static constexpr int LANES = 4;
    static constexpr int NOWN = 5;
    __device__ static __forceinline__ int owned(int x, int k)
    {
        const int c = x + 4 * k;
        return (c < 3 * NATOMS) ? c : -1;
    }
    template <class QF>
    __device__ static __forceinline__ int eval_coop(QF qf, int x, unsigned, double& V, double gown[NOWN], double* = nullptr)
    {
        constexpr int NC = 3 * NATOMS;
        double xa[NC];
#pragma unroll
        for (int c = 0; c < NC; c++) xa[c] = qf(c) * 0.52918;
        const double e0 = PesNH3OH::displaced(xa, -1);
        V = (x == 0) ? e0 : 0.0;
#pragma unroll 1
        for (int k = 0; k < NOWN; k++) {
            const int I = x + 4 * k;
            double d = 0.0;
            if (I < NC) {
                const double en = PesNH3OH::displaced(xa, I);
                d = (en - e0) / PesNH3OH::PASO;
                d = d * 0.52918;
            }
            // gown[k] with k a loop variable of a rolled loop: selects keep the array in registers
#pragma unroll
            for (int kk = 0; kk < NOWN; kk++) gown[kk] = (kk == k) ? d : gown[kk];
        }
        return 0;
    }
};
#endif

}  // namespace crcl
