// pes_ch4h.cuh -- CBE-2009 CH4 + H -> CH3 + H2 surface (Corchado, Bravo, Espinosa-Garcia,
// J. Chem. Phys. 130, 184314), one thread per image, FP64.
//
// Replaces /root/reference/src/egrad_ch4h.f: egrad_ch4h :74-132, POT_ch4h :161-285,
// coorden :287, refangles :367, stretch :506, opbend :713, ipbend :865, calcdelta :986,
// opforce :1240, ipforce :1350, switchf :1545, constants :1852-1886 scaled once as
// PREPOT_ch4h :1781-1793 does.  The POTLIB wrappers in util_ch4h.f are identity maps for the
// flags Caracal sets (NFLAG(1)=NFLAG(2)=1, ICARTR=1) and have no counterpart here.
//
// This is a re-derivation for the GPU, not a transcription.  The reference scatters
// Cartesian derivatives term by term into pdot(150) through index tables and re-evaluates
// the same exponentials inside triple loops (~370 libm calls per image).  Here every term
// is differentiated once with respect to the eleven distances it depends on
// (rch[4], rbh[4], rcb) plus explicit vector parts for the two angle types, and the chain
// rule to Cartesians is applied once at the end (~40 libm calls per image).  Results agree
// with the literal oracle to rounding (tests/test_pes_parity.py).
//
// Units as in the reference: bohr -> Angstrom by *0.52918 (:231); energy 1e5 J/mol ->
// hartree by *0.03812 (:267); gradient by *0.0201723 (:276).  NB 0.03812*0.52918 =
// 0.02017234..., so the reference's gradient is 2e-6 (relative) inconsistent with its
// energy; reproduced, not fixed.  No clamps on acos / 1/sqrt(1-x^2) arguments (:939-947,
// :1088, :1218): collinear/planar arrangements give NaN as in the reference.
#pragma once
#include "crcl_common.cuh"

namespace crcl {
namespace ch4h {

// BLOCK DATA PTPACM_ch4h after PREPOT scaling (fact1 = 0.041840, fact2 = 6.022045)
constexpr double R0CH = 1.08898, A1CH = 1.78374, B1CH = 0.14201, C1CH = 2.21773;
constexpr double R0HH = 0.74239, AHH = 1.9589, R0CB = 1.08898, ACB = 1.4173200;
constexpr double D1CH = 111.266 * 0.041840, D3CH = 48.96226 * 0.041840;
constexpr double D1HH = 108.382 * 0.041840, D3HH = 38.42657 * 0.041840;
constexpr double D1CB = 56.505 * 0.041840, D3CB = 19.612 * 0.041840;
constexpr double A3S = 0.1475300, B3S = -2.9926300;
constexpr double APHI = 0.5307000, BPHI = 0.4012200, CPHI = 1.9235100;
constexpr double ATHETA = 0.9119800, BTHETA = 0.3537500, CTHETA = 1.8970500;
constexpr double FCH3 = 0.0693700 * 6.022045, HCH3 = 0.1387400 * 6.022045;
constexpr double FKINF = 0.4291400 * 6.022045, AK = 0.1353000 * 6.022045;
constexpr double AA1 = 1.265960, AA2 = 0.000710, AA3 = 0.985920, AA4 = 2.785060;
// switchf_ch4h :1598-1601
constexpr double A1S = 1.5132681e-7, B1S = -4.3792246, A2S = 1.9202402e-7, B2S = -12.323018;
constexpr double PI = 3.141592653589793;
constexpr double TAU_TET = 1.9106332362490186;   // acos(-1/3), the tetrahedral angle (refangles_ch4h)

// Morse-like singlet/triplet pair for one bond: vq, vj and their r- and a-derivatives
struct Leps {
    double vq, vj, dvq, dvj;  // d/dr
};
CRCL_HD __forceinline__ Leps leps(double d1, double d3, double a, double dr)
{
    const double X1 = CRCL_EXP(-a * dr), X2 = X1 * X1;
    Leps o;
    o.vq = 0.5 * ((d1 + d3) * X2 - 2.0 * (d1 - d3) * X1);
    o.vj = 0.5 * ((d1 - d3) * X2 - 2.0 * (d1 + d3) * X1);
    o.dvq = -a * ((d1 + d3) * X2 - (d1 - d3) * X1);
    o.dvj = -a * ((d1 - d3) * X2 - (d1 + d3) * X1);
    return o;
}

CRCL_HD __forceinline__ void cross(const double a[3], const double b[3], double c[3])
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
CRCL_HD __forceinline__ double dot(const double a[3], const double b[3])
{
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

}  // namespace ch4h

namespace ch4h {
// the namespace constants above as the constants struct of the one-lane template
struct K6 {
    static constexpr int NATOMS = 6, ID = CRCL_PES_CH4H;
    static constexpr bool HAS_OH = false;
    static constexpr double R0CH = ch4h::R0CH;
    static constexpr double A1CH = ch4h::A1CH;
    static constexpr double B1CH = ch4h::B1CH;
    static constexpr double C1CH = ch4h::C1CH;
    static constexpr double R0HH = ch4h::R0HH;
    static constexpr double AHH = ch4h::AHH;
    static constexpr double R0CB = ch4h::R0CB;
    static constexpr double ACB = ch4h::ACB;
    static constexpr double D1CH = ch4h::D1CH;
    static constexpr double D3CH = ch4h::D3CH;
    static constexpr double D1HH = ch4h::D1HH;
    static constexpr double D3HH = ch4h::D3HH;
    static constexpr double D1CB = ch4h::D1CB;
    static constexpr double D3CB = ch4h::D3CB;
    static constexpr double A3S = ch4h::A3S;
    static constexpr double B3S = ch4h::B3S;
    static constexpr double APHI = ch4h::APHI;
    static constexpr double BPHI = ch4h::BPHI;
    static constexpr double CPHI = ch4h::CPHI;
    static constexpr double ATHETA = ch4h::ATHETA;
    static constexpr double BTHETA = ch4h::BTHETA;
    static constexpr double CTHETA = ch4h::CTHETA;
    static constexpr double FCH3 = ch4h::FCH3;
    static constexpr double HCH3 = ch4h::HCH3;
    static constexpr double FKINF = ch4h::FKINF;
    static constexpr double AK = ch4h::AK;
    static constexpr double AA1 = ch4h::AA1;
    static constexpr double AA2 = ch4h::AA2;
    static constexpr double AA3 = ch4h::AA3;
    static constexpr double AA4 = ch4h::AA4;
    static constexpr double A1S = ch4h::A1S;
    static constexpr double B1S = ch4h::B1S;
    static constexpr double A2S = ch4h::A2S;
    static constexpr double B2S = ch4h::B2S;
    static constexpr double FKH2OEQ = 0.0, ALPH2O = 0.0, ANH2OEQ = 0.0;   // unused without HAS_OH
    static constexpr double MORSE_R0 = 0.0, MORSE_A = 0.0, MORSE_D = 0.0;
    static constexpr bool SPHI_ANY_R = false;
    static constexpr double TAU_PLANAR = 0.5;
};
}  // namespace ch4h

// K: the constants of one member of the CBE family (ch4h::K6 below, ch4oh::K7 in pes_ch4oh.cuh).  With K::HAS_OH the
// abstracting atom is the oxygen of an OH radical (7 atoms) and the three terms egrad_ch4oh.f adds to the template
// are evaluated too: O-H Morse bond (:616-659, :769-774), four H-O-H bends (:1059-1171); the switched C-O triplet
// depth (:590-592) is the constant D3CB because the shipped BLOCK DATA has a3cb = 0 (:2104).
template <class K>
struct PesCBE1 {
    static constexpr int NATOMS = K::NATOMS;
    static constexpr int ID = K::ID;
    // one-lane trajectory interface (traj_kernel.cuh): the thread of a bead owns all 3 NATOMS components
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

    // q, g: [atom][xyz] in bohr / hartree bohr^-1, atom order H, C, H, H, H, H_b
    // (nnc=2, nnb=6, nnh=3,4,5,1; egrad_ch4h.f:1852-1854)
    CRCL_HD static __forceinline__ int eval(const double* __restrict__ q, double& V,
                                               double* __restrict__ g)
    {
        using namespace ch4h;
        constexpr int HA[4] = {2, 3, 4, 0};  // 0-based atom index of methane hydrogen x
        constexpr int CA = 1, BA = 5;
        double c[4][3], ubh[4][3], ucb[3];  // unit vectors H-C, H-Hb, Hb-C
        double rch[4], irch[4], rbh[4], rcb;
        {
            double t[3];
#pragma unroll
            for (int d = 0; d < 3; d++) t[d] = q[3 * BA + d] * 0.52918 - q[3 * CA + d] * 0.52918;
            double ircb;
            sqrt_rsqrt(dot(t, t), rcb, ircb);
#pragma unroll
            for (int d = 0; d < 3; d++) ucb[d] = t[d] * ircb;
#pragma unroll
            for (int x = 0; x < 4; x++) {
                double tc[3], tb[3];
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    // the reference converts each coordinate first (q(I)=R(I)*0.52918) and
                    // differences afterwards; do the same so near-cancelling differences agree
                    const double h = q[3 * HA[x] + d] * 0.52918;
                    tc[d] = h - q[3 * CA + d] * 0.52918;
                    tb[d] = h - q[3 * BA + d] * 0.52918;
                }
                double irbh;
                sqrt_rsqrt(dot(tc, tc), rch[x], irch[x]);
                sqrt_rsqrt(dot(tb, tb), rbh[x], irbh);
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    c[x][d] = tc[d] * irch[x];
                    ubh[x][d] = tb[d] * irbh;
                }
            }
        }
        double tno[3] = {0, 0, 0}, rno = 1.0, irno = 1.0;   // H(O) - O in Angstrom (coorden_ch4oh :351,:361)
        if constexpr (K::HAS_OH) {
#pragma unroll
            for (int d = 0; d < 3; d++) tno[d] = q[3 * 6 + d] * 0.52918 - q[3 * BA + d] * 0.52918;
            sqrt_rsqrt(dot(tno, tno), rno, irno);
        }
        double gO[3] = {0, 0, 0}, gBx[3] = {0, 0, 0};   // explicit vector parts on H(O) and on the abstracting atom
        // accumulators: dV/d(distance) and explicit vector parts
        double Dch[4] = {0, 0, 0, 0}, Dbh[4] = {0, 0, 0, 0}, Dcb = 0.0;
        double gH[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, gC[3] = {0, 0, 0};
        double en = 0.0;

        // ---- switching functions (switchf_ch4h) ----
        double s1[4], ds1[4], s2[4], ds2[4], s3[4], ds3[4], sphi[4], dsphi[4], sth[4], dsth[4];
#pragma unroll
        for (int x = 0; x < 4; x++) {
            const double r = rch[x], dr = r - K::R0CH;
            double omt, ms2;
            {
                const double u = r - K::B1S, u2 = u * u, u4 = u2 * u2, u7 = u4 * u2 * u, u8 = u4 * u4;
                const double arg = K::A1S * dr * u8;
                if (arg < 19.0) {
                    one_minus_tanh(arg, omt, ms2);
                    s1[x] = omt;
                    ds1[x] = K::A1S * (u8 + 8.0 * dr * u7) * ms2;
                } else {
                    s1[x] = 0.0;
                    ds1[x] = 0.0;
                }
            }
            {
                const double u = r - K::B2S, u2 = u * u, u4 = u2 * u2, u5 = u4 * u, u6 = u4 * u2;
                const double arg = K::A2S * dr * u6;
                if (arg < 19.0) {
                    one_minus_tanh(arg, omt, ms2);
                    s2[x] = omt;
                    ds2[x] = K::A2S * (u6 + 6.0 * dr * u5) * ms2;
                } else {
                    s2[x] = 0.0;
                    ds2[x] = 0.0;
                }
            }
            {
                const double u = r - K::B3S;
                const double arg = K::A3S * dr * u * u;
                if (arg < 19.0) {
                    one_minus_tanh(arg, omt, ms2);
                    s3[x] = omt;
                    ds3[x] = K::A3S * (3.0 * r * r - 2.0 * r * (K::R0CH + 2.0 * K::B3S) + K::B3S * (K::B3S + 2.0 * K::R0CH)) * ms2;
                } else {
                    s3[x] = 0.0;
                    ds3[x] = 0.0;
                }
            }
            if (K::SPHI_ANY_R || r < 3.8) {   // egrad_geh4oh.f:1793-1804 drops the cut for sphi only
                const double u = r - K::CPHI, ex = CRCL_EXP(K::BPHI * u * u * u);
                one_minus_tanh(K::APHI * dr * ex, omt, ms2);
                sphi[x] = omt;
                dsphi[x] = K::APHI * (1.0 + 3.0 * K::BPHI * dr * u * u) * ex * ms2;
            } else {
                sphi[x] = 0.0;
                dsphi[x] = 0.0;
            }
            if (r < 3.8) {
                const double v = r - K::CTHETA, ev = CRCL_EXP(K::BTHETA * v * v * v);
                one_minus_tanh(K::ATHETA * dr * ev, omt, ms2);
                sth[x] = omt;
                dsth[x] = K::ATHETA * (1.0 + 3.0 * K::BTHETA * dr * v * v) * ev * ms2;
            } else {
                sth[x] = 0.0;
                dsth[x] = 0.0;
            }
        }
        // reference angles theta0(i,j) = tau + ta (sphi_i sphi_j - 1) + tb (sth_k sth_l - 1)
        // with {k,l} the complement of {i,j} (refangles_ch4h)
        const double tau = TAU_TET;
        const double ta = tau - K::TAU_PLANAR * PI, tb = tau - 2.0 * PI / 3.0;   // halfpi, or taugeh (egrad_geh4oh.f:368)
        auto theta0 = [&](int i, int j, int k, int l) {
            return tau + ta * (sphi[i] * sphi[j] - 1.0) + tb * (sth[k] * sth[l] - 1.0);
        };

        // ---- stretching (stretch_ch4h): LEPS for each (C-H_i, C-H_b, H_b-H_i) triple ----
        {
            const double rav = (rch[0] + rch[1] + rch[2] + rch[3]) / 4.0;
            const double arga = K::C1CH * (rav - K::R0CH);
            double ach, dach;  // dach = d ach / d rch_x (same for every x)
            if (arga < 19.0) {
                double omt, ms2;
                one_minus_tanh(arga, omt, ms2);
                ach = K::A1CH + K::B1CH * (2.0 - omt) * 0.5;
                dach = -K::B1CH * K::C1CH * 0.5 * ms2 * 0.25;
            } else {
                ach = K::A1CH + K::B1CH;
                dach = 0.0;
            }
            const Leps cb = leps(K::D1CB, K::D3CB, K::ACB, rcb - K::R0CB);
            double Dach = 0.0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const double dr = rch[i] - K::R0CH;
                const Leps ch = leps(K::D1CH, K::D3CH, ach, dr);
                const Leps bh = leps(K::D1HH, K::D3HH, K::AHH, rbh[i] - K::R0HH);
                const double a = ch.vj, b = cb.vj, cc = bh.vj;
                double vjs, ivjs;
                sqrt_rsqrt((sqr(a - b) + sqr(b - cc) + sqr(cc - a)) * 0.5, vjs, ivjs);
                const double vj = -vjs;
                en += ch.vq + cb.vq + bh.vq + vj;
                const double h = -0.5 * ivjs;
                const double wa = (2.0 * a - b - cc) * h, wb = (2.0 * b - a - cc) * h,
                             wc = (2.0 * cc - a - b) * h;
                const double dch = ch.dvq + wa * ch.dvj;
                Dch[i] += dch;
                Dbh[i] += bh.dvq + wc * bh.dvj;
                Dcb += cb.dvq + wb * cb.dvj;
                // d/d(ach): CRCL_EXP(-a dr) depends on a exactly as on r with dr/a swapped
                Dach += dch * CRCL_DIV(dr, ach);
            }
            const double t = Dach * dach;
#pragma unroll
            for (int x = 0; x < 4; x++) Dch[x] += t;
        }

        // ---- out-of-plane bending (opbend_ch4h, calcdelta_ch4h, opforce_ch4h) ----
        {
            const double s3p = s3[0] * s3[1] * s3[2] * s3[3];
            (void)s3p;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int j = (i + 1) & 3, k = (i + 2) & 3, l = (i + 3) & 3;
                const double pj = s3[j], pk = s3[k], pl = s3[l];
                const double sw = (1.0 - s3[i]) * pj * pk * pl;
                const double fd = sw * K::FCH3, hd = sw * K::HCH3;
                // a = H_k - H_j, b = H_l - H_j (Angstrom), from unit vectors and lengths
                double a[3], b[3], n[3];
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    const double hj = c[j][d] * rch[j];
                    a[d] = c[k][d] * rch[k] - hj;
                    b[d] = c[l][d] * rch[l] - hj;
                }
                cross(a, b, n);
                const double inn = CRCL_RSQRT(dot(n, n));
                double u[3];
                u[0] = dot(n, c[j]) * inn;
                u[1] = dot(n, c[k]) * inn;
                u[2] = dot(n, c[l]) * inn;
                // right-handedness: each positive argd toggles k<->l (:821-833); an odd count
                // leaves them swapped, which flips the sign of a x b
                const int npos = (u[0] > 0.0) + (u[1] > 0.0) + (u[2] > 0.0);
                const double sg = (npos & 1) ? -1.0 : 1.0;
                double nh[3] = {sg * n[0] * inn, sg * n[1] * inn, sg * n[2] * inn};
                const int m3[3] = {j, k, l};
                // complement pairs of (i,m): for m=j -> (k,l), m=k -> (j,l), m=l -> (j,k)
                const int ca[3] = {k, j, j}, cb2[3] = {l, l, k};
                double sum2 = 0.0, sum4 = 0.0;
                double G[3] = {0, 0, 0};
#pragma unroll
                for (int t = 0; t < 3; t++) {
                    const int m = m3[t];
                    const double um = sg * u[t];
                    const double del = CRCL_ACOS(um) - theta0(i, m, ca[t], cb2[t]);
                    const double d2 = del * del;
                    sum2 += d2;
                    sum4 += d2 * d2;
                    const double w = 2.0 * fd * del + 4.0 * hd * d2 * del;
                    const double wu = -w * CRCL_RSQRT(1.0 - um * um);  // dV/du_m
                    // u_m = nhat . chat_m : direct part through chat_m
                    const double f = wu * irch[m];
#pragma unroll
                    for (int d = 0; d < 3; d++) {
                        const double v = f * (nh[d] - um * c[m][d]);
                        gH[m][d] += v;
                        gC[d] -= v;
                        G[d] += wu * (c[m][d] - um * nh[d]);
                    }
                    // through theta0(i,m): -w * dtheta0(i,m,x)
                    Dch[i] -= w * ta * dsphi[i] * sphi[m];
                    Dch[m] -= w * ta * sphi[i] * dsphi[m];
                    Dch[ca[t]] -= w * tb * dsth[ca[t]] * sth[cb2[t]];
                    Dch[cb2[t]] -= w * tb * sth[ca[t]] * dsth[cb2[t]];
                }
                // through nhat: d(n_eff)/dH_k = sg * (b x .), dH_l = sg * (. x a)
                {
                    const double s = sg * inn;
                    double gk[3], gl[3];
                    cross(b, G, gk);
                    cross(G, a, gl);
#pragma unroll
                    for (int d = 0; d < 3; d++) {
                        gH[k][d] += s * gk[d];
                        gH[l][d] += s * gl[d];
                        gH[j][d] -= s * (gk[d] + gl[d]);
                    }
                }
                en += fd * sum2 + hd * sum4;
                // force-constant derivatives (opforce_ch4h)
                const double fs = K::FCH3 * sum2 + K::HCH3 * sum4;
                Dch[i] -= fs * ds3[i] * pj * pk * pl;
                Dch[j] += fs * (1.0 - s3[i]) * ds3[j] * pk * pl;
                Dch[k] += fs * (1.0 - s3[i]) * pj * ds3[k] * pl;
                Dch[l] += fs * (1.0 - s3[i]) * pj * pk * ds3[l];
            }
        }

        // ---- in-plane bending (ipbend_ch4h, ipforce_ch4h) ----
        {
            constexpr double f0 = K::FKINF + K::AK, f2 = K::FKINF;
            double f1[4], df1c[4], df1h[4];
#pragma unroll
            for (int x = 0; x < 4; x++) {
                const double dr = rch[x] - K::R0CH, dh = rbh[x] - K::R0HH;
                const double e1 = CRCL_EXP(-K::AA1 * rbh[x] * rbh[x]);
                const double e2 = CRCL_EXP(-K::AA4 * dh * dh);
                const double a1 = 1.0 - e1;
                const double a2 = K::AA2 + K::AA3 * e2;
                const double E = CRCL_EXP(-a2 * dr * dr);
                f1[x] = a1 * E;
                df1c[x] = -2.0 * dr * a1 * a2 * E;
                df1h[x] = 2.0 * K::AA1 * rbh[x] * e1 * E + 2.0 * K::AA3 * K::AA4 * dh * e2 * dr * dr * a1 * E;
            }
            constexpr int PI_[6] = {0, 0, 0, 1, 1, 2}, PJ_[6] = {1, 2, 3, 2, 3, 3};
            constexpr int PK_[6] = {2, 1, 1, 0, 0, 0}, PL_[6] = {3, 3, 2, 3, 2, 1};
#pragma unroll
            for (int p = 0; p < 6; p++) {
                const int i = PI_[p], j = PJ_[p], k = PK_[p], l = PL_[p];
                const double fk0 = f0 + f0 * (s1[i] * s1[j] - 1.0) + (f0 - f2) * (s2[k] * s2[l] - 1.0);
                const double ff = f1[i] * f1[j];
                const double Kf = fk0 * ff;
                const double cs = dot(c[i], c[j]);
                const double del = CRCL_ACOS(cs) - theta0(i, j, k, l);
                en += 0.5 * Kf * del * del;
                const double w = Kf * del;                 // dV/d(delta)
                const double wc = -w * CRCL_RSQRT(1.0 - cs * cs);  // dV/d(cos)
                const double fi = wc * irch[i], fj = wc * irch[j];
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    const double vi = fi * (c[j][d] - cs * c[i][d]);
                    const double vj = fj * (c[i][d] - cs * c[j][d]);
                    gH[i][d] += vi;
                    gH[j][d] += vj;
                    gC[d] -= vi + vj;
                }
                const double hd2 = 0.5 * del * del;
                // theta0 and force-constant dependence on rch
                Dch[i] += -w * ta * dsphi[i] * sphi[j] + hd2 * (f0 * ds1[i] * s1[j] * ff + fk0 * df1c[i] * f1[j]);
                Dch[j] += -w * ta * sphi[i] * dsphi[j] + hd2 * (f0 * s1[i] * ds1[j] * ff + fk0 * f1[i] * df1c[j]);
                Dch[k] += -w * tb * dsth[k] * sth[l] + hd2 * (f0 - f2) * ds2[k] * s2[l] * ff;
                Dch[l] += -w * tb * sth[k] * dsth[l] + hd2 * (f0 - f2) * s2[k] * ds2[l] * ff;
                Dbh[i] += hd2 * fk0 * df1h[i] * f1[j];
                Dbh[j] += hd2 * fk0 * f1[i] * df1h[j];
            }
        }

        if constexpr (K::HAS_OH) {
            // ---- O-H Morse bond (stretch_ch4oh :616-621, :652-659, :769-774) ----
            const double ex = CRCL_EXP(-K::MORSE_A * (rno - K::MORSE_R0)), om = 1.0 - ex;
            en += K::MORSE_D * (om * om);
            const double de = 2.0 * K::MORSE_A * K::MORSE_D * om * ex * irno;
#pragma unroll
            for (int d = 0; d < 3; d++) {
                gBx[d] -= de * tno[d];
                gO[d] += de * tno[d];
            }
            // ---- H_i-O-H(O) bends, force constant switched off with r(O-H_i) (ipbend_ch4oh :1059-1171) ----
#pragma unroll
            for (int i = 0; i < 4; i++) {
                double tb[3];   // the reference's tbh(i,:) = O - H_i
#pragma unroll
                for (int d = 0; d < 3; d++) tb[d] = -ubh[i][d] * rbh[i];
                double cs = -dot(tno, tb) * CRCL_RCP(rno * rbh[i]);
                cs = fmin(1.0, fmax(-1.0, cs));
                const double dang = CRCL_ACOS(cs) - K::ANH2OEQ;
                const double arga = K::ALPH2O * (rbh[i] - K::R0HH);
                double omt, ms2;
                one_minus_tanh(arga, omt, ms2);
                const double fk = (arga < 19.0) ? K::FKH2OEQ * omt : 0.0;
                en += 0.5 * fk * dang * dang;
                Dbh[i] += K::FKH2OEQ * K::ALPH2O * ms2 * (0.5 * dang * dang);   // dkdr is not guarded (:1104-1107)
                const double dstda = fk * dang;
                double pv[3], v1[3], v2[3];
                cross(tno, tb, pv);
                double rp = CRCL_SQRT(fmax(dot(pv, pv), 1.0e-300));
                if (rp < 1.0e-6) rp = 1.0e-6;
                const double terma = dstda * CRCL_RCP(rbh[i] * rbh[i] * rp), termc = dstda * CRCL_RCP(rno * rno * rp);
                cross(tb, pv, v1);
                cross(tno, pv, v2);
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    const double hi = -terma * v1[d], ho = -termc * v2[d];
                    gH[i][d] += hi;
                    gO[d] += ho;
                    gBx[d] += -hi - ho;
                }
            }
        }

        // ---- chain rule to Cartesians, unit conversion ----
        V = en * 0.03812;
        constexpr double GF = 0.0201723;
        double gB[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            double cC = gC[d] - Dcb * ucb[d];
            double cB = Dcb * ucb[d];
            if constexpr (K::HAS_OH) cB += gBx[d];
#pragma unroll
            for (int x = 0; x < 4; x++) {
                const double vc = Dch[x] * c[x][d], vb = Dbh[x] * ubh[x][d];
                g[3 * HA[x] + d] = (gH[x][d] + vc + vb) * GF;
                cC -= vc;
                cB -= vb;
            }
            g[3 * CA + d] = cC * GF;
            gB[d] = cB * GF;
            g[3 * BA + d] = gB[d];
            if constexpr (K::HAS_OH) g[3 * 6 + d] = gO[d] * GF;
        }
        return 0;
    }
};

using PesCH4H = PesCBE1<ch4h::K6>;

}  // namespace crcl

// =================================================================================================
// Four-lane cooperative evaluation (device only).
//
// ncu on the one-thread-per-bead kernel (profiles/r1_recross_ch4h_nb16.md) shows the BASELINE
// batch (1024 trajectories x 16 beads = 512 warps on 148 SMs) is bound by instruction-cache misses
// and dependent-issue latency at 3.5 warps/SM, not by the FP64 pipe.  The surface is a sum over
// the four methane hydrogens, so four adjacent lanes evaluate one image: lane x owns hydrogen x,
// works in a frame rotated by x (local index t <-> hydrogen (x+t)&3, all register indices stay
// compile-time), exchanges the per-hydrogen scalars with width-4 shuffles, and the partial
// derivative accumulators are reduce-scattered back to the owner lane by the inverse rotation.
// 4x the warps, ~1/4 of the instruction stream per warp.
// =================================================================================================
#ifdef __CUDACC__
namespace crcl {

// K as for PesCBE1: ch4h::K6 (6 atoms), or a 7-atom member with K::HAS_OH, whose added terms split the same way --
// lane x evaluates the H_x-O-H(O) bend of its own hydrogen, lane 0 the O-H(O) Morse bond.
template <class K>
struct PesCBE4 {
    static constexpr int NATOMS = K::NATOMS;
    static constexpr int ID = K::ID;
    static constexpr int LANES = 4;
    static constexpr int NOWN = K::HAS_OH ? 6 : 5;  // components owned per lane (5,5,4,4 of 18; 6,5,5,5 of 21)
    static constexpr bool SPREAD_OK = true;          // one-bead trajectories may run as eight quads (PesSpreadQ, traj_inst.cuh)
#ifndef CRCL_CBE_SHFL_GATHER
    // per-hydrogen quantities travel between the four lanes of a bead through shared memory: 18 doubles per lane,
    // written as nine 16-byte stores and read back as 16-byte loads per neighbour (instead of the 102 SHFL of the 51
    // doubles a lane gathers), in two phases so that the in-plane term's share is not live through the out-of-plane
    // term; the lane's own late-use values (its switching functions, the three unit vectors of the final chain rule)
    // are parked in the same block and re-read where they are needed: 24 doubles, stride 26 (conflict-free 16-byte
    // accesses).  Measured: 12.77 -> 12.18 ms per 1000 child steps, spill traffic 748 -> 77 MB (profiles/r2m_*).
    // -DCRCL_CBE_SHFL_GATHER restores the shuffles (A/B builds).
    static constexpr int COOP_SCRATCH = 26;
#endif

    // component (atom*3+xyz) number k owned by lane x, or -1: the lane's hydrogen, then C and H_b
    // spread as lane0: Cx,Cy  lane1: Cz,Bx  lane2: By  lane3: Bz; with H(O): lane2: x  lane3: y  lane0: z
    __device__ static __forceinline__ int owned(int x, int k)
    {
        if (k < 3) return 3 * ((x == 3) ? 0 : x + 2) + k;
        if (k == 3) return (x == 0) ? 3 : (x == 1) ? 5 : (x == 2) ? 16 : 17;
        if (k == 4) return (x == 0) ? 4 : (x == 1) ? 15 : (K::HAS_OH ? (x == 2 ? 18 : 19) : -1);
        return (x == 0) ? 20 : -1;
    }

    // q(c): position component c of this image (any callable); x: lane 0..3; mask: shuffle mask.
    // Returns the energy on lane 0 (0 on the others) and the gradient of the owned components.
    template <class QF>
    __device__ static __forceinline__ int eval_coop(QF q, int x, unsigned mask, double& V, double gown[NOWN],
                                                   double* scr = nullptr)
    {
        using namespace ch4h;
        auto shf = [&](double v, int src) { return __shfl_sync(mask, v, src & 3, 4); };
        const int ha = (x == 3) ? 0 : x + 2;
        double co[3], ubo[3], ucb[3], rcho, rbho, rcb, ircho;
        {
            double tc[3], tb[3], t[3];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                const double C = q(3 + d) * 0.52918, B = q(15 + d) * 0.52918, h = q(3 * ha + d) * 0.52918;
                t[d] = B - C;
                tc[d] = h - C;
                tb[d] = h - B;
            }
            double ircb, irbho;
            sqrt_rsqrt(dot(t, t), rcb, ircb);
            sqrt_rsqrt(dot(tc, tc), rcho, ircho);
            sqrt_rsqrt(dot(tb, tb), rbho, irbho);
#pragma unroll
            for (int d = 0; d < 3; d++) {
                ucb[d] = t[d] * ircb;
                co[d] = tc[d] * ircho;
                ubo[d] = tb[d] * irbho;
            }
        }
        // own switching functions
        double sw[10];  // s1,ds1,s2,ds2,s3,ds3,sphi,dsphi,sth,dsth
        {
            const double r = rcho, dr = r - K::R0CH;
            double omt, ms2;
            {
                const double u = r - K::B1S, u2 = u * u, u4 = u2 * u2, u7 = u4 * u2 * u, u8 = u4 * u4;
                const double arg = K::A1S * dr * u8;
                one_minus_tanh(arg, omt, ms2);
                const bool on = arg < 19.0;
                sw[0] = on ? omt : 0.0;
                sw[1] = on ? K::A1S * (u8 + 8.0 * dr * u7) * ms2 : 0.0;
            }
            {
                const double u = r - K::B2S, u2 = u * u, u4 = u2 * u2, u5 = u4 * u, u6 = u4 * u2;
                const double arg = K::A2S * dr * u6;
                one_minus_tanh(arg, omt, ms2);
                const bool on = arg < 19.0;
                sw[2] = on ? omt : 0.0;
                sw[3] = on ? K::A2S * (u6 + 6.0 * dr * u5) * ms2 : 0.0;
            }
            {
                const double u = r - K::B3S;
                const double arg = K::A3S * dr * u * u;
                one_minus_tanh(arg, omt, ms2);
                const bool on = arg < 19.0;
                sw[4] = on ? omt : 0.0;
                sw[5] = on ? K::A3S * (3.0 * r * r - 2.0 * r * (K::R0CH + 2.0 * K::B3S) + K::B3S * (K::B3S + 2.0 * K::R0CH)) * ms2 : 0.0;
            }
            {
                const bool on = r < 3.8, onphi = K::SPHI_ANY_R || on;
                const double u = r - K::CPHI, ex = CRCL_EXP(K::BPHI * u * u * u);
                one_minus_tanh(K::APHI * dr * ex, omt, ms2);
                sw[6] = onphi ? omt : 0.0;
                sw[7] = onphi ? K::APHI * (1.0 + 3.0 * K::BPHI * dr * u * u) * ex * ms2 : 0.0;
                const double v = r - K::CTHETA, ev = CRCL_EXP(K::BTHETA * v * v * v);
                one_minus_tanh(K::ATHETA * dr * ev, omt, ms2);
                sw[8] = on ? omt : 0.0;
                sw[9] = on ? K::ATHETA * (1.0 + 3.0 * K::BTHETA * dr * v * v) * ev * ms2 : 0.0;
            }
        }
        // own f1 and derivatives (ipforce_ch4h)
        double f1o, df1co, df1ho;
        {
            const double dr = rcho - K::R0CH, dh = rbho - K::R0HH;
            const double e1 = CRCL_EXP(-K::AA1 * rbho * rbho);
            const double e2 = CRCL_EXP(-K::AA4 * dh * dh);
            const double a1 = 1.0 - e1, a2 = K::AA2 + K::AA3 * e2;
            const double E = CRCL_EXP(-a2 * dr * dr);
            f1o = a1 * E;
            df1co = -2.0 * dr * a1 * a2 * E;
            df1ho = 2.0 * K::AA1 * rbho * e1 * E + 2.0 * K::AA3 * K::AA4 * dh * e2 * dr * dr * a1 * E;
        }
        // rotated gathers: local t <-> hydrogen (x+t)&3
        double c[4][3], rch[4], irch[4], s1[4], ds1[4], s2[4], ds2[4], s3[4], ds3[4], sphi[4], dsphi[4], sth[4], dsth[4];
        double f1[3], df1c[3], df1h[3];
#ifndef CRCL_CBE_SHFL_GATHER
        // scr: this bead's [lane][18] block; own values out, a warp-level fence, the neighbours' in -- in TWO phases: what the
        // stretch and the out-of-plane term read now, the in-plane term's share (s1, s2, f1 and derivatives: 21 doubles)
        // only after the out-of-plane term, behind a second fence, so that they are not live through it
        const double2* nb[3];
        {
            double2* mine = reinterpret_cast<double2*>(scr + x * COOP_SCRATCH);
            __syncwarp();   // the previous evaluation's reads of this block are complete
            mine[0] = make_double2(co[0], co[1]);
            mine[1] = make_double2(co[2], rcho);
            mine[2] = make_double2(ircho, sw[4]);
            mine[3] = make_double2(sw[5], sw[6]);
            mine[4] = make_double2(sw[7], sw[8]);
            mine[5] = make_double2(sw[9], sw[0]);
            mine[6] = make_double2(sw[1], sw[2]);
            mine[7] = make_double2(sw[3], f1o);
            mine[8] = make_double2(df1co, df1ho);
            mine[9] = make_double2(ubo[0], ubo[1]);     // own values parked for the chain rule at the end
            mine[10] = make_double2(ubo[2], ucb[0]);
            mine[11] = make_double2(ucb[1], ucb[2]);
            __syncwarp();
            c[0][0] = co[0], c[0][1] = co[1], c[0][2] = co[2];
            rch[0] = rcho, irch[0] = ircho;
            s3[0] = sw[4], sphi[0] = sw[6], sth[0] = sw[8];
#pragma unroll
            for (int t = 1; t < 4; t++) {
                const double2* o = reinterpret_cast<const double2*>(scr + ((x + t) & 3) * COOP_SCRATCH);
                nb[t - 1] = o;
                const double2 v0 = o[0], v1 = o[1], v2 = o[2];
                c[t][0] = v0.x, c[t][1] = v0.y, c[t][2] = v1.x;
                rch[t] = v1.y, irch[t] = v2.x;
                s3[t] = v2.y;
                sphi[t] = reinterpret_cast<const double*>(o)[7], sth[t] = reinterpret_cast<const double*>(o)[9];
            }
        }
        auto gather_inplane = [&]() {
            __syncwarp();   // keeps the loads below from being scheduled above the out-of-plane term
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const double2* o = t ? nb[t - 1] : reinterpret_cast<const double2*>(scr + x * COOP_SCRATCH);
                s1[t] = reinterpret_cast<const double*>(o)[11], s2[t] = reinterpret_cast<const double*>(o)[13];
                if (t < 3) f1[t] = reinterpret_cast<const double*>(o)[15];
            }
        };
#else
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int src = x + t;
#pragma unroll
            for (int d = 0; d < 3; d++) c[t][d] = t ? shf(co[d], src) : co[d];
            rch[t] = t ? shf(rcho, src) : rcho;
            irch[t] = t ? shf(ircho, src) : ircho;
            s1[t] = t ? shf(sw[0], src) : sw[0];
            ds1[t] = t ? shf(sw[1], src) : sw[1];
            s2[t] = t ? shf(sw[2], src) : sw[2];
            ds2[t] = t ? shf(sw[3], src) : sw[3];
            s3[t] = t ? shf(sw[4], src) : sw[4];
            ds3[t] = t ? shf(sw[5], src) : sw[5];
            sphi[t] = t ? shf(sw[6], src) : sw[6];
            dsphi[t] = t ? shf(sw[7], src) : sw[7];
            sth[t] = t ? shf(sw[8], src) : sw[8];
            dsth[t] = t ? shf(sw[9], src) : sw[9];
            if (t < 3) {
                f1[t] = t ? shf(f1o, src) : f1o;
                df1c[t] = t ? shf(df1co, src) : df1co;
                df1h[t] = t ? shf(df1ho, src) : df1ho;
            }
        }
        auto gather_inplane = [&]() {};
#endif
        // The derivatives of the switching functions are used once or twice each, late in their terms: 26 doubles that would
        // be live (and spilled: ncu attributed 14 % of the kernel's stall samples to reloads in this function) from the
        // gather to the end.  With the exchange block they are read where they are used, straight from the owner's slot
        // (volatile: a load the compiler may neither hoist nor keep).
#ifndef CRCL_CBE_SHFL_GATHER
        const double* nbd[4] = {scr + x * COOP_SCRATCH, reinterpret_cast<const double*>(nb[0]),
                                reinterpret_cast<const double*>(nb[1]), reinterpret_cast<const double*>(nb[2])};
        auto LDV = [&](int t, int off) { return *reinterpret_cast<const volatile double*>(nbd[t] + off); };
        auto DS3 = [&](int t) { return LDV(t, 6); };
        auto DSPHI = [&](int t) { return LDV(t, 8); };
        auto DSTH = [&](int t) { return LDV(t, 10); };
        auto DS1 = [&](int t) { return LDV(t, 12); };
        auto DS2 = [&](int t) { return LDV(t, 14); };
        auto DF1C = [&](int t) { return LDV(t, 16); };
        auto DF1H = [&](int t) { return LDV(t, 17); };
#else
        auto DS3 = [&](int t) { return ds3[t]; };
        auto DSPHI = [&](int t) { return dsphi[t]; };
        auto DSTH = [&](int t) { return dsth[t]; };
        auto DS1 = [&](int t) { return ds1[t]; };
        auto DS2 = [&](int t) { return ds2[t]; };
        auto DF1C = [&](int t) { return df1c[t]; };
        auto DF1H = [&](int t) { return df1h[t]; };
#endif
        const double tau = TAU_TET;
        const double ta = tau - K::TAU_PLANAR * PI, tb = tau - 2.0 * PI / 3.0;
        auto theta0 = [&](int i, int j, int k, int l) {
            return tau + ta * (sphi[i] * sphi[j] - 1.0) + tb * (sth[k] * sth[l] - 1.0);
        };
        // partial accumulators in the local frame
        double Dch[4] = {0, 0, 0, 0}, Dbh[3] = {0, 0, 0}, Dcb = 0.0;
        double gH[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, gC[3] = {0, 0, 0};
        double en = 0.0;

        // ---- stretching, own triple (C-H_x, C-H_b, H_b-H_x) ----
        {
            const double rav = (rch[0] + rch[1] + rch[2] + rch[3]) / 4.0;
            const double arga = K::C1CH * (rav - K::R0CH);
            double omt, ms2;
            one_minus_tanh(arga, omt, ms2);
            const bool on = arga < 19.0;
            const double ach = on ? K::A1CH + K::B1CH * (2.0 - omt) * 0.5 : K::A1CH + K::B1CH;
            const double dach = on ? -K::B1CH * K::C1CH * 0.5 * ms2 * 0.25 : 0.0;
            const Leps cb = leps(K::D1CB, K::D3CB, K::ACB, rcb - K::R0CB);
            const double dr = rcho - K::R0CH;
            const Leps ch = leps(K::D1CH, K::D3CH, ach, dr);
            const Leps bh = leps(K::D1HH, K::D3HH, K::AHH, rbho - K::R0HH);
            const double a = ch.vj, b = cb.vj, cc = bh.vj;
            double vjs, ivjs;
            sqrt_rsqrt((sqr(a - b) + sqr(b - cc) + sqr(cc - a)) * 0.5, vjs, ivjs);
            const double vj = -vjs;
            en += ch.vq + cb.vq + bh.vq + vj;
            const double h = -0.5 * ivjs;
            const double wa = (2.0 * a - b - cc) * h, wb = (2.0 * b - a - cc) * h, wc = (2.0 * cc - a - b) * h;
            const double dch = ch.dvq + wa * ch.dvj;
            Dbh[0] += bh.dvq + wc * bh.dvj;
            Dcb += cb.dvq + wb * cb.dvj;
            const double tt = dch * CRCL_DIV(dr, ach) * dach;
            Dch[0] += dch + tt;
            Dch[1] += tt;
            Dch[2] += tt;
            Dch[3] += tt;
        }
        // ---- out-of-plane bending, centre = own hydrogen (local 0), (j,k,l) = local (1,2,3) ----
        {
            const double pj = s3[1], pk = s3[2], pl = s3[3];
            const double swi = (1.0 - s3[0]) * pj * pk * pl;
            const double fd = swi * K::FCH3, hd = swi * K::HCH3;
            double a[3], b[3], n[3];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                const double hj = c[1][d] * rch[1];
                a[d] = c[2][d] * rch[2] - hj;
                b[d] = c[3][d] * rch[3] - hj;
            }
            cross(a, b, n);
            const double inn = CRCL_RSQRT(dot(n, n));
            double u[3] = {dot(n, c[1]) * inn, dot(n, c[2]) * inn, dot(n, c[3]) * inn};
            const int npos = (u[0] > 0.0) + (u[1] > 0.0) + (u[2] > 0.0);
            const double sg = (npos & 1) ? -1.0 : 1.0;
            const double nh[3] = {sg * n[0] * inn, sg * n[1] * inn, sg * n[2] * inn};
            double sum2 = 0.0, sum4 = 0.0, G[3] = {0, 0, 0};
#pragma unroll
            for (int t = 0; t < 3; t++) {
                const int m = t + 1;
                const int ca = (t == 0) ? 2 : 1, cb2 = (t == 2) ? 2 : 3;
                const double um = sg * u[t];
                const double del = CRCL_ACOS(um) - theta0(0, m, ca, cb2);
                const double d2 = del * del;
                sum2 += d2;
                sum4 += d2 * d2;
                const double w = 2.0 * fd * del + 4.0 * hd * d2 * del;
                const double wu = -w * CRCL_RSQRT(1.0 - um * um);
                const double f = wu * irch[m];
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    const double v = f * (nh[d] - um * c[m][d]);
                    gH[m][d] += v;
                    gC[d] -= v;
                    G[d] += wu * (c[m][d] - um * nh[d]);
                }
                Dch[0] -= w * ta * DSPHI(0) * sphi[m];
                Dch[m] -= w * ta * sphi[0] * DSPHI(m);
                Dch[ca] -= w * tb * DSTH(ca) * sth[cb2];
                Dch[cb2] -= w * tb * sth[ca] * DSTH(cb2);
            }
            {
                const double s = sg * inn;
                double gk[3], gl[3];
                cross(b, G, gk);
                cross(G, a, gl);
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    gH[2][d] += s * gk[d];
                    gH[3][d] += s * gl[d];
                    gH[1][d] -= s * (gk[d] + gl[d]);
                }
            }
            en += fd * sum2 + hd * sum4;
            const double fs = K::FCH3 * sum2 + K::HCH3 * sum4;
            Dch[0] -= fs * DS3(0) * pj * pk * pl;
            Dch[1] += fs * (1.0 - s3[0]) * DS3(1) * pk * pl;
            Dch[2] += fs * (1.0 - s3[0]) * pj * DS3(2) * pl;
            Dch[3] += fs * (1.0 - s3[0]) * pj * pk * DS3(3);
        }
        // ---- in-plane bending: pair local (0,1) on every lane, local (0,2) on lanes 0 and 1 ----
        gather_inplane();
        {
            constexpr double f0 = K::FKINF + K::AK, f2 = K::FKINF;
#pragma unroll
            for (int pp = 0; pp < 2; pp++) {
                const int j = pp + 1, k = pp ? 1 : 2, l = 3;
                if (pp == 1 && x >= 2) break;
                const double fk0 = f0 + f0 * (s1[0] * s1[j] - 1.0) + (f0 - f2) * (s2[k] * s2[l] - 1.0);
                const double ff = f1[0] * f1[j];
                const double Kf = fk0 * ff;
                const double cs = dot(c[0], c[j]);
                const double del = CRCL_ACOS(cs) - theta0(0, j, k, l);
                en += 0.5 * Kf * del * del;
                const double w = Kf * del;
                const double wc = -w * CRCL_RSQRT(1.0 - cs * cs);
                const double fi = wc * irch[0], fj = wc * irch[j];
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    const double vi = fi * (c[j][d] - cs * c[0][d]);
                    const double vj = fj * (c[0][d] - cs * c[j][d]);
                    gH[0][d] += vi;
                    gH[j][d] += vj;
                    gC[d] -= vi + vj;
                }
                const double hd2 = 0.5 * del * del;
                Dch[0] += -w * ta * DSPHI(0) * sphi[j] + hd2 * (f0 * DS1(0) * s1[j] * ff + fk0 * DF1C(0) * f1[j]);
                Dch[j] += -w * ta * sphi[0] * DSPHI(j) + hd2 * (f0 * s1[0] * DS1(j) * ff + fk0 * f1[0] * DF1C(j));
                Dch[k] += -w * tb * DSTH(k) * sth[l] + hd2 * (f0 - f2) * DS2(k) * s2[l] * ff;
                Dch[l] += -w * tb * sth[k] * DSTH(l) + hd2 * (f0 - f2) * s2[k] * DS2(l) * ff;
                Dbh[0] += hd2 * fk0 * DF1H(0) * f1[j];
                Dbh[j] += hd2 * fk0 * f1[0] * DF1H(j);
            }
        }
        double gO[3] = {0, 0, 0}, gBx[3] = {0, 0, 0};   // explicit vector parts on H(O) and on the abstracting atom
        if constexpr (K::HAS_OH) {
            double tno[3];
#pragma unroll
            for (int d = 0; d < 3; d++) tno[d] = q(18 + d) * 0.52918 - q(15 + d) * 0.52918;
            double rno, irno;
            sqrt_rsqrt(dot(tno, tno), rno, irno);
            // O-H(O) Morse bond (stretch_ch4oh :616-621, :652-659, :769-774): lane 0 only
            const double m0 = (x == 0) ? 1.0 : 0.0;
            const double ex = CRCL_EXP(-K::MORSE_A * (rno - K::MORSE_R0)), om = 1.0 - ex;
            en += m0 * K::MORSE_D * (om * om);
            const double de = m0 * 2.0 * K::MORSE_A * K::MORSE_D * om * ex * irno;
#pragma unroll
            for (int d = 0; d < 3; d++) {
                gBx[d] -= de * tno[d];
                gO[d] += de * tno[d];
            }
            // H_x-O-H(O) bend of the own hydrogen (ipbend_ch4oh :1059-1171)
            double tbv[3];
#pragma unroll
            for (int d = 0; d < 3; d++) tbv[d] = -ubo[d] * rbho;
            double cs = -dot(tno, tbv) * CRCL_RCP(rno * rbho);
            cs = fmin(1.0, fmax(-1.0, cs));
            const double dang = CRCL_ACOS(cs) - K::ANH2OEQ;
            const double arga = K::ALPH2O * (rbho - K::R0HH);
            double omt, ms2;
            one_minus_tanh(arga, omt, ms2);
            const double fk = (arga < 19.0) ? K::FKH2OEQ * omt : 0.0;
            en += 0.5 * fk * dang * dang;
            Dbh[0] += K::FKH2OEQ * K::ALPH2O * ms2 * (0.5 * dang * dang);
            const double dstda = fk * dang;
            double pv[3], v1[3], v2[3];
            cross(tno, tbv, pv);
            double rp = CRCL_SQRT(fmax(dot(pv, pv), 1.0e-300));
            if (rp < 1.0e-6) rp = 1.0e-6;
            const double terma = dstda * CRCL_RCP(rbho * rbho * rp), termc = dstda * CRCL_RCP(rno * rno * rp);
            cross(tbv, pv, v1);
            cross(tno, pv, v2);
#pragma unroll
            for (int d = 0; d < 3; d++) {
                const double hi = -terma * v1[d], ho = -termc * v2[d];
                gH[0][d] += hi;
                gO[d] += ho;
                gBx[d] += -hi - ho;
            }
        }
        // ---- reduce-scatter to the owner lane: hydrogen y collects local t from lane y-t ----
        double DchT = Dch[0], DbhT = Dbh[0], gHT[3] = {gH[0][0], gH[0][1], gH[0][2]};
#pragma unroll
        for (int t = 1; t < 4; t++) {
            DchT += shf(Dch[t], x - t + 4);
            if (t < 3) DbhT += shf(Dbh[t], x - t + 4);
#pragma unroll
            for (int d = 0; d < 3; d++) gHT[d] += shf(gH[t][d], x - t + 4);
        }
        // own hydrogen's chain rule; its share of the C and H_b gradients; all-reduce those
        double gB[3];
#ifndef CRCL_CBE_SHFL_GATHER
        double cov[3], ubv[3], ucv[3];
        {
            const double2* m = reinterpret_cast<const double2*>(scr + x * COOP_SCRATCH);
            const double2 m0 = m[0], m9 = m[9], m10 = m[10], m11 = m[11];
            cov[0] = m0.x, cov[1] = m0.y, cov[2] = reinterpret_cast<const double*>(m)[2];
            ubv[0] = m9.x, ubv[1] = m9.y, ubv[2] = m10.x;
            ucv[0] = m10.y, ucv[1] = m11.x, ucv[2] = m11.y;
        }
#else
        const double(&cov)[3] = co, (&ubv)[3] = ubo, (&ucv)[3] = ucb;
#endif
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const double vc = DchT * cov[d], vb = DbhT * ubv[d], vcb = Dcb * ucv[d];
            gHT[d] += vc + vb;
            gC[d] -= vc + vcb;
            gB[d] = vcb - vb;
            if constexpr (K::HAS_OH) gB[d] += gBx[d];   // (x + 0.0 is not a no-op in IEEE arithmetic: keep it out of the 6-atom code)
        }
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
#pragma unroll
            for (int d = 0; d < 3; d++) {
                gC[d] += __shfl_xor_sync(mask, gC[d], o, 4);
                gB[d] += __shfl_xor_sync(mask, gB[d], o, 4);
                if constexpr (K::HAS_OH) gO[d] += __shfl_xor_sync(mask, gO[d], o, 4);
            }
            en += __shfl_xor_sync(mask, en, o, 4);
        }
        constexpr double GF = 0.0201723;
        V = (x == 0) ? en * 0.03812 : 0.0;
        gown[0] = gHT[0] * GF;
        gown[1] = gHT[1] * GF;
        gown[2] = gHT[2] * GF;
        gown[3] = ((x == 0) ? gC[0] : (x == 1) ? gC[2] : (x == 2) ? gB[1] : gB[2]) * GF;
        gown[4] = ((x == 0) ? gC[1] : (x == 1) ? gB[0] : K::HAS_OH ? ((x == 2) ? gO[0] : gO[1]) : 0.0) * GF;
        if constexpr (K::HAS_OH) gown[5] = (x == 0) ? gO[2] * GF : 0.0;
        return 0;
    }
};

using PesCH4H4 = PesCBE4<ch4h::K6>;

}  // namespace crcl
#endif
