// pes_brh2.cuh -- BrH2 DIM-3C surface (Clary 1982 / Last-Baer 1981), one thread per image, FP64.
//
// Replaces /root/reference/src/egrad_brh2.f (egrad_brh2 :33-80, initialize_brh2 :84-186, POT_brh2
// :188-440, VHX :442-463, EISPACK RSP/TRED3/TQL2/TRBAK3 :485-1059, constants :465-483, all D0).
// Atom order H, Br, H: R1 = r(H1-Br), R2 = r(Br-H3), R3 = r(H1-H3).
//
// Restructured for the GPU, not a transcription:
//  * the EISPACK chain (Householder + QL with its 2**-37 threshold, which the results depend on at the 1e-10
//    level) is specialised to the 4x4 case and the lowest root, fully unrolled, in registers (lowest_root);
//  * the reference assembles three packed derivative matrices dH/dR_i (30 entries) and contracts each with
//    u u^T (:323-402).  Here the Hellmann-Feynman sum is taken once with respect to the 16 quantities the
//    Hamiltonian is built from -- the two HBr base curves per side (all four HBr states are the singlet
//    sigma curve and multiples of one anti-Morse curve, VHX :450-461), the two H2 curves and the four
//    angular functions cos^2 A, sin 2A, cos^2 B, sin 2B -- and chained to R1, R2, R3 afterwards;
//  * one exponential per Morse leg, the overlap polynomials of the three-centre term share their exps.
// Reference behaviour kept: the triangle-inequality clamps (:201-203), |cos| clamps (:214,236), the
// collinear switch LCOL = sin^2 A < 1e-6 that zeroes the sin 2A / sin 2B couplings and their angular
// derivatives but not their radial ones (:218-225,308-320,325), G divided by RHX once (:156).
#pragma once
#include "crcl_common.cuh"

namespace crcl {
namespace brh2 {

constexpr double EPS = 1.e-6;
constexpr double RHH = 1.4016, AHH = 1.0291, DHH = 0.17447, BHH = 0.018;
constexpr double XIH = 1.0, XIX = 1.6, G0 = 0.22, ALFW = 1.0;
constexpr double RHX = 2.673, AHX = 0.957, DHX = 0.1439, BHX = 0.012;
constexpr double ETAHH = 0.393764, ETA3S = 0.322, ETA1P = 0.20286, ETA3P = 0.1771;
constexpr double XIB = 0.5 * (XIH + XIX);
constexpr double C3 = 1.0 / 3.0;
constexpr double RT34 = 0.4330127018922193;   // 0.25 sqrt(3)
constexpr double RT38 = 0.5 * RT34;
constexpr double GS = G0 / RHX;

// the two base curves of one H-Br leg: singlet sigma vs and the anti-Morse va, with derivatives
CRCL_HD __forceinline__ void hx(double R, double& vs, double& va, double& dvs, double& dva)
{
    const double d = R - RHX, d2 = d * d;
    const double e1 = CRCL_EXP(-AHX * d), e2 = CRCL_EXP(-BHX * d2 * d);
    const double t = DHX * e1 * e2, t1 = 2.0 * AHX * t, t2 = 3.0 * BHX * d2;
    vs = t * (e1 - 2.0);
    va = t * (e1 + 2.0);
    dvs = t1 * (1.0 - e1) - t2 * vs;
    dva = -t1 * (1.0 + e1) - t2 * va;
}

// ---- lowest root of the packed symmetric 4x4 Hamiltonian ------------------------------------------------------
// The reference diagonalises with EISPACK's RSP (egrad_brh2.f:485-1059): Householder tridiagonalisation of the
// packed matrix (TRED3), QL with implicit shifts that stops at an off-diagonal threshold of 2**-37 (TQL2, :810),
// back-transformation (TRBAK3).  That threshold leaves the eigenvector -- hence the Hellmann-Feynman gradient --
// with a truncation error of up to a few 1e-10 relative; a diagonalisation converged to rounding (a Jacobi sweep
// was measured) differs from the reference by 2.7e-10 on 1 of 1e5 images, above the 1e-10 parity bar.  So the same
// three stages are done here, for the fixed size 4 and for the lowest root only, entirely in registers: every
// loop bound of TRED3 / TRBAK3 is a compile-time constant after unrolling, and the data-dependent sweep range
// l..m-1 of TQL2 becomes a predicated unrolled loop.  a[] is the packed lower triangle, index r(r+1)/2 + c.
CRCL_HD __forceinline__ double sgn(double a, double b) { return (b >= 0.0) ? fabs(a) : -fabs(a); }
CRCL_HD __forceinline__ constexpr int pk(int r, int c) { return r >= c ? r * (r + 1) / 2 + c : c * (c + 1) / 2 + r; }

// returns the lowest eigenvalue and its eigenvector u; 1 in *fail if TQL2 exceeds its 30 iterations (the
// reference STOPs, :379-385)
CRCL_HD __forceinline__ double lowest_root(double (&a)[10], double (&u)[4], int& fail)
{
    double d[4], e[4];
    // TRED3 (:560-671): rows 3, 2, 1, 0
#pragma unroll
    for (int i = 3; i >= 0; i--) {
        const int l = i;   // number of off-diagonal entries of row i
        double h = 0.0, scale = 0.0;
#pragma unroll
        for (int k = 0; k < l; k++) {
            d[k] = a[pk(i, k)];
            scale += fabs(d[k]);
        }
        if (l < 1 || scale == 0.0) {
            e[i] = 0.0;
        } else {
#pragma unroll
            for (int k = 0; k < l; k++) {
                d[k] = d[k] / scale;
                h += d[k] * d[k];
            }
            double f = d[l - 1];
            double g = -sgn(CRCL_SQRT(h), f);   // h in [1/l, 1] after the scaling
            e[i] = scale * g;
            h = h - f * g;
            d[l - 1] = f - g;
            a[pk(i, l - 1)] = scale * d[l - 1];
            if (l != 1) {
                f = 0.0;
#pragma unroll
                for (int j = 0; j < l; j++) {
                    g = 0.0;
#pragma unroll
                    for (int k = 0; k < l; k++) g += a[pk(j, k)] * d[k];
                    e[j] = g / h;
                    f += e[j] * d[j];
                }
                const double hh = f / (h + h);
#pragma unroll
                for (int j = 0; j < l; j++) {
                    f = d[j];
                    g = e[j] - hh * f;
                    e[j] = g;
#pragma unroll
                    for (int k = 0; k <= j; k++) a[pk(j, k)] = a[pk(j, k)] - f * e[k] - g * d[k];
                }
            }
        }
        d[i] = a[pk(i, i)];
        a[pk(i, i)] = scale * CRCL_SQRT0(h);
    }
    // TQL2 (:808-978) on the tridiagonal (d, e); z accumulates the rotations
    double z[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) z[r][c] = (r == c) ? 1.0 : 0.0;
    constexpr double MACHEP = 7.275957614183426e-12;   // 2**-37
    e[0] = e[1];
    e[1] = e[2];
    e[2] = e[3];
    e[3] = 0.0;
    double f = 0.0, b = 0.0;
    fail = 0;
#pragma unroll
    for (int l = 0; l < 3; l++) {   // l = 3 never iterates (e[3] = 0): only its d[3] += f remains, below
        const double h0 = MACHEP * (fabs(d[l]) + fabs(e[l]));
        if (b < h0) b = h0;
        int m = 3;
#pragma unroll
        for (int t = 2; t >= l; t--)
            if (fabs(e[t]) <= b) m = t;   // first m >= l with |e[m]| <= b (e[3] = 0)
        if (m != l) {
            int iter = 0;
            do {
                if (iter == 30) {
                    fail = 1;
                    break;
                }
                iter++;
                const double g0 = d[l];
                double p = (d[l + 1] - g0) / (2.0 * e[l]);
                double r = CRCL_SQRT(p * p + 1.0);
                d[l] = e[l] / (p + sgn(r, p));
                const double h = g0 - d[l];
#pragma unroll
                for (int i = l + 1; i < 4; i++) d[i] = d[i] - h;
                f = f + h;
                p = (m == 3) ? d[3] : ((m == 2) ? d[2] : d[1]);   // d[m], m > l >= 0
                double c = 1.0, s = 0.0;
#pragma unroll
                for (int i = 2; i >= l; i--)
                    if (i < m) {
                        const double g = c * e[i], hh = c * p;
                        if (!(fabs(p) < fabs(e[i]))) {
                            c = e[i] / p;
                            r = CRCL_SQRT(c * c + 1.0);
                            e[i + 1] = s * p * r;
                            s = c / r;
                            c = 1.0 / r;
                        } else {
                            c = p / e[i];
                            r = CRCL_SQRT(c * c + 1.0);
                            e[i + 1] = s * e[i] * r;
                            s = 1.0 / r;
                            c = c * s;
                        }
                        p = c * d[i] - s * g;
                        d[i + 1] = hh + s * (c * g + s * d[i]);
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const double zz = z[k][i + 1];
                            z[k][i + 1] = s * z[k][i] + c * zz;
                            z[k][i] = c * z[k][i] - s * zz;
                        }
                    }
                e[l] = s * p;
                d[l] = c * p;
            } while (fabs(e[l]) > b);
        }
        d[l] = d[l] + f;
    }
    d[3] = d[3] + f;
    // the ordering pass (:953-971) moves the first strict minimum to column 1: take that column directly
    double e0 = d[0];
#pragma unroll
    for (int r = 0; r < 4; r++) u[r] = z[r][0];
#pragma unroll
    for (int k = 1; k < 4; k++) {
        const bool lo = d[k] < e0;
        e0 = lo ? d[k] : e0;
#pragma unroll
        for (int r = 0; r < 4; r++) u[r] = lo ? z[r][k] : u[r];
    }
    // TRBAK3 (:980-1059) on that one column
#pragma unroll
    for (int i = 1; i < 4; i++) {
        const double h = a[pk(i, i)];
        if (h != 0.0) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < i; k++) s += a[pk(i, k)] * u[k];
            s = (s / h) / h;
#pragma unroll
            for (int k = 0; k < i; k++) u[k] = u[k] - s * a[pk(i, k)];
        }
    }
    return e0;
}

// energy and dE/d(R1,R2,R3) in the potential's own ordering (R1, R2 the two H-Br legs, R3 = H-H)
CRCL_HD __forceinline__ void pot(double R1, double R2, double R3, double& V, double& d1, double& d2, double& d3, int& fail)
{
    if (R1 > R2 + R3) R1 = R2 + R3;
    if (R2 > R1 + R3) R2 = R1 + R3;
    if (R3 > R1 + R2) R3 = R1 + R2;
    const double R1S = R1 * R1, R2S = R2 * R2, R3S = R3 * R3;
    const double i12 = CRCL_RCP(R1 * R2), i13 = CRCL_RCP(R1 * R3), i23 = CRCL_RCP(R2 * R3);
    const double T3 = 0.5 * (R1S + R2S - R3S), T2 = 0.5 * (R1S + R3S - R2S), T1 = 0.5 * (R2S + R3S - R1S);
    // angle A at H1 (between R1 and R3), angle B at H3 (between R2 and R3); G is the angle at Br
    double csa = T2 * i13;
    if (fabs(csa) > 1.0) csa = copysign(1.0, csa);
    const double ca2 = csa * csa, sa2 = 1.0 - ca2;
    const bool lcol = sa2 < EPS;
    const double sna = CRCL_SQRT0(sa2), s2a = 2.0 * csa * sna;
    const double ka = lcol ? 0.0 : 2.0 * (sa2 - ca2) / sna;       // d sin2A / d cosA
    double csb = T1 * i23;
    if (fabs(csb) > 1.0) csb = copysign(1.0, csb);
    const double cb2 = csb * csb, sb2 = 1.0 - cb2;
    const double snb = CRCL_SQRT0(sb2), s2b = 2.0 * csb * snb;
    const double kb = lcol ? 0.0 : 2.0 * (sb2 - cb2) / snb;
    // d cosA / dR_i and d cosB / dR_i
    const double a1 = T3 * i13 / R1, a2 = -R2 * i13, a3 = T1 * i13 / R3;
    const double b1 = -R1 * i23, b2 = T3 * i23 / R2, b3 = T2 * i23 / R3;
    const double csg = T3 * i12, csg2 = csg * csg;
    const double g1 = 2.0 * csg * T2 * i12 / R1, g2 = 2.0 * csg * T1 * i12 / R2, g3 = -2.0 * csg * R3 * i12;

    // diatomic curves
    double v1hh, v3hh, dv1hh, dv3hh;
    {
        const double d = R3 - RHH, dd = d * d;
        const double e1 = CRCL_EXP(-AHH * d), e2 = CRCL_EXP(-BHH * d * dd);
        const double t = DHH * e1 * e2, t1 = 2.0 * AHH * t, t2 = 3.0 * BHH * dd;
        v1hh = t * (e1 - 2.0);
        v3hh = ETAHH * t * (e1 + 2.0);
        dv1hh = t1 * (1.0 - e1) - t2 * v1hh;
        dv3hh = -t1 * ETAHH * (1.0 + e1) - t2 * v3hh;
    }
    double vs1, va1, dvs1, dva1, vs2, va2, dvs2, dva2;
    hx(R1, vs1, va1, dvs1, dva1);
    hx(R2, vs2, va2, dvs2, dva2);
    // sigma / pi combinations (S1x = 1S + 3 3S, S2x = 3 1S + 3S, S3x = 1S - 3S, likewise P) per leg
    constexpr double PS1 = ETA1P + 3.0 * ETA3P, PS2 = 3.0 * ETA1P + ETA3P, PS3 = ETA1P - ETA3P;
    const double S11 = vs1 + 3.0 * ETA3S * va1, S21 = 3.0 * vs1 + ETA3S * va1, S31 = vs1 - ETA3S * va1;
    const double S12 = vs2 + 3.0 * ETA3S * va2, S22 = 3.0 * vs2 + ETA3S * va2, S32 = vs2 - ETA3S * va2;
    const double T11 = S11 - PS1 * va1, T21 = S21 - PS2 * va1, T31 = S31 - PS3 * va1;   // S - P
    const double T12 = S12 - PS1 * va2, T22 = S22 - PS2 * va2, T32 = S32 - PS3 * va2;
    const double P11 = PS1 * va1, P21 = PS2 * va1, P31 = PS3 * va1;
    const double P12 = PS1 * va2, P22 = PS2 * va2, P32 = PS3 * va2;

    double h[10];   // packed lower triangle: (0,0) (1,0) (1,1) (2,0) (2,1) (2,2) (3,0) (3,1) (3,2) (3,3)
    h[0] = v1hh + 0.25 * (S11 * ca2 + P11 * sa2 + S12 * cb2 + P12 * sb2);
    h[2] = v1hh + 0.25 * (S11 * sa2 + P11 * ca2 + S12 * sb2 + P12 * cb2);
    h[5] = v3hh + 0.25 * (S21 * ca2 + P21 * sa2 + S22 * cb2 + P22 * sb2);
    h[9] = v3hh + 0.25 * (S21 * sa2 + P21 * ca2 + S22 * sb2 + P22 * cb2);
    h[1] = lcol ? 0.0 : 0.125 * (T11 * s2a - T12 * s2b);
    h[3] = RT34 * (S31 * ca2 + P31 * sa2 - S32 * cb2 - P32 * sb2);
    h[4] = lcol ? 0.0 : RT38 * (T31 * s2a + T32 * s2b);
    h[6] = h[4];
    h[7] = RT34 * (S31 * sa2 + P31 * ca2 - S32 * sb2 - P32 * cb2);
    h[8] = lcol ? 0.0 : 0.125 * (T21 * s2a - T22 * s2b);
    double u[4];
    const double e0 = lowest_root(h, u, fail);

    // Hellmann-Feynman weights: diagonal w_ii, off-diagonal counted twice
    const double w0 = u[0] * u[0], w1 = u[1] * u[1], w2 = u[2] * u[2], w3 = u[3] * u[3];
    const double x10 = 2.0 * u[1] * u[0], x20 = 2.0 * u[2] * u[0], x31 = 2.0 * u[3] * u[1], x32 = 2.0 * u[3] * u[2];
    const double xm = 2.0 * (u[2] * u[1] + u[3] * u[0]);   // H(5) and H(7) are the same element
    // dE / d(S, P) per leg; the sin 2A coupling enters the radial derivative even when LCOL
    const double eS11 = 0.25 * (w0 * ca2 + w1 * sa2) + 0.125 * x10 * s2a, eP11 = 0.25 * (w0 * sa2 + w1 * ca2) - 0.125 * x10 * s2a;
    const double eS21 = 0.25 * (w2 * ca2 + w3 * sa2) + 0.125 * x32 * s2a, eP21 = 0.25 * (w2 * sa2 + w3 * ca2) - 0.125 * x32 * s2a;
    const double eS31 = RT34 * (x20 * ca2 + x31 * sa2) + RT38 * xm * s2a, eP31 = RT34 * (x20 * sa2 + x31 * ca2) - RT38 * xm * s2a;
    const double eS12 = 0.25 * (w0 * cb2 + w1 * sb2) - 0.125 * x10 * s2b, eP12 = 0.25 * (w0 * sb2 + w1 * cb2) + 0.125 * x10 * s2b;
    const double eS22 = 0.25 * (w2 * cb2 + w3 * sb2) - 0.125 * x32 * s2b, eP22 = 0.25 * (w2 * sb2 + w3 * cb2) + 0.125 * x32 * s2b;
    const double eS32 = -RT34 * (x20 * cb2 + x31 * sb2) + RT38 * xm * s2b, eP32 = -RT34 * (x20 * sb2 + x31 * cb2) - RT38 * xm * s2b;
    // ... folded onto the two base curves of each leg
    const double evs1 = eS11 + 3.0 * eS21 + eS31;
    const double eva1 = ETA3S * (3.0 * eS11 + eS21 - eS31) + PS1 * eP11 + PS2 * eP21 + PS3 * eP31;
    const double evs2 = eS12 + 3.0 * eS22 + eS32;
    const double eva2 = ETA3S * (3.0 * eS12 + eS22 - eS32) + PS1 * eP12 + PS2 * eP22 + PS3 * eP32;
    // dE / d(cos^2 A), d(cos^2 B), d(sin 2A), d(sin 2B)
    const double eca2 = 0.25 * (T11 * (w0 - w1) + T21 * (w2 - w3)) + RT34 * T31 * (x20 - x31);
    const double ecb2 = 0.25 * (T12 * (w0 - w1) + T22 * (w2 - w3)) - RT34 * T32 * (x20 - x31);
    const double es2a = 0.125 * (T11 * x10 + T21 * x32) + RT38 * T31 * xm;
    const double es2b = -0.125 * (T12 * x10 + T22 * x32) + RT38 * T32 * xm;
    const double fa = 2.0 * csa * eca2 + ka * es2a, fb = 2.0 * csb * ecb2 + kb * es2b;   // dE/dcosA, dE/dcosB
    const double D1 = evs1 * dvs1 + eva1 * dva1 + fa * a1 + fb * b1;
    const double D2 = evs2 * dvs2 + eva2 * dva2 + fa * a2 + fb * b2;
    const double D3 = (w0 + w1) * dv1hh + (w2 + w3) * dv3hh + fa * a3 + fb * b3;

    // three-centre term (:406-427)
    double shh, dshh, shx1, dshx1, shx2, dshx2;
    {
        const double t = XIH * R3, ex = CRCL_EXP(-t);
        shh = (1.0 + t * (1.0 + C3 * t)) * ex;
        dshh = -XIH * t * (1.0 + t) * C3 * ex;
    }
    constexpr double TX = XIB * RHX;
    const double iden = 1.0 / (2.0 * RHX * (1.0 + TX * (1.0 + C3 * TX)) * CRCL_EXP(-TX));
    {
        const double t = XIB * R1, ex = CRCL_EXP(-t) * iden;
        shx1 = R1 * (1.0 + t * (1.0 + C3 * t)) * ex;
        dshx1 = (1.0 + t * (1.0 - C3 * t * t)) * ex;
    }
    {
        const double t = XIB * R2, ex = CRCL_EXP(-t) * iden;
        shx2 = R2 * (1.0 + t * (1.0 + C3 * t)) * ex;
        dshx2 = (1.0 + t * (1.0 - C3 * t * t)) * ex;
    }
    const double rd = R1 - R2;
    const double tg = GS * CRCL_EXP(-ALFW * rd * rd);
    const double ov = shh * (shx1 + shx2) + shx1 * shx2;
    const double tw = 2.0 * ALFW * rd * csg2;
    V = e0 + ov * tg * csg2 + DHH;
    d1 = D1 + (dshx1 * (shh + shx2) * csg2 + ov * (g1 - tw)) * tg;
    d2 = D2 + (dshx2 * (shh + shx1) * csg2 + ov * (g2 + tw)) * tg;
    d3 = D3 + (dshh * (shx1 + shx2) * csg2 + ov * g3) * tg;
}

}  // namespace brh2

struct PesBrH2 {
    static constexpr int NATOMS = 3;
    static constexpr int ID = CRCL_PES_BRH2;
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
    CRCL_HD static __forceinline__ int eval(const double* __restrict__ q, double& V, double* __restrict__ g)
    {
        // legs: H1-Br (atoms 0,1), Br-H3 (atoms 1,2), H1-H3 (atoms 0,2)
        double v1[3], v2[3], v3[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            v1[d] = q[3 + d] - q[d];
            v2[d] = q[6 + d] - q[3 + d];
            v3[d] = q[6 + d] - q[d];
        }
        double R1, R2, R3, iR1, iR2, iR3;
        sqrt_rsqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2], R1, iR1);
        sqrt_rsqrt(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2], R2, iR2);
        sqrt_rsqrt(v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2], R3, iR3);
        double d1, d2, d3;
        int fail;
        brh2::pot(R1, R2, R3, V, d1, d2, d3, fail);
        const double f1 = d1 * iR1, f2 = d2 * iR2, f3 = d3 * iR3;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            g[d] = -f1 * v1[d] - f3 * v3[d];
            g[3 + d] = f1 * v1[d] - f2 * v2[d];
            g[6 + d] = f2 * v2[d] + f3 * v3[d];
        }
        return fail;
    }
};

}  // namespace crcl
