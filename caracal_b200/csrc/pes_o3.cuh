// pes_o3.cuh -- O3 1 1A" surface (Varga, Paukku, Truhlar 2017: permutationally invariant polynomials in mixed
// exponential-Gaussian variables + fitted two-body term + D3(BJ) dispersion), one thread per image, FP64.
//
// Replaces /root/reference/src/egrad_o3.f (egrad_o3 :33-58, pot_o3 :83-118, o3pes :144-184, EvMorse / EvMono /
// EvPoly / evbas :323-500, ev2gm2 :503-565, the six derivative routines :567-931, d3disp / edisp :933-1090,
// coefficients :1107-1172).  SURVEY.md 8f row N4.
//
// Restructured for the GPU, not a transcription: the reference evaluates the 67 polynomials and then, in a second
// set of hand-expanded routines, their derivatives with respect to the three distances through dense 3 x 67 arrays
// in COMMON.  Here every polynomial is a value with its three partial derivatives (forward-mode: the product rule in
// the operand order the source uses, so the numbers agree to rounding), the recurrences are straight-line code with
// compile-time indices (registers, no arrays in memory), the 56 fitted coefficients are folded in as soon as a
// polynomial is final, and the chain rule to Cartesians is three pair vectors.  The two-body term uses the
// geometric progression of its eight exponents (one pow-free product per term).
#pragma once
#include "crcl_common.cuh"

namespace crcl {
namespace o3 {

struct J {   // value and d/dR1, d/dR2, d/dR3
    double v, a, b, c;
};
CRCL_HD __forceinline__ J mul(const J& x, const J& y)
{
    return {x.v * y.v, x.a * y.v + x.v * y.a, x.b * y.v + x.v * y.b, x.c * y.v + x.v * y.c};
}
CRCL_HD __forceinline__ J sub(const J& x, const J& y) { return {x.v - y.v, x.a - y.a, x.b - y.b, x.c - y.c}; }
CRCL_HD __forceinline__ J add3(const J& x, const J& y, const J& z)
{
    return {x.v + y.v + z.v, x.a + y.a + z.a, x.b + y.b + z.b, x.c + y.c + z.c};
}

// fitted two-body term, eight Gaussians with exponents alpha beta^k (ev2gm2_o3), kcal/mol and kcal/(mol A)
CRCL_HD __forceinline__ void v2(double r, double& v, double& dv)
{
    constexpr double alpha = 9.439784362354936e-1, beta = 1.262242998506810e0;
    constexpr double a[8] = {-1.488979427684798e3, 1.881435846488955e4,  -1.053475425838226e5, 2.755135591229064e5,
                             -4.277588997761775e5, 4.404104009614092e5, -2.946204062950765e5, 1.176861219078620e5};
    const double r2 = r * r;
    double bk = 1.0, sv = 0.0, sg = 0.0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const double ex = a[k] * CRCL_EXP(-alpha * bk * r2);
        sv += ex;
        sg -= 2.0 * alpha * bk * r * ex;
        bk *= beta;
    }
    v = sv * 627.509523475149e-3;
    dv = sg * 627.509523475149e-3;
}

// D3(BJ) dispersion of one O-O pair (d3disp_o3 / edisp_o3 with the fixed C6 = 12.8), kcal/mol and kcal/(mol A)
CRCL_HD __forceinline__ void disp(double R, double& e, double& de)
{
    constexpr double autoang = 0.52917726, autokcal = 627.509541, s8 = 2.0, a1 = 0.5299, a2 = 2.20, c6 = 12.8;
    const double r2r4 = (double)2.59361680f;   // r2r4(8): REAL*4 literal in the source
    const double c8 = 3.0 * c6 * r2r4 * r2r4;
    const double t = a1 * sqrt(c8 / c6) + a2, t2 = t * t, t6 = t2 * t2 * t2, t8 = t6 * t2;
    const double r = CRCL_DIV(R, autoang), q2 = r * r, q4 = q2 * q2, q6 = q4 * q2, q8 = q4 * q4;
    const double i6 = CRCL_RCP(q6 + t6), i8 = CRCL_RCP(q8 + t8);
    e = (-c6 * i6 - s8 * c8 * i8) * autokcal;
    de = (6.0 * c6 * q4 * r * i6 * i6 + 8.0 * s8 * c8 * q6 * r * i8 * i8) * CRCL_DIV(autokcal, autoang);
}

#define O3P0(k, x, y) const J p##k = mul(p##x, p##y);
#define O3P1(k, x, y, s) const J p##k = sub(mul(p##x, p##y), p##s);
#define O3P2(k, x, y, s, t) const J p##k = sub(sub(mul(p##x, p##y), p##s), p##t);
#define O3P3(k, x, y, s, t, u) const J p##k = sub(sub(sub(mul(p##x, p##y), p##s), p##t), p##u);
#define O3B(k, coef)                 \
    acc.v = fma(coef, p##k.v, acc.v); \
    acc.a = fma(coef, p##k.a, acc.a); \
    acc.b = fma(coef, p##k.b, acc.b); \
    acc.c = fma(coef, p##k.c, acc.c);

// V (kcal/mol, the fit's own zero) and dV/dR (kcal/(mol A)) from the three distances in Angstrom:
// R1 = r(O1 O2), R2 = r(O1 O3), R3 = r(O2 O3)
CRCL_HD __forceinline__ void pes(const double (&R)[3], double& V, double (&dV)[3])
{
    constexpr double A = 0.83, AB = 3.70, RA = 1.25, RB = 1.13;
    double m[3], dm[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double u = R[i] - RB;
        m[i] = CRCL_EXP(-CRCL_DIV(R[i] - RA, A) - CRCL_DIV(u * u, AB));
        dm[i] = (CRCL_DIV(-2.0 * u, AB) - 1.0 / A) * m[i];
    }
    // monomials (EvMono): rm1 = ms(3), rm2 = ms(2), rm3 = ms(1); pair and triple products
    const J q1 = {m[2], 0.0, 0.0, dm[2]}, q2 = {m[1], 0.0, dm[1], 0.0}, q3 = {m[0], dm[0], 0.0, 0.0};
    const J q4 = mul(q1, q2), q5 = mul(q1, q3), q6 = mul(q2, q3), q7 = mul(q1, q6);
    const J p1 = add3(q1, q2, q3), p2 = add3(q4, q5, q6), p4 = q7;
    J acc = {0.0, 0.0, 0.0, 0.0};
    O3P2(3, 1, 1, 2, 2)
    O3P3(5, 1, 2, 4, 4, 4)
    O3P1(6, 1, 3, 5)
    O3P0(7, 1, 4)
    O3P2(8, 2, 2, 7, 7)
    O3P1(9, 2, 3, 7)
    O3P1(10, 1, 6, 9)
    O3P0(11, 2, 4)
    O3P0(12, 3, 4)
    O3P1(13, 1, 8, 11)
    O3P1(14, 2, 6, 12)
    O3P1(15, 1, 10, 14)
    O3P0(16, 4, 4)
    O3P0(17, 4, 5)
    O3P0(18, 4, 6)
    O3P1(19, 2, 8, 17)
    O3P3(20, 1, 13, 17, 19, 19)
    O3P1(21, 2, 10, 18)
    O3P1(22, 1, 15, 21)
    O3P0(23, 1, 16)
    O3P0(24, 4, 8)
    O3P1(25, 3, 11, 23)
    O3P0(26, 4, 10)
    O3P1(27, 1, 19, 24)
    O3P1(28, 6, 8, 23)
    O3P1(29, 2, 15, 26)
    O3P1(30, 1, 22, 29)
    O3P0(31, 2, 16)
    O3P0(32, 3, 16)
    O3P1(33, 1, 24, 31)
    O3P0(34, 4, 14)
    O3P0(35, 4, 15)
    O3P1(36, 2, 19, 33)
    O3P1(37, 3, 19, 31)
    O3P1(38, 8, 10, 32)
    O3P1(39, 2, 22, 35)
    O3P1(40, 1, 30, 39)
    O3P0(41, 4, 16)
    O3P0(42, 4, 17)
    O3P0(43, 4, 19)
    O3P0(44, 6, 16)
    O3P0(45, 4, 20)
    O3P0(46, 4, 21)
    O3P0(47, 4, 22)
    O3P1(48, 1, 36, 43)
    O3P2(49, 1, 37, 45, 48)
    O3P1(50, 8, 15, 44)
    O3P1(51, 2, 30, 47)
    O3P1(52, 1, 40, 51)
    O3P0(53, 1, 41)
    O3P0(54, 4, 24)
    O3P1(55, 3, 31, 53)
    O3P1(56, 1, 43, 54)
    O3P0(57, 10, 16)
    O3P0(58, 4, 28)
    O3P0(59, 4, 29)
    O3P0(60, 4, 30)
    O3P1(61, 2, 36, 56)
    O3P1(62, 3, 36, 54)
    O3P1(63, 10, 19, 53)
    O3P1(64, 8, 22, 57)
    O3P1(65, 2, 40, 60)
    O3P1(66, 1, 52, 65)
    O3B(2, -0.128814549305e+03)
    O3B(4, 0.104229418850e+04)
    O3B(5, 0.811983220935e+03)
    O3B(7, -0.443324528752e+03)
    O3B(8, 0.904506805268e+03)
    O3B(9, -0.501026918125e+04)
    O3B(11, 0.197209669844e+05)
    O3B(12, -0.251424247013e+05)
    O3B(13, -0.138013677810e+03)
    O3B(14, 0.169329202490e+05)
    O3B(16, 0.352627837493e+05)
    O3B(17, -0.334337897178e+05)
    O3B(18, 0.720500009412e+05)
    O3B(19, 0.116986232065e+05)
    O3B(20, -0.521801943104e+04)
    O3B(21, -0.332486978745e+05)
    O3B(23, -0.177870892015e+05)
    O3B(24, 0.335720198273e+05)
    O3B(25, 0.268323174511e+05)
    O3B(26, -0.933618945467e+05)
    O3B(27, -0.592242307973e+04)
    O3B(28, 0.287777488764e+04)
    O3B(29, 0.393607079595e+05)
    O3B(31, -0.330171644074e+04)
    O3B(32, 0.200362806379e+05)
    O3B(33, -0.975981166385e+04)
    O3B(34, -0.266133829509e+05)
    O3B(35, 0.746650532707e+05)
    O3B(36, 0.120290055844e+05)
    O3B(37, -0.464904653691e+04)
    O3B(38, 0.244129022324e+04)
    O3B(39, -0.273870502550e+05)
    O3B(41, -0.122995471301e+05)
    O3B(42, 0.722408057250e+04)
    O3B(43, 0.290562738593e+05)
    O3B(44, -0.212778140565e+05)
    O3B(45, -0.206522536997e+05)
    O3B(46, 0.263237776823e+05)
    O3B(47, -0.370869661253e+05)
    O3B(48, -0.230814810543e+04)
    O3B(49, 0.210942828415e+04)
    O3B(50, -0.177122132932e+04)
    O3B(51, 0.102466183681e+05)
    O3B(53, -0.556070962327e+03)
    O3B(54, 0.150041418580e+05)
    O3B(55, -0.117657568996e+05)
    O3B(56, -0.442916346835e+04)
    O3B(57, 0.140353851796e+05)
    O3B(58, 0.753137518090e+04)
    O3B(59, -0.911889476033e+04)
    O3B(60, 0.816658803687e+04)
    O3B(61, 0.316234339832e+04)
    O3B(62, -0.202814305330e+04)
    O3B(63, 0.791358291948e+03)
    O3B(64, 0.115450049995e+03)
    O3B(65, -0.158182849802e+04)
    double vv = 240.486 + acc.v;
    double g[3] = {acc.a, acc.b, acc.c};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double e2, d2, ed, dd;
        v2(R[i], e2, d2);
        disp(R[i], ed, dd);
        vv += e2 + ed;
        g[i] += d2 + dd;
    }
    V = vv;
    dV[0] = g[0];
    dV[1] = g[1];
    dV[2] = g[2];
}
#undef O3P0
#undef O3P1
#undef O3P2
#undef O3P3
#undef O3B

}  // namespace o3

struct PesO3 {
    static constexpr int NATOMS = 3;
    static constexpr int ID = CRCL_PES_O3;
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
        constexpr double Cconv = 0.52917721092, Econv = 0.159360144e-2, Gconv = 0.843297564e-3, Eref = -0.19172848;
        double v12[3], v13[3], v23[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            v12[d] = (q[d] - q[3 + d]) * Cconv;
            v13[d] = (q[d] - q[6 + d]) * Cconv;
            v23[d] = (q[3 + d] - q[6 + d]) * Cconv;
        }
        double R[3], iR[3];
        sqrt_rsqrt(v12[0] * v12[0] + v12[1] * v12[1] + v12[2] * v12[2], R[0], iR[0]);
        sqrt_rsqrt(v13[0] * v13[0] + v13[1] * v13[1] + v13[2] * v13[2], R[1], iR[1]);
        sqrt_rsqrt(v23[0] * v23[0] + v23[1] * v23[1] + v23[2] * v23[2], R[2], iR[2]);
        double vk, dV[3];
        o3::pes(R, vk, dV);
        V = vk * Econv + Eref;
        const double f1 = dV[0] * iR[0] * Gconv, f2 = dV[1] * iR[1] * Gconv, f3 = dV[2] * iR[2] * Gconv;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            g[d] = f1 * v12[d] + f2 * v13[d];
            g[3 + d] = -f1 * v12[d] + f3 * v23[d];
            g[6 + d] = -f2 * v13[d] - f3 * v23[d];
        }
        return 0;
    }
};

}  // namespace crcl
