// xi.cuh -- reaction coordinate of the BIMOLEC mechanism family on one structure
// (the ring-polymer centroid), its gradient, and the umbrella "hams" force.
//
// Replaces calc_xi.f90:108-502 (+ calc_com.f90:36-58) and the contraction of
// umbrella.f90:144-174.  The reference assembles the dense (3N)^2 Hessians d2s0/d2s1 on
// every call, including the calls that discard them (modes used by SHAKE and by child
// trajectories), and umbrella contracts d2xi with dxi/m in O((3N)^2).  Here the Hessian is
// never formed: each bond / fragment-pair block is (+-)(r^2 I - r r^T)/r^3, so the
// contraction H.v collapses to one 3-vector per bond and per fragment pair, O(N).
#pragma once
#include "crcl_common.cuh"

namespace crcl {

constexpr int XI_MAXBOND = 4;
constexpr int XI_MAXREAC = 4;
constexpr int XI_MAXAT = 8;  // in-register path; larger systems use the split path

struct Mech {
    int form_num, break_num;
    int bf[XI_MAXBOND][2], bb[XI_MAXBOND][2];  // 0-based
    double fref[XI_MAXBOND], bref[XI_MAXBOND];
    int sum_reacs;
    int frag[XI_MAXAT];  // fragment of each atom, -1 = none
    double mass_reac[XI_MAXREAC];
    double wfrag[XI_MAXAT];  // mass[a] / mass_reac[frag[a]] (0 if the atom is in no fragment)
    double wk[XI_MAXREAC][XI_MAXAT];  // wfrag[a] if frag[a] == k else 0: COM_k = sum_a wk[k][a] x_a
    double R_inf;
    double inv_form, inv_break, inv_pairs;  // 1/form_num, 1/break_num, 1/(number of reactant pairs)
    // umbr_type family: 0 BIMOLEC family (calc_xi.f90:108-502), 1 unimolecular CYCLOREVER / REARRANGE /
    // DECOM_1BOND / ELIMINATION (:673-938: s0 from the reactant bond lengths), 2 ATOM_SHIFT (:523-672)
    int type;
    double freac[XI_MAXBOND], breac[XI_MAXBOND];  // form_reac, break_reac (bonds_ref.f90:81-109)
    int shift_atom, shift_c1, shift_c2;           // ATOM_SHIFT: atom, first coordinate, second (-1: none)
    double shift_lo, shift_hi, shift2_lo, shift2_hi;
    int valid;
};

// s0, s1 of ATOM_SHIFT on structure x (calc_xi.f90:523-560)
CRCL_HD __forceinline__ void xi_shift_s(const Mech& M, const double* x, double& s0, double& s1)
{
    const double a = x[3 * M.shift_atom + M.shift_c1];
    if (M.shift_c2 < 0) {
        s1 = a - M.shift_hi;
        s0 = a - M.shift_lo;
    } else {
        const double b = x[3 * M.shift_atom + M.shift_c2];
        s1 = ((a - M.shift_hi) + (b - M.shift2_hi)) / 2.0;
        s0 = ((a - M.shift_lo) + (b - M.shift2_lo)) / 2.0;
    }
}

// Host-side construction from the MECHA{} tables (1-based atom indices as in the key file,
// calc_rate_read.f90:430-870).  Returns 0, or a negative code: -1 limits exceeded, -2 bad index.
inline int build_mech(Mech& M, int natoms, const double* mass, int form_num, const int* bond_form,
                      int break_num, const int* bond_break, const double* form_ref,
                      const double* break_ref, int sum_reacs, const int* n_reac, const int* at_reac,
                      double R_inf)
{
    if (form_num < 0 || break_num < 0 || form_num > XI_MAXBOND || break_num > XI_MAXBOND ||
        sum_reacs < 2 || sum_reacs > XI_MAXREAC || natoms > XI_MAXAT)
        return -1;
    M = Mech();
    M.form_num = form_num;
    M.break_num = break_num;
    for (int i = 0; i < form_num; i++) {
        M.bf[i][0] = bond_form[2 * i] - 1;
        M.bf[i][1] = bond_form[2 * i + 1] - 1;
        M.fref[i] = form_ref[i];
        if (M.bf[i][0] < 0 || M.bf[i][0] >= natoms || M.bf[i][1] < 0 || M.bf[i][1] >= natoms) return -2;
    }
    for (int i = 0; i < break_num; i++) {
        M.bb[i][0] = bond_break[2 * i] - 1;
        M.bb[i][1] = bond_break[2 * i + 1] - 1;
        M.bref[i] = break_ref[i];
        if (M.bb[i][0] < 0 || M.bb[i][0] >= natoms || M.bb[i][1] < 0 || M.bb[i][1] >= natoms) return -2;
    }
    M.sum_reacs = sum_reacs;
    for (int a = 0; a < XI_MAXAT; a++) M.frag[a] = -1;
    int off = 0;
    for (int k = 0; k < sum_reacs; k++) {
        M.mass_reac[k] = 0.0;
        for (int i = 0; i < n_reac[k]; i++) {
            const int a = at_reac[off + i] - 1;
            if (a < 0 || a >= natoms) return -2;
            M.frag[a] = k;
            M.mass_reac[k] += mass[a];  // calc_rate_read.f90: fragment mass = sum of its atoms
        }
        off += n_reac[k];
    }
    for (int a = 0; a < XI_MAXAT; a++) {
        M.wfrag[a] = (a < natoms && M.frag[a] >= 0) ? mass[a] / M.mass_reac[M.frag[a]] : 0.0;
        for (int k = 0; k < XI_MAXREAC; k++) M.wk[k][a] = (M.frag[a] == k) ? M.wfrag[a] : 0.0;
    }
    M.R_inf = R_inf;
    M.inv_form = form_num ? 1.0 / form_num : 0.0;
    M.inv_break = break_num ? 1.0 / break_num : 0.0;
    M.inv_pairs = 1.0 / (double)((sum_reacs * sum_reacs - sum_reacs) / 2);
    M.valid = 1;
    return 0;
}

// unimolecular mechanisms: bond lists as build_mech sets them (without fragments) + reactant references
inline void mech_set_unimol(Mech& M, const double* form_reac, const double* break_reac)
{
    M.type = 1;
    for (int i = 0; i < M.form_num; i++) M.freac[i] = form_reac[i];
    for (int i = 0; i < M.break_num; i++) M.breac[i] = break_reac[i];
}
// ATOM_SHIFT: shift_coord 1..3 = x,y,z; 4 = (x+y)/2, 5 = (x+z)/2, 6 = (y+z)/2 (calc_rate_read.f90:805-817)
inline int mech_set_atom_shift(Mech& M, int natoms, int shift_atom, int shift_coord, double lo, double hi, double lo2,
                               double hi2)
{
    if (shift_atom < 1 || shift_atom > natoms || shift_coord < 1 || shift_coord > 6) return -2;
    M.type = 2;
    M.shift_atom = shift_atom - 1;
    M.shift_c1 = (shift_coord < 4) ? shift_coord - 1 : ((shift_coord == 6) ? 1 : 0);
    M.shift_c2 = (shift_coord < 4) ? -1 : ((shift_coord == 4) ? 1 : 2);
    M.shift_lo = lo;
    M.shift_hi = hi;
    M.shift2_lo = lo2;
    M.shift2_hi = hi2;
    M.valid = 1;
    return 0;
}

// M w with M = (r^2 I - r r^T)/r^3
CRCL_HD __forceinline__ void proj(const double r[3], double rinv, const double w[3], double o[3])
{
    const double r3 = rinv * rinv * rinv;
    const double rr = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    const double rw = r[0] * w[0] + r[1] * w[1] + r[2] * w[2];
#pragma unroll
    for (int d = 0; d < 3; d++) o[d] = (rr * w[d] - r[d] * rw) * r3;
}

// xi only (no gradient), with the averages as multiplications by host-computed reciprocals (exact
// for 1, 2 and 4 bonds / pairs, one ulp otherwise): what a child trajectory needs every step (verlet.f90:1049-1050 calls
// umbrella mode 1 only to learn the sign of xi_real, recross.f90:597-602).
// Every array index below is a compile-time constant after unrolling, except the bond atoms, which
// index the caller's x (shared memory in the trajectory kernels): nothing lives in local memory.
template <int NAT>
CRCL_HD __forceinline__ double xi_value(const Mech& M, const double* x, double xi_ideal, int mode)
{
    if (M.type == 2) {
        double s0, s1;
        xi_shift_s(M, x, s0, s1);
        return (mode == 1) ? s0 / (s0 - s1) : xi_ideal * s1 + (1 - xi_ideal) * s0;
    }
    if (M.type == 1) {   // unimolecular: both dividing surfaces from the same bonds (value only; divisions as written)
        double s1 = 0.0, s0 = 0.0;
        const double fnum = (double)M.form_num, bnum = (double)M.break_num;
#pragma unroll
        for (int i = 0; i < XI_MAXBOND; i++)
            if (i < M.break_num) {
                const int a1 = M.bb[i][0], a2 = M.bb[i][1];
                const double dx = x[3 * a1] - x[3 * a2], dy = x[3 * a1 + 1] - x[3 * a2 + 1], dz = x[3 * a1 + 2] - x[3 * a2 + 2];
                const double r = sqrt(dx * dx + dy * dy + dz * dz);
                s1 += (r - M.bref[i]) / bnum;
                s0 += (r - M.breac[i]) / bnum;
            }
#pragma unroll
        for (int i = 0; i < XI_MAXBOND; i++)
            if (i < M.form_num) {
                const int a1 = M.bf[i][0], a2 = M.bf[i][1];
                const double dx = x[3 * a1] - x[3 * a2], dy = x[3 * a1 + 1] - x[3 * a2 + 1], dz = x[3 * a1 + 2] - x[3 * a2 + 2];
                const double r = sqrt(dx * dx + dy * dy + dz * dz);
                s1 -= (r - M.fref[i]) / fnum;
                s0 -= (r - M.freac[i]) / fnum;
            }
        return (mode == 1) ? s0 / (s0 - s1) : xi_ideal * s1 + (1 - xi_ideal) * s0;
    }
    double s1 = 0.0;
#pragma unroll
    for (int i = 0; i < XI_MAXBOND; i++)
        if (i < M.break_num) {
            const int a1 = M.bb[i][0], a2 = M.bb[i][1];
            const double dx = x[3 * a1] - x[3 * a2], dy = x[3 * a1 + 1] - x[3 * a2 + 1], dz = x[3 * a1 + 2] - x[3 * a2 + 2];
            s1 += (CRCL_SQRT(dx * dx + dy * dy + dz * dz) - M.bref[i]) * M.inv_break;
        }
#pragma unroll
    for (int i = 0; i < XI_MAXBOND; i++)
        if (i < M.form_num) {
            const int a1 = M.bf[i][0], a2 = M.bf[i][1];
            const double dx = x[3 * a1] - x[3 * a2], dy = x[3 * a1 + 1] - x[3 * a2 + 1], dz = x[3 * a1 + 2] - x[3 * a2 + 2];
            s1 -= (CRCL_SQRT(dx * dx + dy * dy + dz * dz) - M.fref[i]) * M.inv_form;
        }
    // calc_com.f90:36-58; atoms outside fragment k enter with weight 0 (same sums, same order)
    double com[XI_MAXREAC][3];
#pragma unroll
    for (int k = 0; k < XI_MAXREAC; k++) {
        com[k][0] = com[k][1] = com[k][2] = 0.0;
        if (k < M.sum_reacs) {
#pragma unroll
            for (int a = 0; a < NAT; a++) {
#pragma unroll
                for (int d = 0; d < 3; d++) com[k][d] += M.wk[k][a] * x[3 * a + d];
            }
        }
    }
    double s0 = 0.0;
#pragma unroll
    for (int i = 0; i < XI_MAXREAC; i++)
#pragma unroll
        for (int j = i + 1; j < XI_MAXREAC; j++)
            if (j < M.sum_reacs) {
                const double dx = com[j][0] - com[i][0], dy = com[j][1] - com[i][1], dz = com[j][2] - com[i][2];
                s0 += M.R_inf - CRCL_SQRT(dx * dx + dy * dy + dz * dz);
            }
    s0 = s0 * M.inv_pairs;
    return (mode == 1) ? CRCL_DIV(s0, s0 - s1) : xi_ideal * s1 + (1 - xi_ideal) * s0;
}

// mode 1: xi = s0/(s0-s1) (umbrella form); mode 2: xi = xi_ideal*s1 + (1-xi_ideal)*s0.
// x, dxi: [atom][xyz].  If hams != nullptr (mode 1 only) it receives the force of
// umbrella.f90:144-174, -(1/beta) d/dx log f_s with f_s^2 = sum (dxi)^2/m / (2 pi beta).
template <int NAT>
CRCL_HD __forceinline__ void calc_xi(const Mech& M, const double* mass, const double* x,
                                     double xi_ideal, int mode, double& xi, double* dxi,
                                     double* hams, double beta)
{
    double ds0[3 * NAT], ds1[3 * NAT];
#pragma unroll
    for (int t = 0; t < 3 * NAT; t++) {
        ds0[t] = 0.0;
        ds1[t] = 0.0;
    }
    double Rf[XI_MAXBOND][3], Rb[XI_MAXBOND][3], fi[XI_MAXBOND], bi[XI_MAXBOND];
    double s1 = 0.0, s0u = 0.0;   // s0u: s0 of the unimolecular mechanisms (reactant bond references)
    // 1/n of the bond and pair counts from the mechanism (exact for 1, 2, 4; one ulp beside the reference's quotient
    // otherwise): a product instead of a quotient per term
    const double ifnum = M.inv_form, ibnum = M.inv_break, iterms = M.inv_pairs;
    // ATOM_SHIFT has no bond terms, the unimolecular mechanisms no fragment terms
    const int nform = (M.type == 2) ? 0 : M.form_num, nbreak = (M.type == 2) ? 0 : M.break_num;
    const int nreac = (M.type == 0) ? M.sum_reacs : 0;
    for (int i = 0; i < nbreak; i++) {
        const int a1 = M.bb[i][0], a2 = M.bb[i][1];
#pragma unroll
        for (int d = 0; d < 3; d++) Rb[i][d] = x[3 * a1 + d] - x[3 * a2 + d];
        double r;
        sqrt_rsqrt(Rb[i][0] * Rb[i][0] + Rb[i][1] * Rb[i][1] + Rb[i][2] * Rb[i][2], r, bi[i]);
        s1 += (r - M.bref[i]) * ibnum;
        s0u += (r - M.breac[i]) * ibnum;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const double v = Rb[i][d] * bi[i] * ibnum;
            ds1[3 * a1 + d] += v;
            ds1[3 * a2 + d] -= v;
        }
    }
    for (int i = 0; i < nform; i++) {
        const int a1 = M.bf[i][0], a2 = M.bf[i][1];
#pragma unroll
        for (int d = 0; d < 3; d++) Rf[i][d] = x[3 * a1 + d] - x[3 * a2 + d];
        double r;
        sqrt_rsqrt(Rf[i][0] * Rf[i][0] + Rf[i][1] * Rf[i][1] + Rf[i][2] * Rf[i][2], r, fi[i]);
        s1 -= (r - M.fref[i]) * ifnum;
        s0u -= (r - M.freac[i]) * ifnum;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const double v = Rf[i][d] * fi[i] * ifnum;
            ds1[3 * a1 + d] -= v;
            ds1[3 * a2 + d] += v;
        }
    }
    // centres of mass of the reactant fragments (calc_com.f90)
    double com[XI_MAXREAC][3];
#pragma unroll
    for (int k = 0; k < XI_MAXREAC; k++) com[k][0] = com[k][1] = com[k][2] = 0.0;
#pragma unroll
    for (int a = 0; a < NAT; a++) {
        const int k = M.frag[a];
        if (k >= 0) {
#pragma unroll
            for (int d = 0; d < 3; d++) com[k][d] += M.wfrag[a] * x[3 * a + d];
        }
    }
    double s0 = 0.0;
    double Red[XI_MAXREAC * (XI_MAXREAC - 1) / 2][3], ri[XI_MAXREAC * (XI_MAXREAC - 1) / 2];
    int np = 0;
    for (int i = 0; i < nreac; i++)
        for (int j = i + 1; j < nreac; j++, np++) {
#pragma unroll
            for (int d = 0; d < 3; d++) Red[np][d] = com[j][d] - com[i][d];
            double r;
            sqrt_rsqrt(Red[np][0] * Red[np][0] + Red[np][1] * Red[np][1] + Red[np][2] * Red[np][2], r, ri[np]);
            s0 += M.R_inf - r;
#pragma unroll
            for (int a = 0; a < NAT; a++) {
                const int k = M.frag[a];
                if (k == i || k == j) {
                    const double sg = (k == i) ? 1.0 : -1.0;
                    const double w = sg * ri[np] * M.wfrag[a] * iterms;
#pragma unroll
                    for (int d = 0; d < 3; d++) ds0[3 * a + d] += Red[np][d] * w;
                }
            }
        }
    s0 = s0 * iterms;
    if (M.type == 1) {          // calc_xi.f90:722-728, :765 (ds0 = ds1)
        s0 = s0u;
#pragma unroll
        for (int t = 0; t < 3 * NAT; t++) ds0[t] = ds1[t];
    } else if (M.type == 2) {   // calc_xi.f90:523-620
        xi_shift_s(M, x, s0, s1);
        const double w = (M.shift_c2 < 0) ? 1.0 : 0.5;
#pragma unroll
        for (int t = 0; t < 3 * NAT; t++) {
            const bool on = (t == 3 * M.shift_atom + M.shift_c1) || (M.shift_c2 >= 0 && t == 3 * M.shift_atom + M.shift_c2);
            ds0[t] = ds1[t] = on ? w : 0.0;
        }
    }

    if (mode == 1) {
        const double D = s0 - s1;
        const double iD = CRCL_RCP(D);
        xi = s0 * iD;
        const double iD2 = iD * iD;
#pragma unroll
        for (int t = 0; t < 3 * NAT; t++) dxi[t] = (s0 * ds1[t] - s1 * ds0[t]) * iD2;
    } else {
        xi = xi_ideal * s1 + (1 - xi_ideal) * s0;
#pragma unroll
        for (int t = 0; t < 3 * NAT; t++) dxi[t] = xi_ideal * ds1[t] + (1 - xi_ideal) * ds0[t];
    }
    if (!hams) return;

    // ---- hams[b] = coeff2/(coeff1 fs2) * sum_a d2xi(a,b) v_a,  v_a = dxi_a / m_a ----
    double v[3 * NAT], H1v[3 * NAT], H0v[3 * NAT];
    double fs2 = 0.0, d1v = 0.0, d0v = 0.0;
#pragma unroll
    for (int a = 0; a < NAT; a++)
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const int t = 3 * a + d;
            v[t] = CRCL_DIV(dxi[t], mass[a]);
            fs2 += dxi[t] * v[t];
            d1v += ds1[t] * v[t];
            d0v += ds0[t] * v[t];
            H1v[t] = 0.0;
            H0v[t] = 0.0;
        }
    for (int i = 0; i < nform; i++) {  // forming bonds: block = -(r^2 I - r r^T)/r^3
        const int a1 = M.bf[i][0], a2 = M.bf[i][1];
        double w[3], o[3];
#pragma unroll
        for (int d = 0; d < 3; d++) w[d] = v[3 * a1 + d] - v[3 * a2 + d];
        proj(Rf[i], fi[i], w, o);
#pragma unroll
        for (int d = 0; d < 3; d++) {
            H1v[3 * a1 + d] -= o[d] * ifnum;
            H1v[3 * a2 + d] += o[d] * ifnum;
        }
    }
    for (int i = 0; i < nbreak; i++) {  // breaking bonds: block = +(r^2 I - r r^T)/r^3
        const int a1 = M.bb[i][0], a2 = M.bb[i][1];
        double w[3], o[3];
#pragma unroll
        for (int d = 0; d < 3; d++) w[d] = v[3 * a1 + d] - v[3 * a2 + d];
        proj(Rb[i], bi[i], w, o);
#pragma unroll
        for (int d = 0; d < 3; d++) {
            H1v[3 * a1 + d] += o[d] * ibnum;
            H1v[3 * a2 + d] -= o[d] * ibnum;
        }
    }
    np = 0;
    for (int i = 0; i < nreac; i++)
        for (int j = i + 1; j < nreac; j++, np++) {
            double W[3] = {0, 0, 0}, o[3];
#pragma unroll
            for (int a = 0; a < NAT; a++) {
                const int k = M.frag[a];
                if (k == i || k == j) {
                    const double sg = (k == i) ? 1.0 : -1.0;
                    const double w = sg * M.wfrag[a];
#pragma unroll
                    for (int d = 0; d < 3; d++) W[d] += w * v[3 * a + d];
                }
            }
            proj(Red[np], ri[np], W, o);
#pragma unroll
            for (int a = 0; a < NAT; a++) {
                const int k = M.frag[a];
                if (k == i || k == j) {
                    const double sg = (k == i) ? 1.0 : -1.0;
                    const double w = -sg * M.wfrag[a] * iterms;
#pragma unroll
                    for (int d = 0; d < 3; d++) H0v[3 * a + d] += w * o[d];
                }
            }
        }
    if (M.type == 1) {          // d2s0 = d2s1 (calc_xi.f90:913)
#pragma unroll
        for (int t = 0; t < 3 * NAT; t++) H0v[t] = H1v[t];
    }
    const double coeff1 = 2.0 * PI_UMBR * beta;
    fs2 = CRCL_DIV(fs2, coeff1);
    const double pref = CRCL_DIV(-1.0, beta * (coeff1 * fs2));
    const double D = s0 - s1;
    const double iD3 = CRCL_RCP(D * D * D);
    const double cross2 = 2.0 * (s0 * d1v - s1 * d0v);
#pragma unroll
    for (int t = 0; t < 3 * NAT; t++) {
        const double h = ((s0 * H1v[t] + ds0[t] * d1v - ds1[t] * d0v - s1 * H0v[t]) * D -
                          cross2 * (ds0[t] - ds1[t])) *
                         iD3;
        hams[t] = h * pref;
    }
}

// ---- cooperative form for the T threads of one trajectory (fused trajectory kernels) ---------------------------
// calc_xi above keeps ds0/ds1/v/H0v/H1v in per-thread arrays that the bond lists index at run time, i.e. in local
// memory, and the trajectory kernels used to run it redundantly on every thread of a trajectory (ncu, umbrella
// phase of CH4+H x 16 beads: 6 GB of local-memory write-back per 150 steps, 35 % of the stall samples).  Here
// thread t < 3 NAT owns component t = 3 atom + xyz and GATHERS its entries: it walks the bond and fragment-pair
// lists and adds the terms that touch its atom, in the order the scatter loops of calc_xi add them, so every
// number is bit-identical to calc_xi's.  The structure x, the velocities v = dxi/m and ds0/ds1 are exchanged
// through shared memory; no array is indexed at run time.
//   x       shared structure [3 NAT] (the centroid), complete before the call
//   scr     shared scratch [3 * 3 NAT]
//   dxi_sh  shared out [3 NAT];  hams_sh shared out [3 NAT] (written if want_hams; mode 1 only)
// `sync` synchronises the T threads.  The outputs are complete after the caller's next sync.  Every thread
// returns xi.
template <int NAT, int T, class SyncF>
__device__ __forceinline__ double calc_xi_coop(const Mech& M, const double* mass, const double* x, double xi_ideal, int mode,
                                               bool want_hams, double beta, int tig, SyncF sync, double* scr,
                                               double* dxi_sh, double* hams_sh)
{
    constexpr int NC = 3 * NAT, NPASS = (NC + T - 1) / T;
    if constexpr (NPASS > 1) {
        // fewer threads than components (few beads): a thread would repeat the list walks for each of its
        // components (measured: the one-bead start-structure chain of calc_rate ran 5x slower), so every thread
        // evaluates the serial form and thread 0 publishes it -- the same numbers by construction
        double xl[NC], dl[NC], hl[NC], xi;
#pragma unroll
        for (int c = 0; c < NC; c++) xl[c] = x[c];
        calc_xi<NAT>(M, mass, xl, xi_ideal, mode, xi, dl, want_hams ? hl : nullptr, beta);
        sync();
        if (tig == 0) {
#pragma unroll
            for (int c = 0; c < NC; c++) {
                dxi_sh[c] = dl[c];
                if (want_hams) hams_sh[c] = hl[c];
            }
        }
        return xi;
    }
    double* s_ds0 = scr;
    double* s_ds1 = scr + NC;
    double* s_v = scr + 2 * NC;
    // 1/n of the bond and pair counts from the mechanism (exact for 1, 2, 4; one ulp beside the reference's quotient
    // otherwise): a product instead of a quotient per term
    const double ifnum = M.inv_form, ibnum = M.inv_break, iterms = M.inv_pairs;
    const int nform = (M.type == 2) ? 0 : M.form_num, nbreak = (M.type == 2) ? 0 : M.break_num;
    const int nreac = (M.type == 0) ? M.sum_reacs : 0;
    auto sel = [](const double (&r)[3], int d) { return d == 0 ? r[0] : (d == 1 ? r[1] : r[2]); };
    // warps of a CTA-wide trajectory that own no component skip the work and pick xi up from shared memory
    const bool work = (T <= 32) || ((tig & ~31) < NC);
    // fragment centres of mass and pair vectors: the same for every component, kept in registers
    double com[XI_MAXREAC][3];
#pragma unroll
    for (int k = 0; k < XI_MAXREAC; k++) {
        com[k][0] = com[k][1] = com[k][2] = 0.0;
        if (work && k < nreac) {
#pragma unroll
            for (int a = 0; a < NAT; a++)
#pragma unroll
                for (int d = 0; d < 3; d++) com[k][d] += M.wk[k][a] * x[3 * a + d];
        }
    }
    double s0 = 0.0, s1 = 0.0, s0u = 0.0;
#pragma unroll
    for (int p = 0; p < NPASS; p++) {
        if (!work) break;
        const int t = tig + p * T;
        const bool active = t < NC;
        const int tt = active ? t : 0, a = tt / 3, d = tt - 3 * a;
        double ds1 = 0.0, ds0 = 0.0;
        s1 = 0.0;
        s0u = 0.0;
        for (int i = 0; i < nbreak; i++) {
            const int a1 = M.bb[i][0], a2 = M.bb[i][1];
            const double R[3] = {x[3 * a1] - x[3 * a2], x[3 * a1 + 1] - x[3 * a2 + 1], x[3 * a1 + 2] - x[3 * a2 + 2]};
            double r, ri;
            sqrt_rsqrt(R[0] * R[0] + R[1] * R[1] + R[2] * R[2], r, ri);
            s1 += (r - M.bref[i]) * ibnum;
            s0u += (r - M.breac[i]) * ibnum;
            const double v = sel(R, d) * ri * ibnum;
            if (a == a1) ds1 += v;
            if (a == a2) ds1 -= v;
        }
        for (int i = 0; i < nform; i++) {
            const int a1 = M.bf[i][0], a2 = M.bf[i][1];
            const double R[3] = {x[3 * a1] - x[3 * a2], x[3 * a1 + 1] - x[3 * a2 + 1], x[3 * a1 + 2] - x[3 * a2 + 2]};
            double r, ri;
            sqrt_rsqrt(R[0] * R[0] + R[1] * R[1] + R[2] * R[2], r, ri);
            s1 -= (r - M.fref[i]) * ifnum;
            s0u -= (r - M.freac[i]) * ifnum;
            const double v = sel(R, d) * ri * ifnum;
            if (a == a1) ds1 -= v;
            if (a == a2) ds1 += v;
        }
        s0 = 0.0;
        const int ka = M.frag[a];
        const double wa = M.wfrag[a];
#pragma unroll
        for (int i = 0; i < XI_MAXREAC; i++)
#pragma unroll
            for (int j = i + 1; j < XI_MAXREAC; j++)
                if (j < nreac) {
                    const double Red[3] = {com[j][0] - com[i][0], com[j][1] - com[i][1], com[j][2] - com[i][2]};
                    double r, ri;
                    sqrt_rsqrt(Red[0] * Red[0] + Red[1] * Red[1] + Red[2] * Red[2], r, ri);
                    s0 += M.R_inf - r;
                    if (ka == i || ka == j) {
                        const double sg = (ka == i) ? 1.0 : -1.0;
                        ds0 += sel(Red, d) * (sg * ri * wa * iterms);
                    }
                }
        s0 = s0 * iterms;
        if (M.type == 1) {
            s0 = s0u;
            ds0 = ds1;
        } else if (M.type == 2) {
            xi_shift_s(M, x, s0, s1);
            const double w = (M.shift_c2 < 0) ? 1.0 : 0.5;
            const bool on = (tt == 3 * M.shift_atom + M.shift_c1) || (M.shift_c2 >= 0 && tt == 3 * M.shift_atom + M.shift_c2);
            ds0 = ds1 = on ? w : 0.0;
        }
        double dx;
        if (mode == 1) {
            const double D = s0 - s1;
            const double iD = CRCL_RCP(D);
            dx = (s0 * ds1 - s1 * ds0) * (iD * iD);
        } else {
            dx = xi_ideal * ds1 + (1 - xi_ideal) * ds0;
        }
        if (active) {
            dxi_sh[t] = dx;
            s_ds0[t] = ds0;
            s_ds1[t] = ds1;
            s_v[t] = CRCL_DIV(dx, mass[a]);
        }
    }
    double xi = (mode == 1) ? s0 * CRCL_RCP(s0 - s1) : xi_ideal * s1 + (1 - xi_ideal) * s0;
    if (T > 32 && tig == 0) scr[3 * NC] = xi;
    if (T > 32 || want_hams) sync();
    if (!work) xi = scr[3 * NC];
    if (!want_hams || !work) return xi;
    double fs2 = 0.0, d1v = 0.0, d0v = 0.0;
#pragma unroll
    for (int t = 0; t < NC; t++) {
        const double v = s_v[t];
        fs2 += dxi_sh[t] * v;
        d1v += s_ds1[t] * v;
        d0v += s_ds0[t] * v;
    }
    const double coeff1 = 2.0 * PI_UMBR * beta;
    fs2 = CRCL_DIV(fs2, coeff1);
    const double pref = CRCL_DIV(-1.0, beta * (coeff1 * fs2));
    const double D = s0 - s1;
    const double iD3 = CRCL_RCP(D * D * D);
    const double cross2 = 2.0 * (s0 * d1v - s1 * d0v);
#pragma unroll
    for (int p = 0; p < NPASS; p++) {
        const int t = tig + p * T;
        const bool active = t < NC;
        const int tt = active ? t : 0, a = tt / 3, d = tt - 3 * a;
        double H1 = 0.0, H0 = 0.0;
        for (int i = 0; i < nform; i++) {
            const int a1 = M.bf[i][0], a2 = M.bf[i][1];
            if (a == a1 || a == a2) {
                const double R[3] = {x[3 * a1] - x[3 * a2], x[3 * a1 + 1] - x[3 * a2 + 1], x[3 * a1 + 2] - x[3 * a2 + 2]};
                const double ri = CRCL_RSQRT(R[0] * R[0] + R[1] * R[1] + R[2] * R[2]);
                const double w[3] = {s_v[3 * a1] - s_v[3 * a2], s_v[3 * a1 + 1] - s_v[3 * a2 + 1], s_v[3 * a1 + 2] - s_v[3 * a2 + 2]};
                double o[3];
                proj(R, ri, w, o);
                const double od = sel(o, d) * ifnum;
                if (a == a1) H1 -= od;
                if (a == a2) H1 += od;
            }
        }
        for (int i = 0; i < nbreak; i++) {
            const int a1 = M.bb[i][0], a2 = M.bb[i][1];
            if (a == a1 || a == a2) {
                const double R[3] = {x[3 * a1] - x[3 * a2], x[3 * a1 + 1] - x[3 * a2 + 1], x[3 * a1 + 2] - x[3 * a2 + 2]};
                const double ri = CRCL_RSQRT(R[0] * R[0] + R[1] * R[1] + R[2] * R[2]);
                const double w[3] = {s_v[3 * a1] - s_v[3 * a2], s_v[3 * a1 + 1] - s_v[3 * a2 + 1], s_v[3 * a1 + 2] - s_v[3 * a2 + 2]};
                double o[3];
                proj(R, ri, w, o);
                const double od = sel(o, d) * ibnum;
                if (a == a1) H1 += od;
                if (a == a2) H1 -= od;
            }
        }
        const int ka = M.frag[a];
        const double wa = M.wfrag[a];
#pragma unroll
        for (int i = 0; i < XI_MAXREAC; i++)
#pragma unroll
            for (int j = i + 1; j < XI_MAXREAC; j++)
                if (j < nreac && (ka == i || ka == j)) {
                    const double Red[3] = {com[j][0] - com[i][0], com[j][1] - com[i][1], com[j][2] - com[i][2]};
                    const double ri = CRCL_RSQRT(Red[0] * Red[0] + Red[1] * Red[1] + Red[2] * Red[2]);
                    double W[3] = {0, 0, 0}, o[3];
#pragma unroll
                    for (int b = 0; b < NAT; b++) {
                        const int kb = M.frag[b];
                        if (kb == i || kb == j) {
                            const double w = ((kb == i) ? 1.0 : -1.0) * M.wfrag[b];
#pragma unroll
                            for (int e = 0; e < 3; e++) W[e] += w * s_v[3 * b + e];
                        }
                    }
                    proj(Red, ri, W, o);
                    const double sg = (ka == i) ? 1.0 : -1.0;
                    H0 += (-sg * wa * iterms) * sel(o, d);
                }
        if (M.type == 1) H0 = H1;
        const double ds0 = s_ds0[tt], ds1 = s_ds1[tt];
        const double h = ((s0 * H1 + ds0 * d1v - ds1 * d0v - s1 * H0) * D - cross2 * (ds0 - ds1)) * iD3;
        if (active) hams_sh[t] = h * pref;
    }
    return xi;
}


}  // namespace crcl
