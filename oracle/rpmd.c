/*
 * rpmd.c -- CPU oracle: RPMD integrator, reaction coordinate, umbrella bias,
 * Bennett-Chandler constraint, thermostats and the recrossing child body.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Parity UNPINNED by the reference.
 *
 * Literal restatement (graded paths only: no NPT / AFM / mirrors / box walls / bias lists;
 * the periodic wrap of verlet.f90:591-641 in its plain-box form, not the VASP fractional branch) of
 *   verlet.f90:65-1308          orc_verlet          (operation order: SURVEY.md 3.5)
 *   rfft.f90:35-60, irfft.f90:35-62  orc_rfft       (both are forward DFT, real part, 1/sqrt(N))
 *   get_centroid.f90:67-82      orc_get_centroid
 *   mdinit.f90:40-172           orc_mdinit
 *   andersen.f90:36-74          orc_andersen        (normals from a pluggable source)
 *   nhc.f90:34-170              orc_nhc
 *   transrot.f90:36-236 + invert.f90:38-123   orc_transrot (totmass*nbeads twice, F9)
 *   calc_xi.f90:63-502          orc_calc_xi         (BIMOLEC family)
 *   calc_com.f90:36-58          orc_calc_com
 *   umbrella.f90:66-175         orc_umbrella
 *   constrain_q.f90:30-112      orc_constrain_q
 *   constrain_p.f90:30-75       orc_constrain_p
 *   recross_serial.f90:172-229  orc_recross_pair    (one +/- child pair)
 *   gradient.f90:140-207        orc_gradient        (analytic PES dispatch, one bead/call)
 *   rpmd_check.f90:69-112       orc_rpmd_check      (the drivers' per-step guards, as status bits)
 *
 * The reference's RNG (gfortran random_number + Marsaglia polar, andersen.f90:84-126) is
 * not reproducible outside one gfortran build.  Normal deviates come either from an
 * injected array (exact-parity tests) or from the counter-based Philox4x32-10 +
 * Box-Muller stream defined in DESIGN.md (the product's RNG; statistical parity only).
 */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "oracle_real.h"
#include "oracle.h"
#include "rpmd.h"

#define ORC_PI_QMDFF 3.1415926535897932384626433832795029 /* qmdff.f90:44 (verlet) */
#define ORC_PI_UMBR 3.1415926535897932384                 /* umbrella.f90:98, recross_serial.f90:79 */

/* ------------------------------------------------------------------------------------
 * counter-based RNG: Philox4x32-10 (Salmon et al., SC'11; Random123 v1.09) + Box-Muller.
 * counter = (pair, bead, event, traj), key = (seed_lo, seed_hi); see DESIGN.md "RNG".
 * ------------------------------------------------------------------------------------ */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    int r;
    for (r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    philox4x32_10(ctr, key, out);
}

/* two standard normals for (seed, traj, event, bead, pair) */
void oracle_rng_normal_pair(uint64_t seed, uint32_t traj, uint32_t event, uint32_t bead,
                            uint32_t pair, double z[2])
{
    uint32_t ctr[4], key[2], o[4];
    double u1, u2, r, a;
    ctr[0] = pair;
    ctr[1] = bead;
    ctr[2] = event;
    ctr[3] = traj;
    key[0] = (uint32_t)seed;
    key[1] = (uint32_t)(seed >> 32);
    philox4x32_10(ctr, key, o);
    /* 53-bit uniforms in (0,1]: ((hi<<21 | lo>>11) + 1) * 2^-53 */
    u1 = ((double)((((uint64_t)o[0]) << 21) | (o[1] >> 11)) + 1.0) * (1.0 / 9007199254740992.0);
    u2 = ((double)((((uint64_t)o[2]) << 21) | (o[3] >> 11)) + 1.0) * (1.0 / 9007199254740992.0);
    r = sqrt(-2.0 * log(u1));
    a = 6.283185307179586476925286766559 * u2;
    z[0] = r * cos(a);
    z[1] = r * sin(a);
}

/* ------------------------------------------------------------------------------------ */

orc_sys *oracle_sys_create(int natoms, int nbeads, const double *mass, const int *at_move,
                           double beta, double dt, int pes)
{
    orc_sys *s = (orc_sys *)calloc(1, sizeof(orc_sys));
    int i;
    s->natoms = natoms;
    s->nbeads = nbeads;
    s->pes = pes;
    s->beta = beta;
    s->dt = dt;
    s->mass = (double *)malloc(sizeof(double) * natoms);
    s->at_move = (int *)malloc(sizeof(int) * natoms);
    for (i = 0; i < natoms; i++) {
        s->mass[i] = mass[i];
        s->at_move[i] = at_move ? at_move[i] : 1;
    }
    s->q = (double *)calloc((size_t)3 * natoms * nbeads, sizeof(double));
    s->p = (double *)calloc((size_t)3 * natoms * nbeads, sizeof(double));
    s->thermostat = 0;
    s->andersen_step = 0;
    s->nve = 0;
    s->k_force = 0.0;
    s->seed = 0;
    s->traj = 0;
    s->event = 0;
    s->inject = 0;
    s->inject_len = 0;
    s->inject_pos = 0;
    s->custom_grad = 0;
    return s;
}

void oracle_sys_free(orc_sys *s)
{
    if (!s) return;
    free(s->mass);
    free(s->at_move);
    free(s->q);
    free(s->p);
    free(s);
}

void oracle_sys_set_mecha(orc_sys *s, int form_num, const int *bond_form, int break_num,
                          const int *bond_break, const double *form_ref, const double *break_ref,
                          int sum_reacs, const int *n_reac, const int *at_reac, double R_inf)
{
    int i, k, off = 0;
    s->form_num = form_num;
    s->break_num = break_num;
    for (i = 0; i < form_num; i++) {
        s->bond_form[i][0] = bond_form[2 * i];
        s->bond_form[i][1] = bond_form[2 * i + 1];
        s->form_ref[i] = form_ref[i];
    }
    for (i = 0; i < break_num; i++) {
        s->bond_break[i][0] = bond_break[2 * i];
        s->bond_break[i][1] = bond_break[2 * i + 1];
        s->break_ref[i] = break_ref[i];
    }
    s->sum_reacs = sum_reacs;
    for (k = 0; k < sum_reacs; k++) {
        s->n_reac[k] = n_reac[k];
        s->mass_reac[k] = 0.0;
        for (i = 0; i < n_reac[k]; i++) {
            s->at_reac[k][i] = at_reac[off + i];
            /* calc_rate_read.f90: mass_reac = sum of the fragment's atom masses */
            s->mass_reac[k] = s->mass_reac[k] + s->mass[at_reac[off + i]];
        }
        off += n_reac[k];
    }
    s->R_inf = R_inf;
}

/* unimolecular mechanisms: bond lists as set by oracle_sys_set_mecha plus the reactant references */
void oracle_sys_set_unimol(orc_sys *s, const double *form_reac, const double *break_reac)
{
    int i;
    s->umbr_type = 1;
    for (i = 0; i < s->form_num; i++) s->form_reac[i] = form_reac[i];
    for (i = 0; i < s->break_num; i++) s->break_reac[i] = break_reac[i];
}
void oracle_sys_set_atom_shift(orc_sys *s, int shift_atom, int shift_coord, double shift_lo, double shift_hi,
                               double shift2_lo, double shift2_hi)
{
    s->umbr_type = 2;
    s->shift_atom = shift_atom;
    s->shift_coord = shift_coord;
    s->shift_lo = shift_lo;
    s->shift_hi = shift_hi;
    s->shift2_lo = shift2_lo;
    s->shift2_hi = shift2_hi;
}

void oracle_sys_set_thermostat(orc_sys *s, int thermostat, int andersen_step, double kelvin,
                               double nose_q)
{
    s->thermostat = thermostat;
    s->andersen_step = andersen_step;
    s->kelvin = kelvin;
    s->nose_q = nose_q;
}
void oracle_sys_set_kforce(orc_sys *s, double k_force) { s->k_force = k_force; }
void oracle_sys_set_rng(orc_sys *s, uint64_t seed, uint32_t traj, uint32_t event)
{
    s->seed = seed;
    s->traj = traj;
    s->event = event;
}
void oracle_sys_inject_normals(orc_sys *s, const double *z, long n)
{
    s->inject = z;
    s->inject_len = n;
    s->inject_pos = 0;
}
void oracle_sys_set_custom_grad(orc_sys *s, orc_custom_grad_fn fn) { s->custom_grad = fn; }
double *oracle_sys_q(orc_sys *s) { return s->q; }
double *oracle_sys_p(orc_sys *s) { return s->p; }
uint32_t oracle_sys_get_event(orc_sys *s) { return s->event; }
void oracle_sys_get_nhc(orc_sys *s, double *v4q4)
{
    int i;
    for (i = 0; i < 4; i++) {
        v4q4[i] = s->vnh[i];
        v4q4[4 + i] = s->qnh[i];
    }
}

#define Q(s, i, j, k) ((s)->q[((size_t)(k) * (s)->natoms + (j)) * 3 + (i)])
#define P(s, i, j, k) ((s)->p[((size_t)(k) * (s)->natoms + (j)) * 3 + (i)])
#define D3(a, n, i, j, k) ((a)[((size_t)(k) * (n) + (j)) * 3 + (i)])
#define D2(a, i, j) ((a)[(size_t)(j) * 3 + (i)])

/* gradient.f90:140-207 -- one bead per call */
void orc_gradient(orc_sys *s, const double *xyz, double *e, double *g)
{
    int info = 0;
    if (s->custom_grad) {
        s->custom_grad(xyz, e, g, s->natoms);
        return;
    }
    switch (s->pes) {
    case ORC_PES_H3:
        oracle_egrad_h3_real(xyz, s->natoms, 1, e, g, &info);
        break;
    case ORC_PES_OH3:
        oracle_egrad_oh3_real(xyz, s->natoms, 1, e, g, &info);
        break;
    case ORC_PES_CH4H:
        oracle_egrad_ch4h_real(xyz, s->natoms, 1, e, g, &info);
        break;
    case ORC_PES_BRH2:
        oracle_egrad_brh2_real(xyz, s->natoms, 1, e, g, &info);
        break;
    case ORC_PES_O3:
        oracle_egrad_o3_real(xyz, s->natoms, 1, e, g, &info);
        break;
    case ORC_PES_CH4OH:
        oracle_egrad_ch4oh_real(xyz, s->natoms, 1, e, g, &info);
        break;
    case ORC_PES_GEH4OH:
        oracle_egrad_geh4oh_real(xyz, s->natoms, 1, e, g, &info);
        break;
    case ORC_PES_CH4CN:
        oracle_egrad_ch4cn_real(xyz, s->natoms, 1, e, g, &info);
        break;
    case ORC_PES_CLNH3:
        oracle_egrad_clnh3_real(xyz, s->natoms, 1, e, g, &info);
        break;
    case ORC_PES_NH3OH:
        oracle_egrad_nh3oh_real(xyz, s->natoms, 1, e, g, &info);
        break;
    case ORC_PES_H2CO:
        oracle_egrad_h2co_real(xyz, s->natoms, 1, e, g, &info);
        break;
    default:
        *e = 0.0;
        memset(g, 0, sizeof(double) * 3 * s->natoms);
    }
}

/* get_centroid.f90:67-82 */
void orc_get_centroid(orc_sys *s, double *centroid)
{
    int i, j, k;
    for (j = 0; j < s->natoms; j++)
        for (i = 0; i < 3; i++) {
            double c = 0.0;
            for (k = 0; k < s->nbeads; k++) c = c + Q(s, i, j, k);
            D2(centroid, i, j) = c / s->nbeads;
        }
}

/* rfft.f90 / irfft.f90: x <- factor * real(forward DFT(x)), factor = dsqrt(1.d0/N).
 * FFTW's complex forward DFT is restated as the dense cosine sum (exact to rounding). */
static void orc_rfft(double *x, int stride, int N, const double *costab)
{
    double tmp[ORC_MAXBEADS];
    double factor = sqrt(1.0 / N);
    int k, j;
    for (k = 0; k < N; k++) {
        double acc = 0.0;
        for (j = 0; j < N; j++) acc += x[(size_t)j * stride] * costab[(k * j) % N];
        tmp[k] = acc;
    }
    for (k = 0; k < N; k++) x[(size_t)k * stride] = factor * tmp[k];
}

/* calc_com.f90:36-58 */
static void orc_calc_com(const orc_sys *s, const double *coords, double com[ORC_MAXREAC][3])
{
    int k, i, j;
    for (k = 0; k < s->sum_reacs; k++)
        for (j = 0; j < 3; j++) com[k][j] = 0.0;
    for (k = 0; k < s->sum_reacs; k++)
        for (i = 0; i < s->n_reac[k]; i++) {
            int atom = s->at_reac[k][i];
            for (j = 0; j < 3; j++)
                com[k][j] = com[k][j] + s->mass[atom] * D2(coords, j, atom) / (s->mass_reac[k]);
        }
}

#define H4(a, n, i1, j1, i2, j2) ((a)[(((size_t)(j2) * 3 + (i2)) * (n) + (j1)) * 3 + (i1)])

static void add_hess_block(double *h, int n, int a1, int a2, double sgn, const double d[6],
                           double scale)
{
    /* d = dxx,dxy,dxz,dyy,dyz,dzz ; adds sgn*d/scale to block (a1,a2) */
    static const int map[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    int i1, i2;
    for (i1 = 0; i1 < 3; i1++)
        for (i2 = 0; i2 < 3; i2++)
            H4(h, n, i1, a1, i2, a2) = H4(h, n, i1, a1, i2, a2) + sgn * (d[map[i1][i2]] / scale);
}

/* calc_xi.f90:523-672, ATOM_SHIFT: one Cartesian coordinate (or the mean of two) of one atom */
static void orc_calc_xi_shift(const orc_sys *s, const double *coords, double xi_ideal, double *xi_act,
                              double *dxi_act, double *d2xi_act, int mode)
{
    const int n = s->natoms, a = s->shift_atom, sc = s->shift_coord;
    double s0, s1;
    double *ds0 = (double *)calloc((size_t)3 * n, sizeof(double));
    double *ds1 = (double *)calloc((size_t)3 * n, sizeof(double));
    int c1 = 0, c2 = -1, i;
    if (sc < 4) {
        c1 = sc - 1;
        s1 = D2(coords, c1, a) - s->shift_hi;
        s0 = D2(coords, c1, a) - s->shift_lo;
        D2(ds1, c1, a) = 1.0;
        D2(ds0, c1, a) = 1.0;
    } else {
        c1 = (sc == 6) ? 1 : 0;
        c2 = (sc == 4) ? 1 : 2;
        s1 = ((D2(coords, c1, a) - s->shift_hi) + (D2(coords, c2, a) - s->shift2_hi)) / 2.0;
        s0 = ((D2(coords, c1, a) - s->shift_lo) + (D2(coords, c2, a) - s->shift2_lo)) / 2.0;
        D2(ds1, c1, a) = 0.5;
        D2(ds1, c2, a) = 0.5;
        D2(ds0, c1, a) = 0.5;
        D2(ds0, c2, a) = 0.5;
    }
    if (mode == 1) {
        *xi_act = s0 / (s0 - s1);
        for (i = 0; i < 3 * n; i++) dxi_act[i] = (s0 * ds1[i] - s1 * ds0[i]) / ((s0 - s1) * (s0 - s1));
    } else {
        *xi_act = xi_ideal * s1 + (1 - xi_ideal) * s0;
        for (i = 0; i < 3 * n; i++) dxi_act[i] = xi_ideal * ds1[i] + (1 - xi_ideal) * ds0[i];
    }
    if (d2xi_act) { /* d2s0 = d2s1 = 0 (:633-634) */
        int i1, j1, i2, j2;
        memset(d2xi_act, 0, sizeof(double) * 9 * n * n);
        if (mode == 1)
            for (i1 = 0; i1 < 3; i1++)
                for (j1 = 0; j1 < n; j1++)
                    for (i2 = 0; i2 < 3; i2++)
                        for (j2 = 0; j2 < n; j2++)
                            H4(d2xi_act, n, i1, j1, i2, j2) =
                                ((s0 * 0.0 + D2(ds0, i2, j2) * D2(ds1, i1, j1) - D2(ds1, i2, j2) * D2(ds0, i1, j1) -
                                  s1 * 0.0) * (s0 - s1) -
                                 2.0 * (s0 * D2(ds1, i1, j1) - s1 * D2(ds0, i1, j1)) *
                                     (D2(ds0, i2, j2) - D2(ds1, i2, j2))) /
                                ((s0 - s1) * (s0 - s1) * (s0 - s1));
    }
    free(ds0);
    free(ds1);
}

/* calc_xi.f90:63-502, BIMOLEC family, and :673-938, the unimolecular mechanisms (s0 from the reactant
 * reference bond lengths, ds0 = ds1, d2s0 = d2s1).  mode 1: umbrella form, mode 2: recrossing form.
 * d2xi may be NULL (the value is not needed by the caller; the reference always builds it). */
void orc_calc_xi(const orc_sys *s, const double *coords, double xi_ideal, double *xi_act,
                 double *dxi_act, double *d2xi_act, int mode)
{
    const int n = s->natoms;
    const int unimol = (s->umbr_type == 1);
    double R_f[ORC_MAXBOND][3], R_b[ORC_MAXBOND][3], form_act[ORC_MAXBOND], break_act[ORC_MAXBOND];
    double Red[ORC_MAXREAC][ORC_MAXREAC][3], r_eds[ORC_MAXREAC][ORC_MAXREAC];
    double com[ORC_MAXREAC][3];
    double s0, s1, Rinv;
    double *ds0 = (double *)calloc((size_t)3 * n, sizeof(double));
    double *ds1 = (double *)calloc((size_t)3 * n, sizeof(double));
    int i, j, k, l, s0_terms;
    const float fnum = (float)s->form_num, bnum = (float)s->break_num; /* real(form_num) */
    float fterms;

    if (s->umbr_type == 2) {
        free(ds0);
        free(ds1);
        orc_calc_xi_shift(s, coords, xi_ideal, xi_act, dxi_act, d2xi_act, mode);
        return;
    }
    for (i = 0; i < s->form_num; i++) {
        int a1 = s->bond_form[i][0], a2 = s->bond_form[i][1];
        R_f[i][0] = D2(coords, 0, a1) - D2(coords, 0, a2);
        R_f[i][1] = D2(coords, 1, a1) - D2(coords, 1, a2);
        R_f[i][2] = D2(coords, 2, a1) - D2(coords, 2, a2);
        form_act[i] = sqrt(R_f[i][0] * R_f[i][0] + R_f[i][1] * R_f[i][1] + R_f[i][2] * R_f[i][2]);
    }
    for (i = 0; i < s->break_num; i++) {
        int a1 = s->bond_break[i][0], a2 = s->bond_break[i][1];
        R_b[i][0] = D2(coords, 0, a1) - D2(coords, 0, a2);
        R_b[i][1] = D2(coords, 1, a1) - D2(coords, 1, a2);
        R_b[i][2] = D2(coords, 2, a1) - D2(coords, 2, a2);
        break_act[i] = sqrt(R_b[i][0] * R_b[i][0] + R_b[i][1] * R_b[i][1] + R_b[i][2] * R_b[i][2]);
    }
    s1 = 0.0;
    for (i = 0; i < s->break_num; i++) s1 = s1 + (break_act[i] - s->break_ref[i]) / bnum;
    for (i = 0; i < s->form_num; i++) s1 = s1 - (form_act[i] - s->form_ref[i]) / fnum;

    s0 = 0.0;
    if (unimol) { /* calc_xi.f90:722-728 */
        for (i = 0; i < s->break_num; i++) s0 = s0 + (break_act[i] - s->break_reac[i]) / bnum;
        for (i = 0; i < s->form_num; i++) s0 = s0 - (form_act[i] - s->form_reac[i]) / fnum;
    } else
        orc_calc_com(s, coords, com);
    for (i = 0; i < (unimol ? 0 : s->sum_reacs); i++)
        for (j = i + 1; j < s->sum_reacs; j++) {
            Red[i][j][0] = com[j][0] - com[i][0];
            Red[i][j][1] = com[j][1] - com[i][1];
            Red[i][j][2] = com[j][2] - com[i][2];
            r_eds[i][j] = sqrt(Red[i][j][0] * Red[i][j][0] + Red[i][j][1] * Red[i][j][1] +
                               Red[i][j][2] * Red[i][j][2]);
            s0 = s0 + (s->R_inf - r_eds[i][j]);
        }
    s0_terms = (s->sum_reacs * s->sum_reacs - s->sum_reacs) / 2;
    fterms = (float)s0_terms;
    if (!unimol) s0 = s0 / fterms;

    if (mode == 1)
        *xi_act = s0 / (s0 - s1);
    else
        *xi_act = xi_ideal * s1 + (1 - xi_ideal) * s0;

    /* gradient of s1 */
    for (i = 0; i < s->form_num; i++) {
        int a1 = s->bond_form[i][0], a2 = s->bond_form[i][1];
        Rinv = 1.0 / form_act[i];
        for (k = 0; k < 3; k++) {
            D2(ds1, k, a1) = D2(ds1, k, a1) - R_f[i][k] * Rinv / fnum;
            D2(ds1, k, a2) = D2(ds1, k, a2) + R_f[i][k] * Rinv / fnum;
        }
    }
    for (i = 0; i < s->break_num; i++) {
        int a1 = s->bond_break[i][0], a2 = s->bond_break[i][1];
        Rinv = 1.0 / break_act[i];
        for (k = 0; k < 3; k++) {
            D2(ds1, k, a1) = D2(ds1, k, a1) + R_b[i][k] * Rinv / bnum;
            D2(ds1, k, a2) = D2(ds1, k, a2) - R_b[i][k] * Rinv / bnum;
        }
    }
    /* gradient of s0 */
    if (unimol) memcpy(ds0, ds1, sizeof(double) * 3 * n); /* ds0 = ds1 (:765) */
    for (i = 0; i < (unimol ? 0 : s->sum_reacs); i++)
        for (j = i + 1; j < s->sum_reacs; j++) {
            Rinv = 1.0 / r_eds[i][j];
            for (k = 0; k < s->n_reac[i]; k++) {
                int atom = s->at_reac[i][k];
                for (l = 0; l < 3; l++)
                    D2(ds0, l, atom) = D2(ds0, l, atom) + Red[i][j][l] * Rinv * s->mass[atom] /
                                                              s->mass_reac[i] / fterms;
            }
            for (k = 0; k < s->n_reac[j]; k++) {
                int atom = s->at_reac[j][k];
                for (l = 0; l < 3; l++)
                    D2(ds0, l, atom) = D2(ds0, l, atom) - Red[i][j][l] * Rinv * s->mass[atom] /
                                                              s->mass_reac[j] / fterms;
            }
        }
    if (mode == 1) {
        for (i = 0; i < 3 * n; i++)
            dxi_act[i] = (s0 * ds1[i] - s1 * ds0[i]) / ((s0 - s1) * (s0 - s1));
    } else {
        for (i = 0; i < 3 * n; i++) dxi_act[i] = xi_ideal * ds1[i] + (1 - xi_ideal) * ds0[i];
    }

    if (d2xi_act) {
        size_t hs = (size_t)9 * n * n;
        double *d2s0 = (double *)calloc(hs, sizeof(double));
        double *d2s1 = (double *)calloc(hs, sizeof(double));
        double d[6], R3;
        for (i = 0; i < s->form_num; i++) {
            int a1 = s->bond_form[i][0], a2 = s->bond_form[i][1];
            const double *r = R_f[i];
            Rinv = 1.0 / form_act[i];
            R3 = Rinv * Rinv * Rinv;
            d[0] = -(r[1] * r[1] + r[2] * r[2]) * R3;
            d[3] = -(r[2] * r[2] + r[0] * r[0]) * R3;
            d[5] = -(r[0] * r[0] + r[1] * r[1]) * R3;
            d[1] = r[0] * r[1] * R3;
            d[2] = r[0] * r[2] * R3;
            d[4] = r[1] * r[2] * R3;
            add_hess_block(d2s1, n, a1, a1, +1.0, d, fnum);
            add_hess_block(d2s1, n, a1, a2, -1.0, d, fnum);
            add_hess_block(d2s1, n, a2, a1, -1.0, d, fnum);
            add_hess_block(d2s1, n, a2, a2, +1.0, d, fnum);
        }
        for (i = 0; i < s->break_num; i++) {
            int a1 = s->bond_break[i][0], a2 = s->bond_break[i][1];
            const double *r = R_b[i];
            Rinv = 1.0 / break_act[i];
            R3 = Rinv * Rinv * Rinv;
            d[0] = (r[1] * r[1] + r[2] * r[2]) * R3;
            d[3] = (r[2] * r[2] + r[0] * r[0]) * R3;
            d[5] = (r[0] * r[0] + r[1] * r[1]) * R3;
            d[1] = -r[0] * r[1] * R3;
            d[2] = -r[0] * r[2] * R3;
            d[4] = -r[1] * r[2] * R3;
            add_hess_block(d2s1, n, a1, a1, +1.0, d, bnum);
            add_hess_block(d2s1, n, a1, a2, -1.0, d, bnum);
            add_hess_block(d2s1, n, a2, a1, -1.0, d, bnum);
            add_hess_block(d2s1, n, a2, a2, +1.0, d, bnum);
        }
        if (unimol) memcpy(d2s0, d2s1, sizeof(double) * hs); /* d2s0 = d2s1 (:913) */
        for (i = 0; i < (unimol ? 0 : s->sum_reacs); i++)
            for (j = i + 1; j < s->sum_reacs; j++) {
                const double *r = Red[i][j];
                double dm[6], mf;
                int m, ka, la;
                Rinv = 1.0 / r_eds[i][j];
                R3 = Rinv * Rinv * Rinv;
                d[0] = -(r[1] * r[1] + r[2] * r[2]) * R3;
                d[3] = -(r[2] * r[2] + r[0] * r[0]) * R3;
                d[5] = -(r[0] * r[0] + r[1] * r[1]) * R3;
                d[1] = r[0] * r[1] * R3;
                d[2] = r[0] * r[2] * R3;
                d[4] = r[1] * r[2] * R3;
                for (k = 0; k < s->n_reac[i]; k++) {
                    ka = s->at_reac[i][k];
                    for (l = 0; l < s->n_reac[i]; l++) {
                        la = s->at_reac[i][l];
                        mf = s->mass[ka] / s->mass_reac[i] * s->mass[la] / s->mass_reac[i];
                        for (m = 0; m < 6; m++) dm[m] = d[m] * mf;
                        add_hess_block(d2s0, n, ka, la, +1.0, dm, fterms);
                    }
                    for (l = 0; l < s->n_reac[j]; l++) {
                        la = s->at_reac[j][l];
                        mf = s->mass[ka] / s->mass_reac[i] * s->mass[la] / s->mass_reac[j];
                        for (m = 0; m < 6; m++) dm[m] = d[m] * mf;
                        add_hess_block(d2s0, n, ka, la, -1.0, dm, fterms);
                    }
                }
                for (k = 0; k < s->n_reac[j]; k++) {
                    ka = s->at_reac[j][k];
                    for (l = 0; l < s->n_reac[i]; l++) {
                        la = s->at_reac[i][l];
                        mf = s->mass[ka] / s->mass_reac[j] * s->mass[la] / s->mass_reac[i];
                        for (m = 0; m < 6; m++) dm[m] = d[m] * mf;
                        add_hess_block(d2s0, n, ka, la, -1.0, dm, fterms);
                    }
                    for (l = 0; l < s->n_reac[j]; l++) {
                        la = s->at_reac[j][l];
                        mf = s->mass[ka] / s->mass_reac[j] * s->mass[la] / s->mass_reac[j];
                        for (m = 0; m < 6; m++) dm[m] = d[m] * mf;
                        add_hess_block(d2s0, n, ka, la, +1.0, dm, fterms);
                    }
                }
            }
        if (mode == 1) {
            int i1, j1, i2, j2;
            for (i1 = 0; i1 < 3; i1++)
                for (j1 = 0; j1 < n; j1++)
                    for (i2 = 0; i2 < 3; i2++)
                        for (j2 = 0; j2 < n; j2++) {
                            H4(d2xi_act, n, i1, j1, i2, j2) =
                                ((s0 * H4(d2s1, n, i1, j1, i2, j2) +
                                  D2(ds0, i2, j2) * D2(ds1, i1, j1) -
                                  D2(ds1, i2, j2) * D2(ds0, i1, j1) -
                                  s1 * H4(d2s0, n, i1, j1, i2, j2)) *
                                     (s0 - s1) -
                                 2.0 * (s0 * D2(ds1, i1, j1) - s1 * D2(ds0, i1, j1)) *
                                     (D2(ds0, i2, j2) - D2(ds1, i2, j2))) /
                                ((s0 - s1) * (s0 - s1) * (s0 - s1));
                        }
        } else {
            size_t t;
            for (t = 0; t < hs; t++) d2xi_act[t] = xi_ideal * d2s1[t] + (1 - xi_ideal) * d2s0[t];
        }
        free(d2s0);
        free(d2s1);
    }
    free(ds0);
    free(ds1);
}

/* umbrella.f90:66-175.  mode 0: bias + hams term; 1: xi (recrossing form) only; 3: xi only */
void orc_umbrella(orc_sys *s, const double *centroid, double xi_ideal, double *xi_real,
                  double *dxi_act, double *grad_xyz, int mode)
{
    const int n = s->natoms;
    double *d2xi;
    double delta, fs2, coeff1, coeff2, dhams;
    int i, j, k, i2, j2;
    if (mode == 1) {
        orc_calc_xi(s, centroid, xi_ideal, xi_real, dxi_act, (double *)0, 2);
        return;
    }
    if (mode == 3) {
        orc_calc_xi(s, centroid, xi_ideal, xi_real, dxi_act, (double *)0, 1);
        return;
    }
    d2xi = (double *)calloc((size_t)9 * n * n, sizeof(double));
    orc_calc_xi(s, centroid, xi_ideal, xi_real, dxi_act, d2xi, 1);
    delta = (*xi_real - xi_ideal);
    for (k = 0; k < s->nbeads; k++)
        for (j = 0; j < n; j++)
            for (i = 0; i < 3; i++)
                D3(grad_xyz, n, i, j, k) =
                    D3(grad_xyz, n, i, j, k) + s->k_force * delta * D2(dxi_act, i, j);
    fs2 = 0.0;
    for (i = 0; i < 3; i++)
        for (j = 0; j < n; j++) fs2 = fs2 + D2(dxi_act, i, j) * D2(dxi_act, i, j) / s->mass[j];
    coeff1 = 2.0 * ORC_PI_UMBR * s->beta;
    fs2 = fs2 / coeff1;
    coeff2 = -1.0 / s->beta;
    for (i = 0; i < 3; i++)
        for (j = 0; j < n; j++) {
            dhams = 0.0;
            for (i2 = 0; i2 < 3; i2++)
                for (j2 = 0; j2 < n; j2++)
                    dhams = dhams + H4(d2xi, n, i2, j2, i, j) * D2(dxi_act, i2, j2) / s->mass[j2];
            dhams = dhams * coeff2 / (coeff1 * fs2);
            for (k = 0; k < s->nbeads; k++)
                D3(grad_xyz, n, i, j, k) = D3(grad_xyz, n, i, j, k) + dhams;
        }
    free(d2xi);
}

/* constrain_q.f90:30-112; returns const_good (0 ok, 1 failed) */
int orc_constrain_q(orc_sys *s, const double *centroid, double xi_ideal, const double *dxi_act,
                    double dt)
{
    const int n = s->natoms;
    double *qtemp = (double *)calloc((size_t)3 * n, sizeof(double));
    double *dxi_new = (double *)calloc((size_t)3 * n, sizeof(double));
    double mult = 0.0, coeff = 0.0, sigma, dsigma, dx, xi_new;
    int iter, i, j, k, maxiters = 200, const_good = 0;
    for (iter = 1; iter <= maxiters; iter++) {
        coeff = mult * dt * dt / s->nbeads;
        for (i = 0; i < 3; i++)
            for (j = 0; j < n; j++)
                D2(qtemp, i, j) = D2(centroid, i, j) + coeff * D2(dxi_act, i, j) / s->mass[j];
        orc_calc_xi(s, qtemp, xi_ideal, &xi_new, dxi_new, (double *)0, 2);
        sigma = xi_new;
        dsigma = 0.0;
        for (i = 0; i < 3; i++)
            for (j = 0; j < n; j++)
                dsigma = dsigma + D2(dxi_new, i, j) * dt * dt * D2(dxi_act, i, j) /
                                      (s->mass[j] * s->nbeads);
        dx = sigma / dsigma;
        mult = mult - dx;
        /* 1.0E-8 / 1.0E-10 are REAL*4 literals (constrain_q.f90:93) */
        if (fabs(dx) < F(1.0E-8) || fabs(sigma) < F(1.0E-10)) break;
        if (iter == maxiters) {
            const_good = 1;
            free(qtemp);
            free(dxi_new);
            return const_good;
        }
    }
    for (i = 0; i < 3; i++)
        for (j = 0; j < n; j++)
            for (k = 0; k < s->nbeads; k++) {
                Q(s, i, j, k) = Q(s, i, j, k) + coeff / s->mass[j] * D2(dxi_act, i, j);
                P(s, i, j, k) = P(s, i, j, k) + mult * dt / s->nbeads * D2(dxi_act, i, j);
            }
    free(qtemp);
    free(dxi_new);
    return const_good;
}

/* constrain_p.f90:30-75 */
void orc_constrain_p(orc_sys *s, const double *dxi_act)
{
    const int n = s->natoms;
    double coeff1 = 0.0, coeff2 = 0.0, lambdas;
    int i, j, k;
    for (i = 0; i < 3; i++)
        for (j = 0; j < n; j++)
            for (k = 0; k < s->nbeads; k++)
                coeff1 = coeff1 + D2(dxi_act, i, j) * P(s, i, j, k) / s->mass[j];
    for (i = 0; i < 3; i++)
        for (j = 0; j < n; j++)
            coeff2 = coeff2 + D2(dxi_act, i, j) * D2(dxi_act, i, j) / s->mass[j];
    lambdas = -coeff1 / coeff2 / s->nbeads;
    for (i = 0; i < 3; i++)
        for (j = 0; j < n; j++)
            for (k = 0; k < s->nbeads; k++)
                P(s, i, j, k) = P(s, i, j, k) + lambdas * D2(dxi_act, i, j);
}

/* next standard normal for (i=xyz, j=atom, k=bead) of the current draw event */
static double orc_next_normal(orc_sys *s, int i, int j, int k)
{
    if (s->inject) {
        /* injected stream is consumed in the reference's loop order xyz -> atom -> bead */
        double z = (s->inject_pos < s->inject_len) ? s->inject[s->inject_pos] : 0.0;
        s->inject_pos++;
        return z;
    } else {
        double z[2];
        uint32_t m = (uint32_t)(j * 3 + i);
        oracle_rng_normal_pair(s->seed, s->traj, s->event, (uint32_t)k, m >> 1, z);
        return z[m & 1];
    }
}

/* andersen.f90:36-74: full resample, loop order xyz -> atom -> bead */
void orc_andersen(orc_sys *s)
{
    double beta_n = s->beta / s->nbeads;
    int i, j, k;
    for (i = 0; i < 3; i++)
        for (j = 0; j < s->natoms; j++) {
            double dp = sqrt(s->mass[j] / beta_n);
            for (k = 0; k < s->nbeads; k++) P(s, i, j, k) = orc_next_normal(s, i, j, k) * dp;
        }
    s->event++;
}

/* nhc.f90:34-170 */
void orc_nhc(orc_sys *s, double dt)
{
    /* ekt=1.380649E-23/4.3597447E-18*kelvin: REAL*4 literals, single-precision division */
    const float ekt_f = 1.380649E-23f / 4.3597447E-18f;
    double ekt = (double)ekt_f * s->kelvin;
    int nc = 5, ns = 3, i, j, k;
    double dtc = dt / (double)nc, w[3], scale = 1.0, eksum = 0.0;
    double dts, dt2, dt4, dt8, expterm;
    double *vnh = s->vnh, *qnh = s->qnh, *gnh = s->gnh;
    w[0] = 1.0 / (2.0 - pow(2.0, 1.0 / 3.0));
    w[1] = 1.0 - 2.0 * w[0];
    w[2] = w[0];
    for (i = 0; i < s->natoms; i++)
        if (s->at_move[i])
            for (j = 0; j < s->nbeads; j++) {
                double dot = P(s, 0, i, j) * P(s, 0, i, j) + P(s, 1, i, j) * P(s, 1, i, j) +
                             P(s, 2, i, j) * P(s, 2, i, j);
                eksum = eksum + dot / (2.0 * s->mass[i]) / s->nbeads / s->nbeads;
            }
    for (i = 1; i <= nc; i++)
        for (j = 0; j < ns; j++) {
            dts = w[j] * dtc;
            dt2 = 0.5 * dts;
            dt4 = 0.25 * dts;
            dt8 = 0.125 * dts;
            gnh[3] = (qnh[2] * vnh[2] * vnh[2] - ekt) / qnh[3];
            vnh[3] = vnh[3] + gnh[3] * dt4;
            gnh[2] = (qnh[1] * vnh[1] * vnh[1] - ekt) / qnh[2];
            expterm = exp(-vnh[3] * dt8);
            vnh[2] = expterm * (vnh[2] * expterm + gnh[2] * dt4);
            gnh[1] = (qnh[0] * vnh[0] * vnh[0] - ekt) / qnh[1];
            expterm = exp(-vnh[2] * dt8);
            vnh[1] = expterm * (vnh[1] * expterm + gnh[1] * dt4);
            gnh[0] = (2.0 * eksum - (double)s->nfree * ekt) / qnh[0];
            expterm = exp(-vnh[1] * dt8);
            vnh[0] = expterm * (vnh[0] * expterm + gnh[0] * dt4);
            expterm = exp(-vnh[0] * dt2);
            scale = scale * expterm;
            eksum = eksum * expterm * expterm;
            gnh[0] = (2.0 * eksum - (double)s->nfree * ekt) / qnh[0];
            expterm = exp(-vnh[1] * dt8);
            vnh[0] = expterm * (vnh[0] * expterm + gnh[0] * dt4);
            gnh[1] = (qnh[0] * vnh[0] * vnh[0] - ekt) / qnh[1];
            expterm = exp(-vnh[2] * dt8);
            vnh[1] = expterm * (vnh[1] * expterm + gnh[1] * dt4);
            gnh[2] = (qnh[1] * vnh[1] * vnh[1] - ekt) / qnh[2];
            expterm = exp(-vnh[3] * dt8);
            vnh[2] = expterm * (vnh[2] * expterm + gnh[2] * dt4);
            gnh[3] = (qnh[2] * vnh[2] * vnh[2] - ekt) / qnh[3];
            vnh[3] = vnh[3] + gnh[3] * dt4;
        }
    for (i = 0; i < s->natoms; i++)
        for (j = 0; j < s->nbeads; j++)
            for (k = 0; k < 3; k++) {
                if (s->at_move[i])
                    P(s, k, i, j) = scale * P(s, k, i, j);
                else
                    P(s, k, i, j) = 0.0;
            }
}

/* invert.f90:38-123 (Gauss-Jordan with full pivoting), n=3; returns 1 if singular */
static int orc_invert3(double a[3][3])
{
    int ipivot[3] = {0, 0, 0}, indxr[3], indxc[3];
    int i, j, k, irow = 0, icol = 0, n = 3;
    double big, temp, pivot;
    for (i = 0; i < n; i++) {
        big = 0.0;
        for (j = 0; j < n; j++)
            if (ipivot[j] != 1)
                for (k = 0; k < n; k++) {
                    if (ipivot[k] == 0) {
                        if (fabs(a[j][k]) >= big) {
                            big = fabs(a[j][k]);
                            irow = j;
                            icol = k;
                        }
                    } else if (ipivot[k] > 1)
                        return 1;
                }
        ipivot[icol] = ipivot[icol] + 1;
        if (irow != icol)
            for (j = 0; j < n; j++) {
                temp = a[irow][j];
                a[irow][j] = a[icol][j];
                a[icol][j] = temp;
            }
        indxr[i] = irow;
        indxc[i] = icol;
        if (a[icol][icol] == 0.0) return 1;
        pivot = a[icol][icol];
        a[icol][icol] = 1.0;
        for (j = 0; j < n; j++) a[icol][j] = a[icol][j] / pivot;
        for (j = 0; j < n; j++)
            if (j != icol) {
                temp = a[j][icol];
                a[j][icol] = 0.0;
                for (k = 0; k < n; k++) a[j][k] = a[j][k] - a[icol][k] * temp;
            }
    }
    for (i = n - 1; i >= 0; i--)
        if (indxr[i] != indxc[i])
            for (k = 0; k < n; k++) {
                temp = a[k][indxr[i]];
                a[k][indxr[i]] = a[k][indxc[i]];
                a[k][indxc[i]] = temp;
            }
    return 0;
}

/* transrot.f90:36-236; returns 1 if the inertia tensor is singular (reference: fatal) */
int orc_transrot(orc_sys *s)
{
    const int n = s->natoms, nb = s->nbeads;
    double *vel = (double *)malloc(sizeof(double) * 3 * n * nb);
    double totmass = 0.0, vtot[3] = {0, 0, 0}, weigh, xtot = 0, ytot = 0, ztot = 0;
    double mang[3] = {0, 0, 0}, vang[3], tensor[3][3];
    double xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0, xdel, ydel, zdel;
    int i, j, k, sing;
    for (i = 0; i < n; i++)
        for (k = 0; k < nb; k++)
            for (j = 0; j < 3; j++) D3(vel, n, j, i, k) = P(s, j, i, k) / s->mass[i];
    for (i = 0; i < n; i++) {
        weigh = s->mass[i];
        for (k = 0; k < nb; k++) {
            totmass = totmass + weigh;
            for (j = 0; j < 3; j++) vtot[j] = vtot[j] + D3(vel, n, j, i, k) * weigh;
        }
    }
    totmass = totmass * nb; /* transrot.f90:79 -- the reference's double count (F9) */
    for (j = 0; j < 3; j++) vtot[j] = vtot[j] / totmass;
    for (i = 0; i < n; i++) {
        weigh = s->mass[i];
        for (j = 0; j < nb; j++) {
            xtot = xtot + Q(s, 0, i, j) * weigh;
            ytot = ytot + Q(s, 1, i, j) * weigh;
            ztot = ztot + Q(s, 2, i, j) * weigh;
        }
    }
    xtot = xtot / totmass;
    ytot = ytot / totmass;
    ztot = ztot / totmass;
    for (i = 0; i < n; i++) {
        weigh = s->mass[i];
        for (k = 0; k < nb; k++) {
            mang[0] = mang[0] + (Q(s, 1, i, k) * D3(vel, n, 2, i, k) -
                                 Q(s, 2, i, k) * D3(vel, n, 1, i, k)) * weigh;
            mang[1] = mang[1] + (Q(s, 2, i, k) * D3(vel, n, 0, i, k) -
                                 Q(s, 0, i, k) * D3(vel, n, 2, i, k)) * weigh;
            mang[2] = mang[2] + (Q(s, 0, i, k) * D3(vel, n, 1, i, k) -
                                 Q(s, 1, i, k) * D3(vel, n, 0, i, k)) * weigh;
        }
    }
    mang[0] = mang[0] - (ytot * vtot[2] - ztot * vtot[1]) * totmass;
    mang[1] = mang[1] - (ztot * vtot[0] - xtot * vtot[2]) * totmass;
    mang[2] = mang[2] - (xtot * vtot[1] - ytot * vtot[0]) * totmass;
    for (i = 0; i < n; i++) {
        weigh = s->mass[i];
        for (k = 0; k < nb; k++) {
            xdel = Q(s, 0, i, k) - xtot;
            ydel = Q(s, 1, i, k) - ytot;
            zdel = Q(s, 2, i, k) - ztot;
            xx = xx + xdel * xdel * weigh;
            xy = xy + xdel * ydel * weigh;
            xz = xz + xdel * zdel * weigh;
            yy = yy + ydel * ydel * weigh;
            yz = yz + ydel * zdel * weigh;
            zz = zz + zdel * zdel * weigh;
        }
    }
    tensor[0][0] = yy + zz;
    tensor[1][0] = -xy;
    tensor[2][0] = -xz;
    tensor[0][1] = -xy;
    tensor[1][1] = xx + zz;
    tensor[2][1] = -yz;
    tensor[0][2] = -xz;
    tensor[1][2] = -yz;
    tensor[2][2] = xx + yy;
    if (n <= 2) {
        double eps = 0.000001;
        tensor[0][0] += eps;
        tensor[1][1] += eps;
        tensor[2][2] += eps;
    }
    sing = orc_invert3(tensor);
    if (sing) {
        free(vel);
        return 1;
    }
    for (i = 0; i < 3; i++) {
        vang[i] = 0.0;
        for (j = 0; j < 3; j++) vang[i] = vang[i] + tensor[i][j] * mang[j];
    }
    for (i = 0; i < n; i++)
        for (k = 0; k < nb; k++)
            for (j = 0; j < 3; j++) D3(vel, n, j, i, k) = D3(vel, n, j, i, k) - vtot[j];
    for (i = 0; i < n; i++)
        for (k = 0; k < nb; k++) {
            xdel = Q(s, 0, i, k) - xtot;
            ydel = Q(s, 1, i, k) - ytot;
            zdel = Q(s, 2, i, k) - ztot;
            D3(vel, n, 0, i, k) = D3(vel, n, 0, i, k) - vang[1] * zdel + vang[2] * ydel;
            D3(vel, n, 1, i, k) = D3(vel, n, 1, i, k) - vang[2] * xdel + vang[0] * zdel;
            D3(vel, n, 2, i, k) = D3(vel, n, 2, i, k) - vang[0] * ydel + vang[1] * xdel;
        }
    for (i = 0; i < n; i++)
        for (k = 0; k < nb; k++)
            for (j = 0; j < 3; j++) {
                P(s, j, i, k) = D3(vel, n, j, i, k) * s->mass[i];
                if (!s->at_move[i]) P(s, j, i, k) = 0.0;
            }
    free(vel);
    return 0;
}

static void orc_mask(orc_sys *s)
{
    int i, j, k;
    for (j = 0; j < s->natoms; j++)
        if (!s->at_move[j])
            for (k = 0; k < s->nbeads; k++)
                for (i = 0; i < 3; i++) P(s, i, j, k) = 0.0;
}

/* mdinit.f90:40-172 (bias_mode 1 -> umbrella mode 1, 2 -> mode 0; 0 -> no umbrella call,
 * which is what dynamic.x's constrain=-1 path amounts to for the graded configs) */
void orc_mdinit(orc_sys *s, double *derivs, double xi_ideal, double *dxi_act, int bias_mode)
{
    const int n = s->natoms;
    double *centroid = (double *)malloc(sizeof(double) * 3 * n);
    double epot, xi_real;
    int i, j;
    for (i = 0; i < s->nbeads; i++)
        orc_gradient(s, s->q + (size_t)i * 3 * n, &epot, derivs + (size_t)i * 3 * n);
    orc_get_centroid(s, centroid);
    if (bias_mode == 1)
        orc_umbrella(s, centroid, xi_ideal, &xi_real, dxi_act, derivs, 1);
    else if (bias_mode == 2)
        orc_umbrella(s, centroid, xi_ideal, &xi_real, dxi_act, derivs, 0);
    if (!s->nve) {
        if (s->thermostat == 0 || s->thermostat == 1) orc_andersen(s);
    } else {
        memset(s->p, 0, sizeof(double) * 3 * n * s->nbeads);
    }
    if (s->thermostat == 2) {
        const double k_B = 0.316679e-5; /* mdinit.f90:62 */
        double ekt, qterm;
        orc_andersen(s);
        s->nfree = 0;
        for (i = 0; i < n; i++)
            if (s->at_move[i]) s->nfree += 3;
        ekt = k_B * s->kelvin;
        qterm = ekt * s->nose_q * s->nose_q;
        for (j = 0; j < 4; j++) {
            s->qnh[j] = qterm;
            s->vnh[j] = 0.0;
            s->gnh[j] = 0.0;
        }
        s->qnh[0] = (double)s->nfree * s->qnh[0];
    }
    free(centroid);
}

/* verlet.f90:65-1308, graded paths.  Returns status bits: 1 SHAKE failed (epot carries the
 * 1e5 penalty), 2 NaN/Inf coordinate (reference: fatal), 4 singular inertia tensor (fatal),
 * 64 periodic wrap gave up (fatal); 8 / 32 from orc_rpmd_check when it is switched on. */
int orc_verlet(orc_sys *s, int istep, double *derivs, double *epot, double xi_ideal,
               double *xi_real, double *dxi_act, int constrain)
{
    const int n = s->natoms, nb = s->nbeads;
    const double dt = s->dt;
    double *centroid = (double *)malloc(sizeof(double) * 3 * n);
    double costab[ORC_MAXBEADS];
    double poly[4][ORC_MAXBEADS];
    int i, j, k, const_good = 0, status = 0;
    size_t t, tot = (size_t)3 * n * nb;

    /* 1: NHC half step */
    if (constrain != 2 && s->thermostat == 2) {
        orc_get_centroid(s, centroid);
        orc_nhc(s, dt);
    }
    /* 2: half kick ; 3: mask */
    for (t = 0; t < tot; t++) s->p[t] = s->p[t] - 0.5 * dt * derivs[t];
    orc_mask(s);
    /* 4: position update */
    if (nb == 1) {
        for (i = 0; i < 3; i++)
            for (j = 0; j < n; j++) Q(s, i, j, 0) = Q(s, i, j, 0) + P(s, i, j, 0) * dt / s->mass[j];
    } else {
        for (k = 0; k < nb; k++) costab[k] = cos(2.0 * ORC_PI_QMDFF * k / nb);
        for (i = 0; i < 3; i++)
            for (j = 0; j < n; j++) {
                orc_rfft(&P(s, i, j, 0), 3 * n, nb, costab);
                orc_rfft(&Q(s, i, j, 0), 3 * n, nb, costab);
            }
        for (j = 0; j < n; j++) {
            double beta_n, twown, pi_n;
            poly[0][0] = 1.0;
            poly[1][0] = 0.0;
            poly[2][0] = dt / s->mass[j];
            poly[3][0] = 1.0;
            beta_n = s->beta / nb;
            twown = 2.0 / beta_n;
            pi_n = ORC_PI_QMDFF / nb;
            for (k = 1; k <= nb / 2; k++) {
                double wk = twown * sin(k * pi_n);
                double wt = wk * dt;
                double wm = wk * s->mass[j];
                double cos_wt = cos(wt), sin_wt = sin(wt);
                poly[0][k] = cos_wt;
                poly[1][k] = -wm * sin_wt;
                poly[2][k] = sin_wt / wm;
                poly[3][k] = cos_wt;
            }
            for (k = 1; k <= (nb - 1) / 2; k++) {
                poly[0][nb - k] = poly[0][k];
                poly[1][nb - k] = poly[1][k];
                poly[2][nb - k] = poly[2][k];
                poly[3][nb - k] = poly[3][k];
            }
            for (k = 0; k < nb; k++)
                for (i = 0; i < 3; i++) {
                    double p_new = P(s, i, j, k) * poly[0][k] + Q(s, i, j, k) * poly[1][k];
                    Q(s, i, j, k) = P(s, i, j, k) * poly[2][k] + Q(s, i, j, k) * poly[3][k];
                    P(s, i, j, k) = p_new;
                }
        }
        for (i = 0; i < 3; i++)
            for (j = 0; j < n; j++) {
                orc_rfft(&P(s, i, j, 0), 3 * n, nb, costab);
                orc_rfft(&Q(s, i, j, 0), 3 * n, nb, costab);
            }
    }
    /* 5: periodic wrap, plain-box branch (verlet.f90:591-641): beads outer, atoms, xyz inner; every
     * shift moves ALL beads of that (atom, xyz); `tries` counts lower and upper shifts together and
     * more than 100 is `fatal` in the reference (here: status bit 64, the trajectory carries on) */
    if (s->periodic) {
        for (i = 0; i < nb; i++)
            for (j = 0; j < n; j++)
                for (k = 0; k < 3; k++) {
                    const double boxlen = s->box[k];
                    int tries = 0, b;
                    while (Q(s, k, j, i) < 0 && tries <= 100) {
                        for (b = 0; b < nb; b++) Q(s, k, j, b) = Q(s, k, j, b) + boxlen;
                        tries = tries + 1;
                    }
                    while (Q(s, k, j, i) > boxlen && tries <= 100) {
                        for (b = 0; b < nb; b++) Q(s, k, j, b) = Q(s, k, j, b) - boxlen;
                        tries = tries + 1;
                    }
                    if (tries > 100) status |= 64;
                }
    }
    /* 6: centroid ; 7: mask */
    orc_get_centroid(s, centroid);
    orc_mask(s);
    /* 9: SHAKE with dxi_act from the previous step */
    if (constrain == 1) const_good = orc_constrain_q(s, centroid, xi_ideal, dxi_act, dt);
    if (const_good == 0)
        *epot = 0.0;
    else {
        *epot = 100000.0;
        status |= 1;
    }
    /* 10: per-bead gradient */
    for (i = 0; i < nb; i++) {
        double epot1;
        orc_gradient(s, s->q + (size_t)i * 3 * n, &epot1, derivs + (size_t)i * 3 * n);
        *epot = *epot + epot1;
    }
    /* 12: umbrella (centroid is the one from step 6, i.e. before SHAKE moved the beads) */
    if (constrain == 0 || constrain == 3)
        orc_umbrella(s, centroid, xi_ideal, xi_real, dxi_act, derivs, 0);
    else if (constrain == 1 || constrain == 2)
        orc_umbrella(s, centroid, xi_ideal, xi_real, dxi_act, derivs, 1);
    /* 13: half kick + mask */
    for (t = 0; t < tot; t++) s->p[t] = s->p[t] - 0.5 * dt * derivs[t];
    orc_mask(s);
    /* 14: RATTLE */
    if (constrain == 1) orc_constrain_p(s, dxi_act);
    /* 15: NHC */
    if (constrain != 2 && s->thermostat == 2) orc_nhc(s, dt);
    /* 16: Andersen */
    if (constrain != 2 && s->thermostat == 1 && s->andersen_step > 0 &&
        (istep % s->andersen_step) == 0)
        orc_andersen(s);
    /* 18: NaN / Inf check */
    for (t = 0; t < tot; t++)
        if (s->q[t] != s->q[t] || s->q[t] > 1.79769313486231570815e308) status |= 2;
    /* 19: transrot */
    if (constrain <= 0)
        if (orc_transrot(s)) status |= 4;
    /* the drivers call rpmd_check right after verlet in the biased / constrained phases
     * (calc_rate.f90:945,1072,1575,1633 with the window's xi; recross.f90:275,476 with xi_ideal twice) */
    if (s->chk_on && constrain >= 0 && constrain != 2)
        status |= orc_rpmd_check(s, *epot, xi_ideal, (constrain == 1) ? xi_ideal : *xi_real);
    free(centroid);
    return status;
}

/* rpmd_check.f90:69-112 without its restart bookkeeping: bit 2 NaN / Inf structure or energy (:76-95),
 * bit 8 act_energy > (ts_energy+energy_tol)*nbeads (:100-106), bit 32 abs(xi_real-xi_ideal) > xi_tol (:112-116) */
int orc_rpmd_check(const orc_sys *s, double act_energy, double xi_ideal, double xi_real)
{
    const double infinity = 1.79769313486231570815e308;
    size_t t, tot = (size_t)3 * s->natoms * s->nbeads;
    int err = 0;
    for (t = 0; t < tot; t++)
        if (s->q[t] != s->q[t] || s->q[t] > infinity) err |= 2;
    if (act_energy != act_energy || act_energy > infinity) err |= 2;
    if (act_energy > (s->chk_energy_ts + s->chk_energy_tol) * s->nbeads) err |= 8;
    if (fabs(xi_real - xi_ideal) > s->chk_xi_tol) err |= 32;
    return err;
}

void oracle_sys_set_rpmd_check(orc_sys *s, int on, double energy_ts, double energy_tol, double xi_tol)
{
    s->chk_on = on;
    s->chk_energy_ts = energy_ts;
    s->chk_energy_tol = energy_tol;
    s->chk_xi_tol = xi_tol;
}

/* pbc_mod: periodic, boxlen_x/y/z (bohr) */
void oracle_sys_set_box(orc_sys *s, int periodic, const double *box)
{
    s->periodic = periodic;
    s->box[0] = box[0];
    s->box[1] = box[1];
    s->box[2] = box[2];
}

/* recross_serial.f90:172-229: one +/- child pair started from q_save (=current s->q).
 * num[child_evol] and *denom are ACCUMULATED into (caller zeroes them). */
int orc_recross_pair(orc_sys *s, double xi_ideal, int child_evol, double *num, double *denom)
{
    const int n = s->natoms, nb = s->nbeads;
    size_t tot = (size_t)3 * n * nb, t;
    double *q_save = (double *)malloc(sizeof(double) * tot);
    double *p_save = (double *)malloc(sizeof(double) * tot);
    double *derivs = (double *)malloc(sizeof(double) * tot);
    double *act = (double *)malloc(sizeof(double) * 3 * n);
    double *dxi = (double *)malloc(sizeof(double) * 3 * n);
    double xi_real, vs, fs, vpot, epot;
    int k, l, m, o, status = 0, st;
    memcpy(q_save, s->q, sizeof(double) * tot);
    orc_andersen(s);
    memcpy(p_save, s->p, sizeof(double) * tot);
    for (k = 1; k <= 2; k++) {
        for (t = 0; t < tot; t++) s->p[t] = (k == 1) ? p_save[t] : -p_save[t];
        memcpy(s->q, q_save, sizeof(double) * tot);
        orc_get_centroid(s, act);
        orc_calc_xi(s, act, xi_ideal, &xi_real, dxi, (double *)0, 2);
        for (l = 0; l < nb; l++)
            orc_gradient(s, s->q + (size_t)l * 3 * n, &vpot, derivs + (size_t)l * 3 * n);
        vs = 0.0;
        for (l = 0; l < 3; l++)
            for (m = 0; m < n; m++)
                for (o = 0; o < nb; o++) vs = vs + D2(dxi, l, m) * P(s, l, m, o) / s->mass[m];
        vs = vs / nb;
        fs = 0.0;
        for (l = 0; l < 3; l++)
            for (m = 0; m < n; m++) fs = fs + D2(dxi, l, m) * D2(dxi, l, m) / s->mass[m];
        fs = sqrt(fs / (2.0 * ORC_PI_UMBR * s->beta));
        if (vs > 0) *denom = *denom + vs / fs;
        for (l = 1; l <= child_evol; l++) {
            st = orc_verlet(s, l, derivs, &epot, xi_ideal, &xi_real, dxi, 2);
            status |= st;
            if (xi_real > 0) num[l - 1] = num[l - 1] + vs / fs;
        }
    }
    memcpy(s->q, q_save, sizeof(double) * tot);
    free(q_save);
    free(p_save);
    free(derivs);
    free(act);
    free(dxi);
    return status;
}
