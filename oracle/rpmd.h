/*
 * rpmd.h -- state struct + entry points of the CPU oracle's integrator part
 * (TEST INFRASTRUCTURE ONLY, see oracle.h).  The struct makes explicit the module
 * globals of the reference (evb_mod.f90:243-291 q_i,p_i,nbeads,beta,andersen_step,...;
 * general.f90 bath block vnh,qnh,gnh,nfree; mass(:), at_move(:)).
 */
#ifndef ORACLE_RPMD_H
#define ORACLE_RPMD_H
#include <stdint.h>

#define ORC_MAXBEADS 1024 /* fftw_mod.f90:6 Nmax */
#define ORC_MAXBOND 8
#define ORC_MAXREAC 4
#define ORC_MAXREACAT 200 /* calc_rate_read.f90:540 */

#define ORC_PES_NONE 0
#define ORC_PES_H3 1
#define ORC_PES_OH3 2
#define ORC_PES_CH4H 3
#define ORC_PES_BRH2 4
#define ORC_PES_O3 5
#define ORC_PES_CH4OH 6
#define ORC_PES_GEH4OH 7
#define ORC_PES_CH4CN 8
#define ORC_PES_CLNH3 9
#define ORC_PES_NH3OH 13
#define ORC_PES_H2CO 14

#ifdef __cplusplus
extern "C" {
#endif

typedef void (*orc_custom_grad_fn)(const double *xyz, double *e, double *g, int natoms);

typedef struct orc_sys {
    int natoms, nbeads, pes;
    double *mass;
    int *at_move;
    double beta, dt, kelvin;
    int thermostat; /* 0 none, 1 Andersen, 2 NHC (dynamic.f90:463-465) */
    int andersen_step, nve;
    double nose_q, vnh[4], qnh[4], gnh[4];
    int nfree;
    /* reaction-coordinate mechanism (0-based atom indices) */
    int form_num, break_num;
    int bond_form[ORC_MAXBOND][2], bond_break[ORC_MAXBOND][2];
    double form_ref[ORC_MAXBOND], break_ref[ORC_MAXBOND];
    int sum_reacs, n_reac[ORC_MAXREAC], at_reac[ORC_MAXREAC][ORC_MAXREACAT];
    double mass_reac[ORC_MAXREAC], R_inf;
    /* umbr_type family: 0 BIMOLEC family (calc_xi.f90:108-502), 1 unimolecular CYCLOREVER / REARRANGE /
     * DECOM_1BOND / ELIMINATION (:673-938), 2 ATOM_SHIFT (:523-672) */
    int umbr_type;
    double form_reac[ORC_MAXBOND], break_reac[ORC_MAXBOND]; /* bonds_ref.f90:81-109 */
    int shift_atom, shift_coord;                           /* 0-based atom, coord 1..6 as in the reference */
    double shift_lo, shift_hi, shift2_lo, shift2_hi;
    double k_force; /* k_force(um_window_act) */
    double *q, *p;  /* [bead][atom][xyz] */
    /* normal-deviate source */
    uint64_t seed;
    uint32_t traj, event;
    const double *inject;
    long inject_len, inject_pos;
    orc_custom_grad_fn custom_grad;
    /* pbc_mod: periodic, boxlen_x/y/z -- the plain-box wrap of verlet.f90:591-641 */
    int periodic;
    double box[3];
    /* rpmd_check.f90 settings: ts_energy, energy_tol (RPMD_EN_TOL), xi_tol */
    int chk_on;
    double chk_energy_ts, chk_energy_tol, chk_xi_tol;
} orc_sys;

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void oracle_rng_normal_pair(uint64_t seed, uint32_t traj, uint32_t event, uint32_t bead,
                            uint32_t pair, double z[2]);

orc_sys *oracle_sys_create(int natoms, int nbeads, const double *mass, const int *at_move,
                           double beta, double dt, int pes);
void oracle_sys_free(orc_sys *s);
void oracle_sys_set_mecha(orc_sys *s, int form_num, const int *bond_form, int break_num,
                          const int *bond_break, const double *form_ref, const double *break_ref,
                          int sum_reacs, const int *n_reac, const int *at_reac, double R_inf);
void oracle_sys_set_thermostat(orc_sys *s, int thermostat, int andersen_step, double kelvin,
                               double nose_q);
void oracle_sys_set_kforce(orc_sys *s, double k_force);
void oracle_sys_set_rng(orc_sys *s, uint64_t seed, uint32_t traj, uint32_t event);
void oracle_sys_inject_normals(orc_sys *s, const double *z, long n);
void oracle_sys_set_custom_grad(orc_sys *s, orc_custom_grad_fn fn);
double *oracle_sys_q(orc_sys *s);
double *oracle_sys_p(orc_sys *s);
void oracle_sys_get_nhc(orc_sys *s, double *v4q4);
uint32_t oracle_sys_get_event(orc_sys *s);
void oracle_sys_set_unimol(orc_sys *s, const double *form_reac, const double *break_reac);
void oracle_sys_set_atom_shift(orc_sys *s, int shift_atom, int shift_coord, double shift_lo, double shift_hi,
                               double shift2_lo, double shift2_hi);

void orc_gradient(orc_sys *s, const double *xyz, double *e, double *g);
void orc_get_centroid(orc_sys *s, double *centroid);
void orc_calc_xi(const orc_sys *s, const double *coords, double xi_ideal, double *xi_act,
                 double *dxi_act, double *d2xi_act, int mode);
void orc_umbrella(orc_sys *s, const double *centroid, double xi_ideal, double *xi_real,
                  double *dxi_act, double *grad_xyz, int mode);
int orc_constrain_q(orc_sys *s, const double *centroid, double xi_ideal, const double *dxi_act,
                    double dt);
void orc_constrain_p(orc_sys *s, const double *dxi_act);
void orc_andersen(orc_sys *s);
void orc_nhc(orc_sys *s, double dt);
int orc_transrot(orc_sys *s);
void orc_mdinit(orc_sys *s, double *derivs, double xi_ideal, double *dxi_act, int bias_mode);
int orc_verlet(orc_sys *s, int istep, double *derivs, double *epot, double xi_ideal,
               double *xi_real, double *dxi_act, int constrain);
int orc_rpmd_check(const orc_sys *s, double act_energy, double xi_ideal, double xi_real);
void oracle_sys_set_rpmd_check(orc_sys *s, int on, double energy_ts, double energy_tol, double xi_tol);
void oracle_sys_set_box(orc_sys *s, int periodic, const double *box);
int orc_recross_pair(orc_sys *s, double xi_ideal, int child_evol, double *num, double *denom);

#ifdef __cplusplus
}
#endif
#endif
