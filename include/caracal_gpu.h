/*
 * caracal_gpu.h -- C-ABI of the B200 (sm_100a) RPMD hot path for Trebonius91/Caracal.
 *
 * Drop-in boundary for ONE path of the reference: propagation of a batch of ring-polymer
 * trajectories together with the per-bead PES gradient.  Plain C, plain pointers and
 * sizes; bind from Fortran with iso_c_binding exactly the way the reference binds its only
 * other native component (src/inter_mace.f90:34-67 -> src/C_API/wrap_mace.c); the module
 * a maintainer would add is fortran/caracal_gpu_mod.f90.
 *
 * Conventions (the reference's own):
 *   - arrays are Fortran column-major X(3,natoms,nbeads[,ntraj]) == C [traj][bead][atom][xyz];
 *   - bohr, hartree, electron masses, atomic time units (dt = fs / 2.41888428E-2);
 *   - atom indices passed in mechanism tables are 1-based, as in the key file;
 *   - every function returns 0 on success or a negative CRCL_E* code; nothing aborts the
 *     process (the reference's `call fatal`, fatal.f90, becomes a per-trajectory status).
 *   - "host" entry points take host pointers and copy H<->D inside the call on the
 *     handle's stream; "_dev" entry points take device pointers, are asynchronous on the
 *     handle's stream and move nothing over PCIe.
 *
 * There is no CPU fallback: without a CUDA device crcl_create fails with CRCL_ENODEV.
 */
#ifndef CARACAL_GPU_H
#define CARACAL_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRCL_VERSION 100

/* PES ids: pot_type strings of gradient.f90:154-207 that are implemented on the device */
#define CRCL_PES_NONE 0
#define CRCL_PES_H3 1    /* "h3"   egrad_h3.f   BKMP2 H + H2, 3 atoms                 */
#define CRCL_PES_OH3 2   /* "oh3"  egrad_oh3.f  Schatz-Elgersma OH + H2, atoms O,H,H,H */
#define CRCL_PES_CH4H 3  /* "ch4h" egrad_ch4h.f CBE CH4 + H, atoms H,C,H,H,H,H         */
#define CRCL_PES_BRH2 4  /* "brh2" egrad_brh2.f DIM-3C Br + H2, atoms H,Br,H   */
#define CRCL_PES_O3 5    /* "o3"   egrad_o3.f   O3 1 1A" PIP surface (Varga, Paukku, Truhlar 2017), atoms O,O,O */
#define CRCL_PES_CH4OH 6 /* "ch4oh" egrad_ch4oh.f Espinosa-Garcia/Corchado CH4 + OH, atoms H,C,H,H,H,O,H */
#define CRCL_PES_GEH4OH 7 /* "geh4oh" egrad_geh4oh.f GeH4 + OH, atoms H,Ge,H,H,H,O,H */
#define CRCL_PES_CH4CN 8 /* "ch4cn" egrad_ch4cn.f CH4 + CN (Espinosa-Garcia, Rangel, Suleimanov 2017), atoms H,C,H,H,H,C,N */
#define CRCL_PES_CLNH3 9 /* "clnh3" egrad_clnh3.f NH3 + Cl (Monge-Palacios, Rangel, Corchado, Espinosa-Garcia 2012), atoms H,N,H,H,Cl */
#define CRCL_PES_NH3OH 13 /* "nh3oh" egrad_nh3oh.f NH3 + OH (Monge-Palacios, Rangel, Espinosa-Garcia 2013), atoms H,N,H,H,O,H; the gradient is the
                             reference's own forward difference of the energy (POT_nh3oh :283-296) */
#define CRCL_PES_H2CO 14 /* "h2co" main_h2co.f90 egrad_h2co: H2CO fit in Morse variables, atoms C,O,H,H; the gradient is the reference's
                            own central difference (step 0.001 bohr); info / status bit 1 where the reference leaves the fit
                            (r(H-H) >= 8 bohr: hcopot needs a parameter file the reference does not ship) */
#define CRCL_PES_QMDFF 10 /* one QMDFF (gradient.f90:341-362): ff_eg + ff_nonb, tables via crcl_set_qmdff */
#define CRCL_PES_DGEVB 11 /* two QMDFFs + DG-EVB coupling (gradient.f90:365-537): crcl_set_qmdff,
                             crcl_set_qmdff2, crcl_set_dgevb */
#define CRCL_PES_WATER 12 /* flexible SPC water box, pes WATER_SPC (gradient.f90:212-213 -> egrad_water.f90): crcl_set_water */
#define CRCL_PES_HOSTCB 100 /* custom_grad / external_grad stay on the host (callback)  */

/* error codes */
#define CRCL_OK 0
#define CRCL_ENODEV (-1)   /* no CUDA device / driver                              */
#define CRCL_EINVAL (-2)   /* bad argument (sizes, ids, null pointers)             */
#define CRCL_ENOMEM (-3)   /* device allocation failed                             */
#define CRCL_ECUDA (-4)    /* CUDA runtime error (crcl_last_error gives the text)  */
#define CRCL_ENOSUP (-5)   /* combination not implemented on the device path       */
#define CRCL_ESTATE (-6)   /* call order (e.g. no mechanism set, no resident state) */

/* per-trajectory status written by the integrator (verlet.f90 / rpmd_check.f90 outcomes): bit flags, OR-ed
 * over the steps of a call; a trajectory with any of the fatal bits (all but PESWARN) is frozen from then on */
#define CRCL_TRAJ_OK 0
#define CRCL_TRAJ_SHAKE_FAIL 1 /* constrain_q.f90:95-98 const_good=1 (epot += 1e5)                      */
#define CRCL_TRAJ_NAN 2        /* verlet.f90:1256-1275 / rpmd_check.f90:76-95 NaN/Inf coordinate or energy */
#define CRCL_TRAJ_SINGULAR 4   /* invert.f90 singular inertia tensor (reference: fatal)                  */
#define CRCL_TRAJ_ENERGY 8     /* rpmd_check.f90:100-106 act_energy > (ts_energy+energy_tol)*nbeads      */
#define CRCL_TRAJ_PESWARN 16   /* PES printed a geometry warning (egrad_h3.f:1465,1471); not fatal       */
#define CRCL_TRAJ_XI_RANGE 32  /* rpmd_check.f90:112-116 abs(xi_real-xi_ideal) > xi_tol                   */
#define CRCL_TRAJ_PBC_FAIL 64  /* verlet.f90:601-640 more than 100 box shifts of one coordinate (fatal)  */
#define CRCL_TRAJ_FATAL (1 | 2 | 4 | 8 | 32 | 64)

/* bead-transform flavour */
#define CRCL_TRANSFORM_REFERENCE 0 /* rfft.f90/irfft.f90 as written: Re(DFT)/sqrt(N) both ways */
#define CRCL_TRANSFORM_EXACT 1     /* orthonormal normal-mode transform (physics)             */

/* integrator code path */
#define CRCL_PATH_AUTO 0  /* fused in-register kernels when the system fits, else split */
#define CRCL_PATH_FUSED 1 /* <= 8 atoms, device PES, power-of-two nbeads <= 128           */
#define CRCL_PATH_SPLIT 2 /* HBM-resident state, one kernel per stage, any natoms/nbeads,
                             device PES or host callback; constrain -1 / 2, thermostat 0 / 1 */

typedef struct crcl_handle_s *crcl_handle;

/* replaces custom_grad(xyz2,e_evb,g_evb) (custom_grad.f90:35) / external_grad as the
 * host-side PES plug-in: one image per call, (3,natoms) in, energy and (3,natoms) out */
typedef void (*crcl_host_grad_fn)(const double *xyz, double *e, double *g, int natoms, void *user);

/* ---- lifetime ---------------------------------------------------------------------- */

/* Makes explicit the module globals the path reads (evb_mod.f90:243-291 nbeads, beta;
 * general.f90 mass(:), at_move(:); dt as passed to verlet).  One handle per MPI rank / GPU. */
int crcl_create(crcl_handle *h, int device, int natoms, int nbeads, const double *mass,
                const int *at_move /* may be NULL = all movable */, double beta, double dt,
                int pes_id);
int crcl_destroy(crcl_handle h);
const char *crcl_last_error(crcl_handle h);
/* launch on an existing CUDA stream (cudaStream_t passed as void*, used as given: NULL is CUDA's
 * default stream, e.g. torch.cuda.current_stream().cuda_stream == 0).  Until this is called the
 * handle uses a private non-blocking stream created by crcl_create. */
int crcl_set_stream(crcl_handle h, void *cuda_stream);
int crcl_synchronize(crcl_handle h);

/* beta, dt, nbeads are mutated mid-run by the drivers (calc_rate.f90:651,1253) */
int crcl_set_beta_dt(crcl_handle h, double beta, double dt);
int crcl_set_transform(crcl_handle h, int mode);
/* Trajectories with fewer threads than components in the packed form -- one bead of anything, up to eight beads of a
 * three- or four-atom system: the start-structure chain of calc_rate.f90:651-1148 is 111 one-bead trajectories one after
 * the other, the constrained recrossing parent (recross.f90:228-330) ONE trajectory of 150 000 steps.  Batches of at most
 * max_beads trajectories x beads run with the components of a trajectory spread over the lanes of a half-warp or warp
 * (0.6-0.75 of the step latency of the packed form, which larger batches keep for its throughput).  Default 256;
 * 0 = always the packed form.  Both forms follow the same operations; sums over components differ in their order (last
 * bits).  No counterpart in the reference. */
int crcl_set_spread_max_beads(crcl_handle h, int max_beads);
int crcl_set_host_gradient_cb(crcl_handle h, crcl_host_grad_fn fn, void *user);
int crcl_set_path(crcl_handle h, int path);
/* Split path only: steps 2..nsteps of one crcl_verlet / work-unit call are replayed from a CUDA graph of the
 * step's 6-14 launches (default on; off = one launch per kernel as the first step always does).  Has no
 * counterpart in the reference: verlet.f90 is called once per step by its drivers. */
int crcl_set_graph(crcl_handle h, int on);

/* MECHA{} section, BIMOLEC family (calc_rate_read.f90:430-870, bonds_ref.f90): 1-based
 * atom pairs bond_form(form_num,2), bond_break(break_num,2) flattened row-wise; reference
 * lengths from the TS structure; reactant fragments at_reac (concatenated, n_reac each);
 * R_inf = dist_inf in bohr. */
int crcl_set_mechanism(crcl_handle h, int form_num, const int *bond_form, int break_num,
                       const int *bond_break, const double *form_ref, const double *break_ref,
                       int sum_reacs, const int *n_reac, const int *at_reac, double R_inf);

/* The other umbr_type families of calc_xi.f90 (both dividing surfaces without fragment centres of mass):
 * unimolecular CYCLOREVER / REARRANGE / DECOM_1BOND / ELIMINATION (calc_xi.f90:673-938): s1 from the TS
 * reference bond lengths as above, s0 from the reactant references form_reac / break_reac
 * (bonds_ref.f90:81-109, REACTANTS_STRUC);
 * ATOM_SHIFT (calc_xi.f90:523-672): one Cartesian coordinate of one atom, shift_coord 1..3 = x,y,z,
 * 4..6 = mean of (x,y), (x,z), (y,z) with the second pair of limits (calc_rate_read.f90:805-849; bohr). */
int crcl_set_mechanism_unimol(crcl_handle h, int form_num, const int *bond_form, int break_num,
                              const int *bond_break, const double *form_ref, const double *break_ref,
                              const double *form_reac, const double *break_reac);
int crcl_set_mechanism_atom_shift(crcl_handle h, int shift_atom, int shift_coord, double shift_lo,
                                  double shift_hi, double shift2_lo, double shift2_hi);

/* Tables of one QMDFF exactly as the reference holds them after prepare.f90 / rdsolvff.f90 /
 * setnonb.f90 / set_periodic.f90 (module qmdff, qmdff.f90:49-110; pbc_mod): the library receives
 * them, it does not rebuild them.  Index lists are 1-based as in the .qmdff file;
 * (94,94) and (n,n) arrays are in Fortran order.  All pointers are host pointers, copied by
 * crcl_set_qmdff. */
typedef struct crcl_qmdff_tables {
    int n;                /* atoms */
    const int *at;        /* at(n) atomic numbers */
    const double *q;      /* q(n) charges */
    const int *molnum;    /* molnum(n); may be NULL when nmols <= 1 */
    int nmols;
    int nbond, nangl, ntors, nhb, nnci, ldvt; /* ldvt = leading dimension of vtors (14) */
    const int *bond;      /* bond(2,nbond) */
    const double *vbond;  /* vbond(3,nbond): r0, k, a */
    const int *angl;      /* angl(3,nangl): centre first (ff_eg.f90:168-176) */
    const double *vangl;  /* vangl(2,nangl): theta0, k */
    const int *tors;      /* tors(6,ntors): i,j,k,l,nt,type (type 2 = inversion, ff_eg.f90:316) */
    const double *vtors;  /* vtors(ldvt,ntors): phi0, k, nt x (n, phase, V) */
    const int *nci;       /* nci(3,nnci): i, j, screening class 1..6 */
    const double *c6xy;   /* c6xy(n,n) */
    const double *r0ab, *zab, *r094, *sr42; /* (94,94): r0ab, zab, r094_mod, sr42 */
    const double *rad;    /* rad(94) */
    double eps1[6], eps2[6];
    int periodic, zahn;   /* pbc_mod: periodic, zahn */
    double box[3];        /* boxlen_x, boxlen_y, boxlen_z */
    double coul_cut, vdw_cut, cut_low, zahn_a, zahn_par;
    double e_zero;        /* E_zero1 (ESHIFT keyword) */
    /* ff_hb.f90: hb(3,nhb) = A,B,H ; vhb(2,nhb); scalehb_glob(94), scalexb_glob(94), q_glob(n).
     * scalehb == NULL switches the H/X-bond terms off entirely (then nhb must be 0). */
    const int *hb;
    const double *vhb;
    const double *scalehb, *scalexb, *q_glob;
} crcl_qmdff_tables;
/* handle must have been created with pes_id = CRCL_PES_QMDFF (or CRCL_PES_DGEVB) and natoms = T->n */
int crcl_set_qmdff(crcl_handle h, const crcl_qmdff_tables *T);
/* second diabatic state: the *_two table set (bond_two, vbond_two, ..., q_two, c6xy_two, E_zero2).
 * Evaluated with the semantics of ff_eg_two.f90 / ff_nonb_two.f90 / ff_hb_two.f90: never periodic,
 * list terms only, Coulomb q_i q_j eps1 / r without cut-off (periodic, nmols, zahn, cut-offs of T are
 * ignored). */
int crcl_set_qmdff2(crcl_handle h, const crcl_qmdff_tables *T);

/* DG-EVB coupling (module evb_mod; read_pes.f90:2046-2225, evb_pars.dat): dg_mode 1..3,
 * coord_def(nat6,5) flattened row-wise = type (1 dist, 2 angle, 3 dihedral, 4 out-of-plane) and up
 * to four 1-based atoms; point_int(nat6,npoints) in Fortran order; alph_opt(npoints);
 * b_vec(mat_size), mat_size = npoints * (1 | 1+nat6 | 1+nat6+nat6(nat6+1)/2); g_thres (1E-10). */
typedef struct crcl_dgevb_params {
    int mode, npoints, nat6;
    const int *coord_def;
    const double *point_int, *alph, *b_vec;
    double g_thres;
} crcl_dgevb_params;
int crcl_set_dgevb(crcl_handle h, const crcl_dgevb_params *P);

/* Flexible SPC water box (pes WATER_SPC): what water_init.f90:53-107 and set_periodic.f90:66-104 leave in the modules.
 * n = 3 nwater atoms ordered O,H,H per molecule; pars = water_pars(1:11) in atomic units (r_0, r_0HH, D_e, a, k_theta,
 * k_rtheta, k_rr, e_H, e_O, sigma_OO, eps_OO); q(n) the charges laid out by water_init.f90:103-107; is_O(n) = name(i)=="O".
 * Then crcl_egrad(..., CRCL_PES_WATER, ...) replaces egrad_water(xyz_act,g_act,e_act) (egrad_water.f90:36), and the
 * integrator runs on the HBM-resident path with it.  Host pointers, copied. */
typedef struct crcl_water_params {
    int n;
    int periodic, zahn;
    double box[3];
    double coul_cut, zahn_a, zahn_par;
    double pars[11];
    const double *q;
    const int *is_O;
} crcl_water_params;
int crcl_set_water(crcl_handle h, const crcl_water_params *P);

/* Smooth particle-mesh Ewald, reciprocal-space part (ewald_recip.f90:30-470).  The set-up quantities
 * are those of module pbc_mod after set_periodic.f90:114-231: boxlen_x/y/z (bohr), a_ewald, nfft (one
 * value for all three dimensions), bsorder (5) and the B-spline moduli bsmod1/2/3(nfft).
 * NB the reference cannot reach ewald_recip (ff_nonb.f90:337 sets ewald=.false.): this entry point
 * stands alone and is not part of CRCL_PES_QMDFF. */
typedef struct crcl_ewald_params {
    double box[3];
    double a_ewald;
    int nfft, bsorder;
    const double *bsmod1, *bsmod2, *bsmod3;
} crcl_ewald_params;
int crcl_set_ewald(crcl_handle h, const crcl_ewald_params *P);
/* ewald_recip(n,xyz,q,energy,grad) for nimg structures: xyz [nimg][n][3], q [n] -> energy [nimg],
 * grad [nimg][n][3] (overwritten, the reference zeroes grad too, :421) */
int crcl_ewald_recip(crcl_handle h, int n, int nimg, const double *xyz, const double *q,
                     double *energy, double *grad);

/* pbc_mod: periodic, boxlen_x/y/z (bohr).  Switches on the wrap of verlet.f90:591-641 (plain-box branch: all beads
 * of an atom are shifted together until every bead lies in [0, L]).  crcl_set_qmdff / crcl_set_water set it from
 * their tables; this entry point serves the host-callback PES (and switches it off again).  A periodic handle always
 * runs on the HBM-resident path (the analytic gas-phase surfaces of the fused kernels have no box). */
int crcl_set_box(crcl_handle h, int periodic, const double *boxlen /* [3] */);

/* rpmd_check (rpmd_check.f90:69-116), the guard the drivers call after every verlet step of the start-structure,
 * umbrella and recrossing-parent phases (calc_rate.f90:945,1072,1575,1633; recross.f90:275,476): when switched on,
 * every step with constrain 0, 1 or 3 tests epot > (energy_ts + energy_tol) * nbeads -> CRCL_TRAJ_ENERGY (a failed
 * SHAKE trips it through its 1e5 penalty, as in the reference), NaN energy -> CRCL_TRAJ_NAN, and for constrain 0 / 3
 * abs(xi_real - xi_ideal) > xi_tol -> CRCL_TRAJ_XI_RANGE (recross passes xi_ideal twice, so constrain 1 never trips
 * it).  The restart bookkeeping (err_count, new start structures) stays with the caller.  Units: hartree, xi. */
int crcl_set_rpmd_check(crcl_handle h, int on, double energy_ts, double energy_tol, double xi_tol);

/* NVT{} section: thermostat 0 none, 1 Andersen, 2 Nose-Hoover chain (dynamic.f90:463-465);
 * andersen_step as evb_mod.f90:289; kelvin and nose_q for nhc.f90 / mdinit.f90:138-146 */
int crcl_set_thermostat(crcl_handle h, int thermostat, int andersen_step, double kelvin,
                        double nose_q);
/* counter-based RNG (replaces random_init_local, andersen.f90:131) */
int crcl_set_seed(crcl_handle h, uint64_t seed);

/* ---- multi-GPU: the one exchange step of the path --------------------------------------------
 * One process per GPU, one handle per process.  Replaces the result traffic of the reference's MPI master/worker
 * schemes: recross.f90:390,411 (`message(child_evol+2)` from every worker to rank 0) and the statistics files of
 * calc_rate.f90:1690-1734.  Rank 0 calls crcl_comm_unique_id, the caller ships the CRCL_UNIQUE_ID_BYTES bytes to the
 * other ranks (mpi_bcast in the Fortran drivers; torch.distributed or a file in Python), every rank calls
 * crcl_comm_init (collective).  From then on crcl_recross_children(_dev) and crcl_umbrella_windows are COLLECTIVE
 * calls: every rank passes the same GLOBAL unit range, runs its contiguous block of it (blocks differ by at most one
 * unit; RNG streams are keyed by the global unit index, so results do not depend on the number of ranks beyond the
 * summation order) and receives the result of the whole job: ncclAllReduce(sum, double, child_evol + 1) of the
 * kappa(t) numerators and the denominator, and of the per-trajectory (average, variance, status) vectors.
 * NCCL is bound at run time (libnccl.so.2, or $CRCL_NCCL_LIB); without it these three calls return CRCL_ESTATE and
 * everything else works. */
#define CRCL_UNIQUE_ID_BYTES 128
int crcl_comm_unique_id(void *id_out /* CRCL_UNIQUE_ID_BYTES bytes */);
int crcl_comm_init(crcl_handle h, int nranks, int rank, const void *unique_id);
int crcl_comm_destroy(crcl_handle h);
/* any pointer may be NULL; nccl_version as ncclGetVersion reports it (0: NCCL not loadable) */
int crcl_comm_info(crcl_handle h, int *nranks, int *rank, int *nccl_version);

/* ---- PES seam: egrad_<pes>(q,Natoms,Nbeads,V,dVdq,info) --------------------------------
 * Same argument order and layout as egrad_h3.f:29 / egrad_ch4h.f:74 / egrad_oh3.f:33;
 * Nbeads generalises to nimg = ntraj*nbeads images.  info = OR of warning bits. */
int crcl_egrad(crcl_handle h, int pes_id, const double *q, int natoms, int nimg, double *V,
               double *dVdq, int *info);
int crcl_egrad_dev(crcl_handle h, int pes_id, const double *d_q, int natoms, int nimg,
                   double *d_V, double *d_dVdq, int *d_info /* one int, may be NULL */);

/* ---- integrator seam: verlet(istep,dt,derivs,epot,...,constrain,...) (verlet.f90:65) ----
 * Advances ntraj independent ring polymers by nsteps steps, istep = istep0+1..istep0+nsteps.
 * constrain: -1 plain MD (dynamic.x), 0 umbrella, 1 SHAKE/RATTLE on xi, 2 child trajectory.
 * q, p, derivs: [ntraj][nbeads][natoms][3], in/out.  xi_ideal, k_force: per trajectory
 * (k_force = k_force(um_window_act)).  Outputs per trajectory from the LAST step: epot,
 * xi_real; dxi [ntraj][natoms][3] is in/out (SHAKE uses the previous step's, verlet.f90:744).
 * status: CRCL_TRAJ_* of the first failing step (trajectory is frozen from then on).
 * traj_id: global trajectory numbers keying the RNG streams (NULL = 0..ntraj-1);
 * event0 [ntraj]: in/out count of Andersen redraws consumed (NULL = start at 0). */
int crcl_verlet(crcl_handle h, int ntraj, int nsteps, int istep0, int constrain,
                const double *xi_ideal, const double *k_force, double *q, double *p,
                double *derivs, double *epot, double *xi_real, double *dxi, int *status,
                const uint32_t *traj_id, uint32_t *event0);

/* the same on state resident in device memory: every pointer is a device pointer, the call is asynchronous on the
 * handle's stream and nothing crosses PCIe.  d_q, d_p, d_derivs, d_epot, d_xi_real, d_dxi, d_status, d_event must be
 * given; d_xi_ideal, d_k_force, d_traj_id may be NULL as above. */
int crcl_verlet_dev(crcl_handle h, int ntraj, int nsteps, int istep0, int constrain, const double *d_xi_ideal,
                    const double *d_k_force, double *d_q, double *d_p, double *d_derivs, double *d_epot,
                    double *d_xi_real, double *d_dxi, int *d_status, const uint32_t *d_traj_id, uint32_t *d_event);

/* mdinit(derivs,xi_ideal,dxi_act,bias_mode,rank) (mdinit.f90:40): gradient of all beads,
 * umbrella (bias_mode 1 -> xi only, 2 -> bias applied, 0 -> none), fresh momenta, NHC reset. */
int crcl_mdinit(crcl_handle h, int ntraj, int bias_mode, const double *xi_ideal,
                const double *k_force, const double *q, double *p, double *derivs, double *dxi,
                const uint32_t *traj_id, uint32_t *event0);

/* calc_xi(coords,xi_ideal,xi_act,dxi_act,d2xi_act,mode) (calc_xi.f90:63) on ncoord
 * structures [ncoord][natoms][3]; mode 1 umbrella form, 2 recrossing form; d2xi may be NULL */
int crcl_calc_xi(crcl_handle h, int ncoord, const double *coords, const double *xi_ideal,
                 int mode, double *xi, double *dxi, double *d2xi);

/* ---- work-unit seam --------------------------------------------------------------------
 * Recrossing children (recross.f90:515-628 worker body / recross_serial.f90:172-229):
 * pairs pair0..pair0+npairs-1; pair g starts from parent snapshot (g mod nparent), draws
 * momenta with RNG stream (seed, traj=g), runs the +p and -p child for child_evol free
 * steps.  kappa_num[child_evol] and *kappa_denom receive this call's sums (reduce across
 * ranks/GPUs by plain addition; with a communicator, crcl_comm_init, the call is collective over the GLOBAL range
 * pair0..pair0+npairs-1 and returns the sums of the whole job on every rank).  status[npairs] may be NULL. */
int crcl_recross_children(crcl_handle h, const double *q_parents, int nparent, int pair0,
                          int npairs, int child_evol, double xi_ideal, double *kappa_num,
                          double *kappa_denom, int *status);
/* same with q_parents / outputs resident in device memory (async on the stream); d_status holds 2*npairs ints,
 * one per child trajectory */
int crcl_recross_children_dev(crcl_handle h, const double *d_q_parents, int nparent, int pair0,
                              int npairs, int child_evol, double xi_ideal, double *d_kappa_num,
                              double *d_kappa_denom, int *d_status);

/* Umbrella window worker body (calc_rate.f90:1387-1700): ntraj trajectories started from
 * q0 [nbeads][natoms][3] in window (xi0, k): mdinit, equi_steps, then sample_steps while
 * accumulating xi.  avg/var [ntraj] as written to statistics/bias_<xi> (:1690-1700). */
int crcl_umbrella_window(crcl_handle h, const double *q0, double xi0, double k_force, int ntraj,
                         int equi_steps, int sample_steps, uint32_t traj_id0, double *avg,
                         double *var, int *status);
/* The master/worker loop over windows (calc_rate.f90:1351-1376) as one batch: window w starts from
 * q0[w] with (xi0[w], k_force[w]); trajectory t of window w is index w*ntraj+t in avg/var/status and
 * uses RNG stream traj_id0 + w*ntraj + t.  Before the sampling phase the forces are recomputed
 * without the bias, as calc_rate.f90:1619-1623 does.  constrain is the flag handed to verlet:
 * 0 as calc_rate.f90 does, or 3 = the same biased dynamics without the removal of net
 * translation / rotation (verlet.f90:1051,1300-1306; what pre_sample.f90:123 uses). */
int crcl_umbrella_windows(crcl_handle h, int nwin, const double *q0, const double *xi0,
                          const double *k_force, int ntraj, int equi_steps, int sample_steps,
                          int constrain, uint32_t traj_id0, double *avg, double *var, int *status);

/* ---- test / introspection hooks ---------------------------------------------------- */
/* n standard normals of stream (seed, traj, event, bead), elements 0..n-1 (DESIGN.md "RNG") */
int crcl_rng_normals(crcl_handle h, uint64_t seed, uint32_t traj, uint32_t event, uint32_t bead,
                     int n, double *out);
/* number of kernel launches issued through this handle since creation */
long long crcl_launch_count(crcl_handle h);
/* duration in ms of the most recent trajectory-kernel launch (CUDA events on the handle's
 * stream; valid after crcl_synchronize or any host-pointer call) */
double crcl_last_kernel_ms(crcl_handle h);
/* durations (ms, CUDA events on the handle's stream around each launch) of the trajectory /
 * egrad kernels launched since the previous call, oldest first, at most max_n (ring of 256);
 * synchronises the stream; returns the number written */
int crcl_kernel_timings(crcl_handle h, double *ms_out, int max_n);
/* times the split path's propagation kernel (half kick + free ring polymer + centroid) alone on
 * ntraj synthetic trajectories resident in HBM: ms_out[0] = mean, ms_out[1] = best of reps launches.
 * Algorithmic traffic is 120 B per (trajectory, bead, atom). */
int crcl_bench_propagate(crcl_handle h, int ntraj, int reps, double *ms_out);
/* sustained FP64 FMA throughput of the device in TFLOP/s (DFMA microbenchmark, used as the
 * roofline denominator because MEASURED_PEAKS.json has no FP64 entry) */
double crcl_measure_fp64_tflops(crcl_handle h, int iters);
/* same for the FP64 tensor-core path (mma.sync.m8n8k4.f64): the evidence behind keeping the bead
 * transform on the DFMA pipe */
double crcl_measure_dmma_tflops(crcl_handle h, int iters);
/* the free ring-polymer step as a stand-alone shared-memory-resident kernel in its FMA form (the loop of the trajectory
 * kernels) and in a DMMA form, on the same synthetic input (csrc/transform_bench.cuh): nbeads = 16 (6 atoms) or 64
 * (4 atoms), ntraj trajectories, the step applied reps times per launch.  out[0] = ms per launch of the FMA form,
 * out[1] = of the DMMA form (best of 3), out[2] = largest difference between the two results relative to the largest
 * |p| / |q| after three steps, out[3] = flops per launch (4 NB x NB matrix-vector products per component),
 * out[4] = largest |q' - q| (the step is not the identity).  Measurement helper, not on the reference's path. */
int crcl_bench_transform(crcl_handle h, int nbeads, int ntraj, int reps, double *out);

#ifdef __cplusplus
}
#endif
#endif
