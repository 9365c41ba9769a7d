!
!     module caracal_gpu: iso_c_binding interface to libcaracal_gpu.so (include/caracal_gpu.h)
!     and thin shims that keep the call sites of the reference drivers unchanged.
!
!     Written in the style of src/inter_mace.f90:34-67 (the reference's only other bind(C)
!     interface).  NOT compiled in this repository's CI: the build image has no Fortran compiler
!     (SURVEY.md F1).  See INTEGRATION.md for where each shim is called from.
!
module caracal_gpu
use, intrinsic :: iso_c_binding
implicit none

integer(c_int), parameter :: CRCL_PES_H3 = 1, CRCL_PES_OH3 = 2, CRCL_PES_CH4H = 3, CRCL_PES_BRH2 = 4, CRCL_PES_O3 = 5, &
                             CRCL_PES_CH4OH = 6, CRCL_PES_GEH4OH = 7, CRCL_PES_CH4CN = 8, CRCL_PES_CLNH3 = 9, &
                             CRCL_PES_NH3OH = 13, CRCL_PES_H2CO = 14, CRCL_PES_WATER = 12
integer(c_int), parameter :: CRCL_PES_QMDFF = 10, CRCL_PES_DGEVB = 11, CRCL_PES_HOSTCB = 100
type(c_ptr), save :: crcl_h = c_null_ptr        ! one handle per MPI rank / GPU
!     per-trajectory status bits (include/caracal_gpu.h): SHAKE_FAIL 1, NAN 2, SINGULAR 4, ENERGY 8, PESWARN 16,
!     XI_RANGE 32, PBC_FAIL 64; CRCL_TRAJ_FATAL = all but PESWARN
integer(c_int), parameter :: CRCL_TRAJ_SHAKE_FAIL = 1, CRCL_TRAJ_NAN = 2, CRCL_TRAJ_SINGULAR = 4, CRCL_TRAJ_ENERGY = 8, &
                             CRCL_TRAJ_PESWARN = 16, CRCL_TRAJ_XI_RANGE = 32, CRCL_TRAJ_PBC_FAIL = 64, CRCL_TRAJ_FATAL = 111
integer(c_int), parameter :: CRCL_UNIQUE_ID_BYTES = 128
!     RNG stream of the trajectory this rank is propagating through verlet_gpu / mdinit_gpu: the library draws
!     Philox(seed, traj, event, bead, component pair); the event counter must survive from call to call, otherwise
!     every Andersen resample repeats the first draw (the drivers call verlet one step at a time)
integer(c_int32_t), target, save :: crcl_traj(1) = 0_c_int32_t, crcl_event(1) = 0_c_int32_t

!     struct crcl_qmdff_tables of include/caracal_gpu.h (field order and types must match)
type, bind(C) :: crcl_qmdff_tables
   integer(c_int) :: n
   type(c_ptr) :: at, q, molnum
   integer(c_int) :: nmols
   integer(c_int) :: nbond, nangl, ntors, nhb, nnci, ldvt
   type(c_ptr) :: bond, vbond, angl, vangl, tors, vtors, nci, c6xy
   type(c_ptr) :: r0ab, zab, r094, sr42, rad
   real(c_double) :: eps1(6), eps2(6)
   integer(c_int) :: periodic, zahn
   real(c_double) :: box(3)
   real(c_double) :: coul_cut, vdw_cut, cut_low, zahn_a, zahn_par
   real(c_double) :: e_zero
   type(c_ptr) :: hb, vhb, scalehb, scalexb, q_glob
end type crcl_qmdff_tables

!     struct crcl_dgevb_params
type, bind(C) :: crcl_dgevb_params
   integer(c_int) :: mode, npoints, nat6
   type(c_ptr) :: coord_def, point_int, alph, b_vec
   real(c_double) :: g_thres
end type crcl_dgevb_params

!     struct crcl_water_params (pes WATER_SPC: water_init.f90:53-107, set_periodic.f90:66-104)
type, bind(C) :: crcl_water_params
   integer(c_int) :: n, periodic, zahn
   real(c_double) :: box(3), coul_cut, zahn_a, zahn_par, pars(11)
   type(c_ptr) :: q, is_O
end type crcl_water_params

!     struct crcl_ewald_params (pbc_mod after set_periodic.f90:114-231)
type, bind(C) :: crcl_ewald_params
   real(c_double) :: box(3), a_ewald
   integer(c_int) :: nfft, bsorder
   type(c_ptr) :: bsmod1, bsmod2, bsmod3
end type crcl_ewald_params

interface
   function crcl_create(h, device, natoms, nbeads, mass, at_move, beta, dt, pes_id) bind(C, name="crcl_create")
      import :: c_ptr, c_int, c_double
      type(c_ptr), intent(out) :: h
      integer(c_int), value :: device, natoms, nbeads, pes_id
      real(c_double), dimension(*), intent(in) :: mass
      integer(c_int), dimension(*), intent(in) :: at_move
      real(c_double), value :: beta, dt
      integer(c_int) :: crcl_create
   end function crcl_create

   function crcl_destroy(h) bind(C, name="crcl_destroy")
      import :: c_ptr, c_int
      type(c_ptr), value :: h
      integer(c_int) :: crcl_destroy
   end function crcl_destroy

   function crcl_set_beta_dt(h, beta, dt) bind(C, name="crcl_set_beta_dt")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      real(c_double), value :: beta, dt
      integer(c_int) :: crcl_set_beta_dt
   end function crcl_set_beta_dt

   function crcl_set_mechanism(h, form_num, bond_form, break_num, bond_break, form_ref, break_ref, &
                               sum_reacs, n_reac, at_reac, r_inf) bind(C, name="crcl_set_mechanism")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: form_num, break_num, sum_reacs
      integer(c_int), dimension(*), intent(in) :: bond_form, bond_break, n_reac, at_reac
      real(c_double), dimension(*), intent(in) :: form_ref, break_ref
      real(c_double), value :: r_inf
      integer(c_int) :: crcl_set_mechanism
   end function crcl_set_mechanism

   function crcl_set_thermostat(h, thermostat, andersen_step, kelvin, nose_q) bind(C, name="crcl_set_thermostat")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: thermostat, andersen_step
      real(c_double), value :: kelvin, nose_q
      integer(c_int) :: crcl_set_thermostat
   end function crcl_set_thermostat

   function crcl_set_seed(h, seed) bind(C, name="crcl_set_seed")
      import :: c_ptr, c_int, c_int64_t
      type(c_ptr), value :: h
      integer(c_int64_t), value :: seed
      integer(c_int) :: crcl_set_seed
   end function crcl_set_seed

   ! egrad_<pes>(q,Natoms,Nbeads,V,dVdq,info)  (egrad_h3.f:29, egrad_ch4h.f:74, egrad_oh3.f:33)
   function crcl_egrad(h, pes_id, q, natoms, nimg, V, dVdq, info) bind(C, name="crcl_egrad")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: pes_id, natoms, nimg
      real(c_double), dimension(*), intent(in) :: q
      real(c_double), dimension(*), intent(out) :: V, dVdq
      integer(c_int), intent(out) :: info
      integer(c_int) :: crcl_egrad
   end function crcl_egrad

   ! verlet (verlet.f90:65) on ntraj ring polymers
   function crcl_verlet(h, ntraj, nsteps, istep0, constrain, xi_ideal, k_force, q, p, derivs, epot, &
                        xi_real, dxi, status, traj_id, event0) bind(C, name="crcl_verlet")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: ntraj, nsteps, istep0, constrain
      real(c_double), dimension(*), intent(in) :: xi_ideal, k_force
      real(c_double), dimension(*), intent(inout) :: q, p, derivs, dxi
      real(c_double), dimension(*), intent(out) :: epot, xi_real
      integer(c_int), dimension(*), intent(inout) :: status
      type(c_ptr), value :: traj_id, event0          ! c_null_ptr or c_loc of uint32 arrays
      integer(c_int) :: crcl_verlet
   end function crcl_verlet

   ! mdinit (mdinit.f90:40)
   function crcl_mdinit(h, ntraj, bias_mode, xi_ideal, k_force, q, p, derivs, dxi, traj_id, event0) &
                        bind(C, name="crcl_mdinit")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: ntraj, bias_mode
      real(c_double), dimension(*), intent(in) :: xi_ideal, k_force, q
      real(c_double), dimension(*), intent(inout) :: p, derivs, dxi
      type(c_ptr), value :: traj_id, event0
      integer(c_int) :: crcl_mdinit
   end function crcl_mdinit

   ! recross worker body (recross.f90:515-628)
   function crcl_recross_children(h, q_parents, nparent, pair0, npairs, child_evol, xi_ideal, &
                                  kappa_num, kappa_denom, status) bind(C, name="crcl_recross_children")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      real(c_double), dimension(*), intent(in) :: q_parents
      integer(c_int), value :: nparent, pair0, npairs, child_evol
      real(c_double), value :: xi_ideal
      real(c_double), dimension(*), intent(out) :: kappa_num
      real(c_double), intent(out) :: kappa_denom
      integer(c_int), dimension(*), intent(out) :: status
      integer(c_int) :: crcl_recross_children
   end function crcl_recross_children

   ! umbrella worker body (calc_rate.f90:1387-1700)
   function crcl_umbrella_window(h, q0, xi0, k_force, ntraj, equi_steps, sample_steps, traj_id0, &
                                 avg, var, status) bind(C, name="crcl_umbrella_window")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      real(c_double), dimension(*), intent(in) :: q0
      real(c_double), value :: xi0, k_force
      integer(c_int), value :: ntraj, equi_steps, sample_steps, traj_id0
      real(c_double), dimension(*), intent(out) :: avg, var
      integer(c_int), dimension(*), intent(out) :: status
      integer(c_int) :: crcl_umbrella_window
   end function crcl_umbrella_window

   ! all windows of the umbrella phase in one batch (master/worker loop calc_rate.f90:1351-1376)
   function crcl_umbrella_windows(h, nwin, q0, xi0, k_force, ntraj, equi_steps, sample_steps, constrain, traj_id0, &
                                  avg, var, status) bind(C, name="crcl_umbrella_windows")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: nwin, ntraj, equi_steps, sample_steps, constrain, traj_id0
      real(c_double), dimension(*), intent(in) :: q0, xi0, k_force
      real(c_double), dimension(*), intent(out) :: avg, var
      integer(c_int), dimension(*), intent(out) :: status
      integer(c_int) :: crcl_umbrella_windows
   end function crcl_umbrella_windows

   ! the other umbr_type families of calc_xi.f90: unimolecular (:673-938) and ATOM_SHIFT (:523-672)
   function crcl_set_mechanism_unimol(h, form_num, bond_form, break_num, bond_break, form_ref, break_ref, &
                                      form_reac, break_reac) bind(C, name="crcl_set_mechanism_unimol")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: form_num, break_num
      integer(c_int), dimension(*), intent(in) :: bond_form, bond_break
      real(c_double), dimension(*), intent(in) :: form_ref, break_ref, form_reac, break_reac
      integer(c_int) :: crcl_set_mechanism_unimol
   end function crcl_set_mechanism_unimol
   function crcl_set_mechanism_atom_shift(h, shift_atom, shift_coord, shift_lo, shift_hi, shift2_lo, shift2_hi) &
                                          bind(C, name="crcl_set_mechanism_atom_shift")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: shift_atom, shift_coord
      real(c_double), value :: shift_lo, shift_hi, shift2_lo, shift2_hi
      integer(c_int) :: crcl_set_mechanism_atom_shift
   end function crcl_set_mechanism_atom_shift

   ! QMDFF tables of module qmdff / pbc_mod (first and second diabatic state), DG-EVB parameters
   function crcl_set_qmdff(h, T) bind(C, name="crcl_set_qmdff")
      import :: c_ptr, c_int, crcl_qmdff_tables
      type(c_ptr), value :: h
      type(crcl_qmdff_tables), intent(in) :: T
      integer(c_int) :: crcl_set_qmdff
   end function crcl_set_qmdff
   function crcl_set_qmdff2(h, T) bind(C, name="crcl_set_qmdff2")
      import :: c_ptr, c_int, crcl_qmdff_tables
      type(c_ptr), value :: h
      type(crcl_qmdff_tables), intent(in) :: T
      integer(c_int) :: crcl_set_qmdff2
   end function crcl_set_qmdff2
   function crcl_set_dgevb(h, P) bind(C, name="crcl_set_dgevb")
      import :: c_ptr, c_int, crcl_dgevb_params
      type(c_ptr), value :: h
      type(crcl_dgevb_params), intent(in) :: P
      integer(c_int) :: crcl_set_dgevb
   end function crcl_set_dgevb

   ! egrad_water(xyz_act,g_act,e_act) (egrad_water.f90:36): tables once, then crcl_egrad(..., CRCL_PES_WATER, ...)
   function crcl_set_water(h, P) bind(C, name="crcl_set_water")
      import :: c_ptr, c_int, crcl_water_params
      type(c_ptr), value :: h
      type(crcl_water_params), intent(in) :: P
      integer(c_int) :: crcl_set_water
   end function crcl_set_water
   ! HBM-resident path: CUDA-graph replay of steps 2..n of one multi-step call (default on)
   function crcl_set_graph(h, on) bind(C, name="crcl_set_graph")
      import :: c_ptr, c_int
      type(c_ptr), value :: h
      integer(c_int), value :: on
      integer(c_int) :: crcl_set_graph
   end function crcl_set_graph

   ! calc_xi(coords,xi_ideal,xi_act,dxi_act,d2xi_act,mode) (calc_xi.f90:63) for ncoord structures; d2xi may be c_null_ptr
   function crcl_calc_xi(h, ncoord, coords, xi_ideal, mode, xi, dxi, d2xi) bind(C, name="crcl_calc_xi")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: ncoord, mode
      real(c_double), dimension(*), intent(in) :: coords, xi_ideal
      real(c_double), dimension(*), intent(out) :: xi, dxi
      type(c_ptr), value :: d2xi
      integer(c_int) :: crcl_calc_xi
   end function crcl_calc_xi
   ! ewald_recip(n,xyz,q,energy,grad) (ewald_recip.f90:30) for nimg structures
   function crcl_set_ewald(h, P) bind(C, name="crcl_set_ewald")
      import :: c_ptr, c_int, crcl_ewald_params
      type(c_ptr), value :: h
      type(crcl_ewald_params), intent(in) :: P
      integer(c_int) :: crcl_set_ewald
   end function crcl_set_ewald
   function crcl_ewald_recip(h, n, nimg, xyz, q, energy, grad) bind(C, name="crcl_ewald_recip")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: n, nimg
      real(c_double), dimension(*), intent(in) :: xyz, q
      real(c_double), dimension(*), intent(out) :: energy, grad
      integer(c_int) :: crcl_ewald_recip
   end function crcl_ewald_recip
   ! 0: the rfft / irfft pair as written (SURVEY.md F2), 1: true normal-mode transform
   function crcl_set_transform(h, mode) bind(C, name="crcl_set_transform")
      import :: c_ptr, c_int
      type(c_ptr), value :: h
      integer(c_int), value :: mode
      integer(c_int) :: crcl_set_transform
   end function crcl_set_transform
   ! largest batch (trajectories x beads) of few-bead trajectories run in the spread (low-latency) form; 0 = never
   function crcl_set_spread_max_beads(h, max_beads) bind(C, name="crcl_set_spread_max_beads")
      import :: c_ptr, c_int
      type(c_ptr), value :: h
      integer(c_int), value :: max_beads
      integer(c_int) :: crcl_set_spread_max_beads
   end function crcl_set_spread_max_beads
   ! 0 automatic, 1 fused in-register kernels, 2 HBM-resident path
   function crcl_set_path(h, path) bind(C, name="crcl_set_path")
      import :: c_ptr, c_int
      type(c_ptr), value :: h
      integer(c_int), value :: path
      integer(c_int) :: crcl_set_path
   end function crcl_set_path
   function crcl_synchronize(h) bind(C, name="crcl_synchronize")
      import :: c_ptr, c_int
      type(c_ptr), value :: h
      integer(c_int) :: crcl_synchronize
   end function crcl_synchronize
   ! text of the last error of this handle (NUL-terminated C string)
   function crcl_last_error(h) bind(C, name="crcl_last_error")
      import :: c_ptr
      type(c_ptr), value :: h
      type(c_ptr) :: crcl_last_error
   end function crcl_last_error

   ! pbc_mod -> periodic wrap of verlet.f90:591-641 (crcl_set_qmdff / crcl_set_water set it from their tables)
   function crcl_set_box(h, periodic, boxlen) bind(C, name="crcl_set_box")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: periodic
      real(c_double), dimension(3), intent(in) :: boxlen
      integer(c_int) :: crcl_set_box
   end function crcl_set_box
   ! rpmd_check.f90:88-116 inside the step: status bits CRCL_TRAJ_ENERGY / CRCL_TRAJ_XI_RANGE
   function crcl_set_rpmd_check(h, on, energy_ts, energy_tol, xi_tol) bind(C, name="crcl_set_rpmd_check")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: on
      real(c_double), value :: energy_ts, energy_tol, xi_tol
      integer(c_int) :: crcl_set_rpmd_check
   end function crcl_set_rpmd_check
   ! multi-GPU: NCCL communicator behind the C-ABI (one handle per rank); afterwards crcl_recross_children and
   ! crcl_umbrella_windows are collective over the GLOBAL unit range with the reduction inside the library
   function crcl_comm_unique_id(id_out) bind(C, name="crcl_comm_unique_id")
      import :: c_int, c_char
      character(kind=c_char), dimension(128), intent(out) :: id_out
      integer(c_int) :: crcl_comm_unique_id
   end function crcl_comm_unique_id
   function crcl_comm_init(h, nranks, rank, unique_id) bind(C, name="crcl_comm_init")
      import :: c_ptr, c_int, c_char
      type(c_ptr), value :: h
      integer(c_int), value :: nranks, rank
      character(kind=c_char), dimension(128), intent(in) :: unique_id
      integer(c_int) :: crcl_comm_init
   end function crcl_comm_init
   function crcl_comm_destroy(h) bind(C, name="crcl_comm_destroy")
      import :: c_ptr, c_int
      type(c_ptr), value :: h
      integer(c_int) :: crcl_comm_destroy
   end function crcl_comm_destroy

   ! custom_grad / external_grad stay on the host: fn(xyz, e, g, natoms, user) is called per bead
   function crcl_set_host_gradient_cb(h, fn, user) bind(C, name="crcl_set_host_gradient_cb")
      import :: c_ptr, c_funptr, c_int
      type(c_ptr), value :: h
      type(c_funptr), value :: fn
      type(c_ptr), value :: user
      integer(c_int) :: crcl_set_host_gradient_cb
   end function crcl_set_host_gradient_cb
end interface

contains

!
!     gpu_init: call once after read_pes / calc_rate_read (all globals below are set by then)
!
subroutine gpu_init(rank, pes_id, seed, psize)
use general      ! natoms, mass(:), at_move(:), kelvin, thermostat, nose_q
use evb_mod      ! nbeads, beta, andersen_step, bond_form, bond_break, form_ref, break_ref, ...
implicit none
include 'mpif.h' ! as the reference's drivers do (calc_rate.f90, recross.f90)
integer, intent(in) :: rank, pes_id
integer(kind=8), intent(in) :: seed      ! RANDOM_SEED of the key file (or a clock value broadcast from rank 0): the SAME
                                         ! on every rank -- streams are told apart by the trajectory number, not the seed
integer, intent(in), optional :: psize   ! number of MPI ranks: > 1 sets up the NCCL communicator behind the C-ABI
character(kind=c_char), dimension(128) :: uid
integer :: ierr
integer(c_int) :: rc, i, k, n
integer(c_int), allocatable :: amove(:), bf(:), bb(:), atr(:)
real(kind=8) :: dt_dummy
allocate(amove(natoms))
amove = 0
do i = 1, natoms
   if (at_move(i)) amove(i) = 1
end do
dt_dummy = 0.d0   ! the time step is passed per call through gpu_set_dt (drivers convert units late)
rc = crcl_create(crcl_h, int(mod(rank, 8), c_int), int(natoms, c_int), int(nbeads, c_int), &
                 mass(1:natoms), amove, beta, dt_dummy, int(pes_id, c_int))
if (rc .ne. 0) then
   write(*,*) "caracal_gpu: crcl_create failed with code", rc
   call fatal
end if
! MECHA{} tables: bond_form(form_num,2) -> flattened pairs, at_reac(sum_reacs,200) -> concatenated
allocate(bf(2*form_num), bb(2*break_num))
do i = 1, form_num
   bf(2*i-1) = bond_form(i,1); bf(2*i) = bond_form(i,2)
end do
do i = 1, break_num
   bb(2*i-1) = bond_break(i,1); bb(2*i) = bond_break(i,2)
end do
n = sum(n_reac(1:sum_reacs))
allocate(atr(n))
n = 0
do k = 1, sum_reacs
   do i = 1, n_reac(k)
      n = n + 1
      atr(n) = at_reac(k,i)
   end do
end do
rc = crcl_set_mechanism(crcl_h, int(form_num, c_int), bf, int(break_num, c_int), bb, form_ref, break_ref, &
                        int(sum_reacs, c_int), int(n_reac(1:sum_reacs), c_int), atr, R_inf)
rc = crcl_set_thermostat(crcl_h, int(thermostat, c_int), int(andersen_step, c_int), kelvin, nose_q)
!     counter-based RNG: one seed for the whole job (replaces random_init_local, andersen.f90:131); the trajectory
!     this rank propagates starts as its rank number, gpu_new_trajectory sets it per window / trajectory / round
rc = crcl_set_seed(crcl_h, int(seed, c_int64_t))
crcl_traj(1) = int(rank, c_int32_t)
crcl_event(1) = 0_c_int32_t
!     the one exchange of the path: rank 0 makes the NCCL id, MPI ships its 128 bytes, every rank joins
if (present(psize)) then
   if (psize .gt. 1) then
      if (rank .eq. 0) rc = crcl_comm_unique_id(uid)
      call mpi_bcast(uid, 128, MPI_CHARACTER, 0, MPI_COMM_WORLD, ierr)
      rc = crcl_comm_init(crcl_h, int(psize, c_int), int(rank, c_int), uid)
      if (rc .ne. 0) then
         write(*,*) "caracal_gpu: crcl_comm_init failed with code", rc
         call fatal
      end if
   end if
end if
end subroutine gpu_init

!
!     gpu_new_trajectory: call where the drivers start a NEW trajectory (a new umbrella window / trajectory,
!     calc_rate.f90:1387; a new recrossing parent, recross.f90:249): gives it its own RNG stream.  traj_id must be
!     unique in the job, e.g. window*umbr_traj + j, and must not depend on the number of ranks.
!
subroutine gpu_new_trajectory(traj_id)
integer, intent(in) :: traj_id
crcl_traj(1) = int(traj_id, c_int32_t)
crcl_event(1) = 0_c_int32_t
end subroutine gpu_new_trajectory

!
!     mdinit_gpu: same argument list as mdinit (mdinit.f90:40); draws the momenta from the current stream and
!     advances its event counter
!
subroutine mdinit_gpu(derivs, xi_ideal, dxi_act, bias_mode, rank)
use general
use evb_mod
integer :: bias_mode, rank
real(kind=8) :: xi_ideal
real(kind=8) :: derivs(3,natoms,nbeads), dxi_act(3,natoms)
integer(c_int) :: rc
real(c_double) :: xi1(1), kf1(1)
rc = crcl_set_thermostat(crcl_h, int(thermostat, c_int), int(andersen_step, c_int), kelvin, nose_q)
xi1(1) = xi_ideal
kf1(1) = 0.d0
if (bias_mode .ne. 0) kf1(1) = k_force(um_window_act)
rc = crcl_mdinit(crcl_h, 1_c_int, int(bias_mode, c_int), xi1, kf1, q_i, p_i, derivs, dxi_act, &
                 c_loc(crcl_traj), c_loc(crcl_event))
if (rc .ne. 0) then
   write(*,*) "caracal_gpu: crcl_mdinit failed with code", rc
   call fatal
end if
end subroutine mdinit_gpu

!
!     gpu_set_qmdff: hand the tables of the first QMDFF (module qmdff, pbc_mod; built by prepare.f90,
!     rdsolvff.f90, setnonb.f90, set_periodic.f90) to the library.  Call once after read_pes.
!     The second state works the same way with the *_two arrays and crcl_set_qmdff2.
!
subroutine gpu_set_qmdff(n_one)
use qmdff        ! at, q, molnum, nmols, bond, vbond, angl, vangl, tors, vtors, nci, c6xy, r0ab, zab,
                 ! r094_mod, sr42, rad, eps1, eps2, hb, vhb, scalehb_glob, scalexb_glob, q_glob, nbond, ...
use pbc_mod      ! periodic, zahn, boxlen_x/y/z, coul_cut, vdw_cut, cut_low, zahn_a, zahn_par
use evb_mod      ! E_zero1
integer, intent(in) :: n_one
type(crcl_qmdff_tables) :: T
integer(c_int) :: rc
T%n = n_one
T%at = c_loc(at);       T%q = c_loc(q);        T%molnum = c_loc(molnum);  T%nmols = nmols
T%nbond = nbond; T%nangl = nangl; T%ntors = ntors; T%nhb = nhb; T%nnci = nnci; T%ldvt = size(vtors,1)
T%bond = c_loc(bond);   T%vbond = c_loc(vbond); T%angl = c_loc(angl);     T%vangl = c_loc(vangl)
T%tors = c_loc(tors);   T%vtors = c_loc(vtors); T%nci = c_loc(nci);       T%c6xy = c_loc(c6xy)
T%r0ab = c_loc(r0ab);   T%zab = c_loc(zab);     T%r094 = c_loc(r094_mod); T%sr42 = c_loc(sr42)
T%rad = c_loc(rad)
T%eps1 = eps1; T%eps2 = eps2
T%periodic = merge(1, 0, periodic); T%zahn = merge(1, 0, zahn)
T%box = (/ boxlen_x, boxlen_y, boxlen_z /)
T%coul_cut = coul_cut; T%vdw_cut = vdw_cut; T%cut_low = cut_low; T%zahn_a = zahn_a; T%zahn_par = zahn_par
T%e_zero = E_zero1
T%hb = c_loc(hb); T%vhb = c_loc(vhb)
T%scalehb = c_loc(scalehb_glob); T%scalexb = c_loc(scalexb_glob); T%q_glob = c_loc(q_glob)
rc = crcl_set_qmdff(crcl_h, T)
if (rc .ne. 0) then
   write(*,*) "caracal_gpu: crcl_set_qmdff failed with code", rc
   call fatal
end if
end subroutine gpu_set_qmdff

!
!     verlet_gpu: same argument list as verlet (verlet.f90:65); operates on the module globals
!     q_i, p_i like the original.  Drop-in for the call sites listed in INTEGRATION.md.
!
subroutine verlet_gpu(istep, dt, derivs, epot, ekin, afm_force, xi_ideal, xi_real, dxi_act, round, &
                      constrain, analyze, rank)
use general
use evb_mod
integer :: istep, round, constrain, rank
real(kind=8) :: dt, epot, ekin, afm_force, xi_ideal, xi_real
real(kind=8) :: derivs(3,natoms,nbeads), dxi_act(3,natoms)
logical :: analyze
integer(c_int) :: rc, st(1)
real(c_double) :: xi1(1), kf1(1), ep1(1), xr1(1)
rc = crcl_set_beta_dt(crcl_h, beta, dt)
rc = crcl_set_thermostat(crcl_h, int(thermostat, c_int), int(andersen_step, c_int), kelvin, nose_q)
xi1(1) = xi_ideal
kf1(1) = 0.d0
if (constrain .ge. 0) kf1(1) = k_force(um_window_act)
st(1) = 0
!     the stream (trajectory number, event counter) lives in the module and is handed in and out on every call:
!     with null pointers every Andersen resample would redraw event 0 of trajectory 0 (ADVICE r1)
rc = crcl_verlet(crcl_h, 1_c_int, 1_c_int, int(istep-1, c_int), int(constrain, c_int), xi1, kf1, &
                 q_i, p_i, derivs, ep1, xr1, dxi_act, st, c_loc(crcl_traj), c_loc(crcl_event))
epot = ep1(1)
xi_real = xr1(1)
!     verlet.f90:1256-1275, invert.f90, verlet.f90:601-640: NaN/Inf coordinate, singular inertia tensor and a
!     runaway periodic wrap are `fatal` in the reference; a failed SHAKE only raises epot (rpmd_check then restarts),
!     CRCL_TRAJ_ENERGY / CRCL_TRAJ_XI_RANGE go back to the caller's rpmd_check logic through epot / xi_real
if (rc .ne. 0 .or. iand(st(1), CRCL_TRAJ_NAN + CRCL_TRAJ_SINGULAR + CRCL_TRAJ_PBC_FAIL) .ne. 0) then
   write(*,*) "Something went wrong during the dynamics (GPU status", st(1), ")"
   call fatal
end if
end subroutine verlet_gpu

end module caracal_gpu
