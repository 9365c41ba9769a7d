"""Host-side mirror of the reference's interface for the RPMD hot path, over the C-ABI.

Names and argument meaning follow the reference:
  egrad_h3 / egrad_oh3 / egrad_ch4h (q, Natoms, Nbeads) -> V, dVdq, info   (egrad_*.f)
  RPMD.mdinit / RPMD.verlet                                                  (mdinit.f90, verlet.f90)
  RPMD.recross_children                                                      (recross.f90 worker body)
  RPMD.umbrella_window                                                       (calc_rate.f90 worker body)
Arrays are numpy float64 in the reference's layout [traj][bead][atom][xyz]
(= Fortran X(3,natoms,nbeads) per trajectory).  Everything computes on the GPU through
libcaracal_gpu.so; there is no CPU path.
"""
import ctypes

import numpy as np

from . import lib as _l

# ---- unit helpers: the literal semantics of the reference's drivers ----------
_F32 = np.float32
EMASS = 5.485799095e-4  # general.f90:252
AMU = {"H": 1.00782503207, "D": 2.0141017778, "C": 12.00000, "N": 14.0030740048, "O": 15.99491461956,
       "F": 18.99840, "S": 32.06000, "CL": 34.96885268, "BR": 78.9183371, "GE": 72.61}   # atommass.f90:58-129


def atomic_mass_au(symbol):
    """atommass.f90:58-216: isotope mass in amu divided by emass."""
    return AMU[symbol.upper()] / EMASS


def dt_au(dt_fs):
    """dt = dt/2.41888428E-2 with a REAL*4 literal (dynamic.f90:589, calc_rate.f90:466)."""
    return float(dt_fs) / float(_F32(2.41888428e-2))


def beta_calc_rate(kelvin):
    """beta = 1/(kelvin*k_B), k_B = 3.16681517576e-06 as REAL*4 (calc_rate.f90:144,476)."""
    return 1.0 / (float(kelvin) * float(_F32(3.16681517576e-06)))


def beta_dynamic(kelvin):
    """beta = 1/(kelvin*0.316679D-5) (dynamic.f90:584)."""
    return 1.0 / (float(kelvin) * 0.316679e-5)


def _dp(a):
    return a.ctypes.data_as(_l.c_double_p) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(_l.c_int_p) if a is not None else None


def _up(a):
    return a.ctypes.data_as(_l.c_u32_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Mechanism:
    """MECHA{} section, BIMOLEC family: 1-based atom indices as in the key file; dist_inf is R_inf in
    BOHR as module evb_mod holds it (the key file's DIST_INF is in Angstrom, calc_rate_read.f90:693)."""

    def __init__(self, bond_form, bond_break, reactants, dist_inf, ts_struc):
        self.bond_form = np.asarray(bond_form, dtype=np.int32).reshape(-1, 2)
        self.bond_break = np.asarray(bond_break, dtype=np.int32).reshape(-1, 2)
        self.reactants = [np.asarray(r, dtype=np.int32) for r in reactants]
        self.R_inf = float(dist_inf)
        ts = _f64(ts_struc)
        # bonds_ref.f90:39-62: reference bond lengths from the TS structure
        self.form_ref = np.array([np.linalg.norm(ts[a - 1] - ts[b - 1]) for a, b in self.bond_form])
        self.break_ref = np.array([np.linalg.norm(ts[a - 1] - ts[b - 1]) for a, b in self.bond_break])


class UnimolMechanism:
    """MECHA{} of the unimolecular types CYCLOREVER / REARRANGE / DECOM_1BOND / ELIMINATION: bond lists plus
    the TS and the reactant structures the reference lengths come from (bonds_ref.f90:39-109)."""
    kind = "unimol"

    def __init__(self, bond_form, bond_break, ts_struc, reac_struc):
        self.bond_form = np.asarray(bond_form, dtype=np.int32).reshape(-1, 2)
        self.bond_break = np.asarray(bond_break, dtype=np.int32).reshape(-1, 2)
        ts, rs = _f64(ts_struc), _f64(reac_struc)
        ln = lambda x, b: np.array([np.linalg.norm(x[a - 1] - x[c - 1]) for a, c in b])
        self.form_ref, self.break_ref = ln(ts, self.bond_form), ln(ts, self.bond_break)
        self.form_reac, self.break_reac = ln(rs, self.bond_form), ln(rs, self.bond_break)


class AtomShiftMechanism:
    """MECHA{ type atom_shift }: shift_atom (1-based), shift_coord 1..6, limits in bohr (calc_rate_read.f90:805-849)"""
    kind = "atom_shift"

    def __init__(self, shift_atom, shift_coord, shift_lo, shift_hi, shift2_lo=0.0, shift2_hi=0.0):
        self.shift_atom, self.shift_coord = int(shift_atom), int(shift_coord)
        self.shift_lo, self.shift_hi, self.shift2_lo, self.shift2_hi = map(float, (shift_lo, shift_hi, shift2_lo, shift2_hi))


class RPMD:
    """One handle per process/GPU: the explicit form of the reference's module globals."""

    def __init__(self, pes, nbeads, mass, beta, dt, device=0, at_move=None):
        self._lib = _l.load()
        self.pes_id = _l.PES_IDS[pes] if isinstance(pes, str) else int(pes)
        self.mass = _f64(mass)
        self.natoms = len(self.mass)
        self.nbeads = int(nbeads)
        self.beta, self.dt = float(beta), float(dt)
        am = None if at_move is None else np.ascontiguousarray(at_move, dtype=np.int32)
        self._h = ctypes.c_void_p()
        rc = self._lib.crcl_create(ctypes.byref(self._h), device, self.natoms, self.nbeads, _dp(self.mass), _ip(am),
                                   self.beta, self.dt, self.pes_id)
        _l.check(rc, None, "crcl_create")

    def close(self):
        if self._h:
            self._lib.crcl_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        _l.check(rc, self._h, what)

    # -- configuration ------------------------------------------------------------------------
    def set_stream(self, cuda_stream):
        self._ck(self._lib.crcl_set_stream(self._h, ctypes.c_void_p(cuda_stream)), "crcl_set_stream")

    def synchronize(self):
        self._ck(self._lib.crcl_synchronize(self._h), "crcl_synchronize")

    def set_beta_dt(self, beta, dt):
        self.beta, self.dt = float(beta), float(dt)
        self._ck(self._lib.crcl_set_beta_dt(self._h, self.beta, self.dt), "crcl_set_beta_dt")

    def set_transform(self, mode):
        self._ck(self._lib.crcl_set_transform(self._h, int(mode)), "crcl_set_transform")

    def set_spread_max_beads(self, max_beads):
        """Largest batch (trajectories x beads) of few-bead trajectories that runs spread over 16 / 32 lanes (0: never)."""
        self._ck(self._lib.crcl_set_spread_max_beads(self._h, int(max_beads)), "crcl_set_spread_max_beads")

    def set_path(self, path):
        self._ck(self._lib.crcl_set_path(self._h, int(path)), "crcl_set_path")

    def set_graph(self, on):
        """Split path: replay steps from a CUDA graph (default on)."""
        self._ck(self._lib.crcl_set_graph(self._h, int(bool(on))), "crcl_set_graph")

    def set_host_gradient(self, fn):
        """fn(xyz[natoms,3]) -> (e, g[natoms,3]): the custom_grad / external_grad plug-in seam."""
        natoms = self.natoms
        CB = ctypes.CFUNCTYPE(None, _l.c_double_p, _l.c_double_p, _l.c_double_p, ctypes.c_int, ctypes.c_void_p)

        def tramp(xyz, e, g, n, user):
            x = np.ctypeslib.as_array(xyz, shape=(natoms, 3))
            ev, gv = fn(x)
            e[0] = float(ev)
            np.ctypeslib.as_array(g, shape=(natoms, 3))[:] = gv
        self._cb = CB(tramp)   # keep alive
        self._ck(self._lib.crcl_set_host_gradient_cb(self._h, ctypes.cast(self._cb, ctypes.c_void_p), None),
                 "crcl_set_host_gradient_cb")

    def set_qmdff(self, T, second=False):
        """T: dict of QMDFF tables named as in the reference's module qmdff (see crcl_qmdff_tables);
        second=True loads the *_two table set of the second diabatic state."""
        keep = {}

        def arr(key, dtype):
            order = "F" if key in ("c6xy", "r0ab", "zab", "r094", "sr42") else "C"
            keep[key] = np.array(T[key], dtype=dtype, order=order)
            return keep[key]
        S = _l.QmdffTables()
        S.n, S.nmols = int(T["n"]), int(T["nmols"])
        S.at, S.q, S.molnum = _ip(arr("at", np.int32)), _dp(arr("q", np.float64)), _ip(arr("molnum", np.int32))
        S.nbond, S.nangl, S.ntors, S.nnci = len(T["bond"]), len(T["angl"]), len(T["tors"]), len(T["nci"])
        S.nhb, S.ldvt = int(T.get("nhb", 0)), int(T["ldvt"])
        S.bond, S.vbond = _ip(arr("bond", np.int32)), _dp(arr("vbond", np.float64))
        S.angl, S.vangl = _ip(arr("angl", np.int32)), _dp(arr("vangl", np.float64))
        S.tors, S.vtors = _ip(arr("tors", np.int32)), _dp(arr("vtors", np.float64))
        S.nci = _ip(arr("nci", np.int32))
        for k in ("c6xy", "r0ab", "zab", "r094", "sr42", "rad"):
            setattr(S, k, _dp(arr(k, np.float64)))
        S.eps1 = (ctypes.c_double * 6)(*T["eps1"])
        S.eps2 = (ctypes.c_double * 6)(*T["eps2"])
        S.periodic, S.zahn = int(T["periodic"]), int(T["zahn"])
        S.box = (ctypes.c_double * 3)(*T["box"])
        S.coul_cut, S.vdw_cut, S.cut_low = float(T["coul_cut"]), float(T["vdw_cut"]), float(T["cut_low"])
        S.zahn_a, S.zahn_par, S.e_zero = float(T["zahn_a"]), float(T["zahn_par"]), float(T["e_zero"])
        if "scalehb" in T:
            S.hb, S.vhb = _ip(arr("hb", np.int32)), _dp(arr("vhb", np.float64))
            S.scalehb, S.scalexb, S.q_glob = _dp(arr("scalehb", np.float64)), _dp(arr("scalexb", np.float64)), \
                _dp(arr("q_glob", np.float64))
        if second:
            self._ck(self._lib.crcl_set_qmdff2(self._h, ctypes.byref(S)), "crcl_set_qmdff2")
        else:
            self._ck(self._lib.crcl_set_qmdff(self._h, ctypes.byref(S)), "crcl_set_qmdff")

    def set_water(self, W):
        """W: dict(n, periodic, zahn, box[3], coul_cut, zahn_a, zahn_par, pars[11], q[n], is_O[n]): the module state
        after water_init.f90 / set_periodic.f90 (pes WATER_SPC)."""
        q = _f64(W["q"])
        o = np.ascontiguousarray(W["is_O"], dtype=np.int32)
        P = _l.WaterParams()
        P.n, P.periodic, P.zahn = int(W["n"]), int(W["periodic"]), int(W["zahn"])
        for d in range(3):
            P.box[d] = float(W["box"][d])
        P.coul_cut, P.zahn_a, P.zahn_par = float(W["coul_cut"]), float(W["zahn_a"]), float(W["zahn_par"])
        for k in range(11):
            P.pars[k] = float(W["pars"][k])
        P.q, P.is_O = _dp(q), _ip(o)
        self._ck(self._lib.crcl_set_water(self._h, ctypes.byref(P)), "crcl_set_water")

    def set_dgevb(self, E):
        """E: dict(mode, coord_def[nat6,5], point_int[npoints,nat6], alph[npoints], b_vec[mat_size], g_thres)"""
        cd = np.ascontiguousarray(E["coord_def"], dtype=np.int32)
        pi = _f64(E["point_int"])
        al, bv = _f64(E["alph"]), _f64(E["b_vec"])
        P = _l.DgevbParams()
        P.mode, P.npoints, P.nat6 = int(E["mode"]), len(al), len(cd)
        P.coord_def, P.point_int, P.alph, P.b_vec = _ip(cd), _dp(pi), _dp(al), _dp(bv)
        P.g_thres = float(E.get("g_thres", 1e-10))
        self._ck(self._lib.crcl_set_dgevb(self._h, ctypes.byref(P)), "crcl_set_dgevb")

    def set_ewald(self, P):
        """P: dict(box[3], a_ewald, nfft, bsorder, bsmod[3, nfft]) as module pbc_mod holds them after
        set_periodic.f90:114-231 (caracal_b200.ewald.ewald_setup builds one for a stand-alone run)."""
        bs = _f64(P["bsmod"]).reshape(3, -1)
        self._ew_keep = bs
        S = _l.EwaldParams()
        for d in range(3):
            S.box[d] = float(P["box"][d])
        S.a_ewald, S.nfft, S.bsorder = float(P["a_ewald"]), int(P["nfft"]), int(P.get("bsorder", 5))
        S.bsmod1, S.bsmod2, S.bsmod3 = (bs[d].ctypes.data_as(_l.c_double_p) for d in range(3))
        self._ck(self._lib.crcl_set_ewald(self._h, ctypes.byref(S)), "crcl_set_ewald")

    def ewald_recip(self, xyz, q):
        """ewald_recip(n,xyz,q,energy,grad) for a batch: xyz [nimg, n, 3] -> energy [nimg], grad [nimg, n, 3]"""
        q = _f64(q)
        x = _f64(xyz).reshape(-1, len(q), 3)
        e = np.zeros(x.shape[0])
        g = np.zeros_like(x)
        self._ck(self._lib.crcl_ewald_recip(self._h, len(q), x.shape[0], _dp(x), _dp(q), _dp(e), _dp(g)),
                 "crcl_ewald_recip")
        return e, g

    def set_mechanism(self, m):
        kind = getattr(m, "kind", "bimolec")
        if kind == "unimol":
            bf = np.ascontiguousarray(m.bond_form, dtype=np.int32)
            bb = np.ascontiguousarray(m.bond_break, dtype=np.int32)
            self._ck(self._lib.crcl_set_mechanism_unimol(self._h, len(bf), _ip(bf), len(bb), _ip(bb), _dp(_f64(m.form_ref)),
                                                         _dp(_f64(m.break_ref)), _dp(_f64(m.form_reac)),
                                                         _dp(_f64(m.break_reac))), "crcl_set_mechanism_unimol")
            return
        if kind == "atom_shift":
            self._ck(self._lib.crcl_set_mechanism_atom_shift(self._h, m.shift_atom, m.shift_coord, m.shift_lo, m.shift_hi,
                                                             m.shift2_lo, m.shift2_hi), "crcl_set_mechanism_atom_shift")
            return
        bf = np.ascontiguousarray(m.bond_form, dtype=np.int32)
        bb = np.ascontiguousarray(m.bond_break, dtype=np.int32)
        nr = np.array([len(r) for r in m.reactants], dtype=np.int32)
        ar = np.ascontiguousarray(np.concatenate(m.reactants), dtype=np.int32)
        fr, br = _f64(m.form_ref), _f64(m.break_ref)
        self._ck(self._lib.crcl_set_mechanism(self._h, len(bf), _ip(bf), len(bb), _ip(bb), _dp(fr), _dp(br), len(nr),
                                              _ip(nr), _ip(ar), m.R_inf), "crcl_set_mechanism")

    def set_thermostat(self, thermostat, andersen_step=0, kelvin=0.0, nose_q=0.0):
        self._ck(self._lib.crcl_set_thermostat(self._h, int(thermostat), int(andersen_step), float(kelvin),
                                               float(nose_q)), "crcl_set_thermostat")

    def set_box(self, periodic, box=None):
        """pbc_mod: periodic, boxlen_x/y/z (bohr) -> periodic wrap of verlet.f90:591-641 (set_qmdff / set_water set it too)"""
        b = _f64([0.0, 0.0, 0.0] if box is None else box)
        self._ck(self._lib.crcl_set_box(self._h, int(bool(periodic)), _dp(b)), "crcl_set_box")

    def set_rpmd_check(self, on, energy_ts=0.0, energy_tol=0.0, xi_tol=0.0):
        """rpmd_check.f90:88-116 after every step of the biased / constrained modes: status bits TRAJ_ENERGY / TRAJ_XI_RANGE"""
        self._ck(self._lib.crcl_set_rpmd_check(self._h, int(bool(on)), float(energy_ts), float(energy_tol), float(xi_tol)),
                 "crcl_set_rpmd_check")

    def set_seed(self, seed):
        self._ck(self._lib.crcl_set_seed(self._h, int(seed)), "crcl_set_seed")

    # -- PES seam -------------------------------------------------------------------------------
    def egrad(self, q, pes_id=None):
        q = _f64(q)
        nimg = q.size // (3 * self.natoms)
        V = np.empty(nimg)
        g = np.empty_like(q)
        info = ctypes.c_int(0)
        self._ck(self._lib.crcl_egrad(self._h, pes_id or self.pes_id, _dp(q), self.natoms, nimg, _dp(V), _dp(g),
                                      ctypes.byref(info)), "crcl_egrad")
        return V, g, info.value

    # -- integrator seam ------------------------------------------------------------------------
    def _shape(self, q):
        q = _f64(q)
        per = self.nbeads * self.natoms * 3
        if q.size % per:
            raise ValueError("q size is not a multiple of nbeads*natoms*3")
        return q.reshape(-1, self.nbeads, self.natoms, 3)

    def mdinit(self, q, bias_mode=0, xi_ideal=None, k_force=None, traj_id=None, event=None):
        q = self._shape(q)
        nt = q.shape[0]
        p = np.zeros_like(q)
        g = np.zeros_like(q)
        dxi = np.zeros((nt, self.natoms, 3))
        xi = None if xi_ideal is None else _f64(np.broadcast_to(xi_ideal, (nt,)))
        kf = None if k_force is None else _f64(np.broadcast_to(k_force, (nt,)))
        tid = None if traj_id is None else np.ascontiguousarray(traj_id, dtype=np.uint32)
        ev = np.zeros(nt, dtype=np.uint32) if event is None else np.ascontiguousarray(event, dtype=np.uint32)
        self._ck(self._lib.crcl_mdinit(self._h, nt, bias_mode, _dp(xi), _dp(kf), _dp(q), _dp(p), _dp(g), _dp(dxi),
                                       _up(tid), _up(ev)), "crcl_mdinit")
        return p, g, dxi, ev

    def verlet(self, q, p, derivs, nsteps=1, istep0=0, constrain=-1, xi_ideal=None, k_force=None, dxi=None,
               status=None, traj_id=None, event=None):
        """Advance in place; returns (epot, xi_real, status)."""
        for a in (q, p, derivs):
            if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous):
                raise ValueError("q, p, derivs must be C-contiguous float64 arrays (updated in place)")
        nt = q.size // (self.nbeads * self.natoms * 3)
        xi = None if xi_ideal is None else _f64(np.broadcast_to(xi_ideal, (nt,)))
        kf = None if k_force is None else _f64(np.broadcast_to(k_force, (nt,)))
        epot = np.zeros(nt)
        xr = np.zeros(nt)
        st = np.zeros(nt, dtype=np.int32) if status is None else status
        tid = None if traj_id is None else np.ascontiguousarray(traj_id, dtype=np.uint32)
        self._ck(self._lib.crcl_verlet(self._h, nt, int(nsteps), int(istep0), int(constrain), _dp(xi), _dp(kf),
                                       _dp(q), _dp(p), _dp(derivs), _dp(epot), _dp(xr), _dp(dxi), _ip(st), _up(tid),
                                       _up(event)), "crcl_verlet")
        return epot, xr, st

    def calc_xi(self, coords, xi_ideal=0.0, mode=1, hams=False):
        c = _f64(coords).reshape(-1, self.natoms, 3)
        n = c.shape[0]
        xid = _f64(np.broadcast_to(xi_ideal, (n,)))
        xi = np.empty(n)
        dxi = np.empty_like(c)
        hh = np.empty_like(c) if hams else None
        self._ck(self._lib.crcl_calc_xi(self._h, n, _dp(c), _dp(xid), int(mode), _dp(xi), _dp(dxi), _dp(hh)),
                 "crcl_calc_xi")
        return (xi, dxi, hh) if hams else (xi, dxi)

    # -- work units -----------------------------------------------------------------------------
    def recross_children(self, q_parents, npairs, child_evol, xi_ideal, pair0=0):
        qp = self._shape(q_parents)
        num = np.zeros(child_evol)
        den = ctypes.c_double(0.0)
        st = np.zeros(max(npairs, 1), dtype=np.int32)
        self._ck(self._lib.crcl_recross_children(self._h, _dp(qp), qp.shape[0], int(pair0), int(npairs),
                                                 int(child_evol), float(xi_ideal), _dp(num), ctypes.byref(den),
                                                 _ip(st)), "crcl_recross_children")
        return num, den.value, st[:npairs]

    def verlet_dev(self, ntraj, nsteps, d_q, d_p, d_derivs, d_epot, d_xi_real, d_dxi, d_status, d_event, istep0=0, constrain=-1,
                   d_xi_ideal=None, d_k_force=None, d_traj_id=None):
        """crcl_verlet_dev: device pointers as ints (torch.Tensor.data_ptr()); asynchronous on the handle's stream"""
        vp = lambda x: ctypes.c_void_p(x) if x else None
        self._ck(self._lib.crcl_verlet_dev(self._h, int(ntraj), int(nsteps), int(istep0), int(constrain), vp(d_xi_ideal),
                                           vp(d_k_force), vp(d_q), vp(d_p), vp(d_derivs), vp(d_epot), vp(d_xi_real), vp(d_dxi),
                                           vp(d_status), vp(d_traj_id), vp(d_event)), "crcl_verlet_dev")

    def recross_children_dev(self, d_q_parents, nparent, npairs, child_evol, xi_ideal, d_num, d_denom, pair0=0,
                             d_status=None):
        """Device-pointer variant (ints from torch.Tensor.data_ptr()); asynchronous on the handle's stream."""
        self._ck(self._lib.crcl_recross_children_dev(self._h, ctypes.c_void_p(d_q_parents), int(nparent), int(pair0),
                                                     int(npairs), int(child_evol), float(xi_ideal),
                                                     ctypes.c_void_p(d_num), ctypes.c_void_p(d_denom),
                                                     ctypes.c_void_p(d_status) if d_status else None),
                 "crcl_recross_children_dev")

    def umbrella_window(self, q0, xi0, k_force, ntraj, equi_steps, sample_steps, traj_id0=0):
        q0 = _f64(q0)
        avg = np.zeros(ntraj)
        var = np.zeros(ntraj)
        st = np.zeros(ntraj, dtype=np.int32)
        self._ck(self._lib.crcl_umbrella_window(self._h, _dp(q0), float(xi0), float(k_force), int(ntraj),
                                                int(equi_steps), int(sample_steps), int(traj_id0), _dp(avg),
                                                _dp(var), _ip(st)), "crcl_umbrella_window")
        return avg, var, st

    def umbrella_windows(self, q0, xi0, k_force, ntraj, equi_steps, sample_steps, traj_id0=0, constrain=0):
        """All windows of the umbrella phase in one batch: q0[nwin, nbeads, natoms, 3], xi0[nwin],
        k_force[nwin] -> avg, var, status of shape [nwin, ntraj].  constrain 0 (calc_rate.f90) or 3 (no
        removal of net translation / rotation)."""
        q0 = _f64(q0)
        xi0, kf = _f64(np.atleast_1d(xi0)), _f64(np.atleast_1d(k_force))
        nwin = len(xi0)
        avg = np.zeros((nwin, ntraj))
        var = np.zeros((nwin, ntraj))
        st = np.zeros((nwin, ntraj), dtype=np.int32)
        self._ck(self._lib.crcl_umbrella_windows(self._h, nwin, _dp(q0), _dp(xi0), _dp(kf), int(ntraj),
                                                 int(equi_steps), int(sample_steps), int(constrain), int(traj_id0),
                                                 _dp(avg), _dp(var), _ip(st)), "crcl_umbrella_windows")
        return avg, var, st

    # -- multi-GPU: NCCL communicator behind the C-ABI ----------------------------------------------
    @staticmethod
    def comm_unique_id():
        """rank 0: the 128 bytes every rank hands to comm_init (ship them with MPI / torch.distributed / a file)"""
        buf = ctypes.create_string_buffer(128)
        _l.check(_l.load().crcl_comm_unique_id(buf), None, "crcl_comm_unique_id")
        return buf.raw

    def comm_init(self, nranks, rank, unique_id):
        """collective; afterwards recross_children(_dev) and umbrella_windows take the GLOBAL unit range on every
        rank and return the result of the whole job (all-reduce inside the library)"""
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        self._ck(self._lib.crcl_comm_init(self._h, int(nranks), int(rank), buf), "crcl_comm_init")

    def comm_destroy(self):
        self._ck(self._lib.crcl_comm_destroy(self._h), "crcl_comm_destroy")

    def comm_info(self):
        n, r, v = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
        self._ck(self._lib.crcl_comm_info(self._h, ctypes.byref(n), ctypes.byref(r), ctypes.byref(v)), "crcl_comm_info")
        return n.value, r.value, v.value

    # -- hooks ------------------------------------------------------------------------------------
    def rng_normals(self, seed, traj, event, bead, n):
        out = np.empty(n)
        self._ck(self._lib.crcl_rng_normals(self._h, int(seed), int(traj), int(event), int(bead), int(n), _dp(out)),
                 "crcl_rng_normals")
        return out

    def launch_count(self):
        return int(self._lib.crcl_launch_count(self._h))

    def last_kernel_ms(self):
        return float(self._lib.crcl_last_kernel_ms(self._h))

    def kernel_timings(self, max_n=256):
        """ms of each trajectory/egrad kernel launched since the previous call (syncs the stream)"""
        buf = np.zeros(max_n)
        n = self._lib.crcl_kernel_timings(self._h, _dp(buf), max_n)
        if n < 0:
            self._ck(n, "crcl_kernel_timings")
        return buf[:n].copy()

    def bench_propagate(self, ntraj, reps=10):
        """(mean ms, best ms, GB/s at the mean) of the split path's propagation kernel on resident data"""
        out = np.zeros(2)
        self._ck(self._lib.crcl_bench_propagate(self._h, int(ntraj), int(reps), _dp(out)), "crcl_bench_propagate")
        gbs = 120.0 * ntraj * self.nbeads * self.natoms / (out[0] * 1e-3) / 1e9
        return out[0], out[1], gbs

    def measure_fp64_tflops(self, iters=8192):
        return float(self._lib.crcl_measure_fp64_tflops(self._h, int(iters)))

    def measure_dmma_tflops(self, iters=8192):
        """FP64 tensor-core (mma.sync.m8n8k4.f64) throughput of this device"""
        return float(self._lib.crcl_measure_dmma_tflops(self._h, int(iters)))

    def bench_transform(self, nbeads, ntraj, reps):
        """free ring-polymer step from shared memory, FMA form against DMMA form (include/caracal_gpu.h
        crcl_bench_transform): dict(ms_dfma, ms_dmma, max_rel_diff, flops, max_dq)"""
        out = np.zeros(5)
        _l.check(self._lib.crcl_bench_transform(self._h, int(nbeads), int(ntraj), int(reps), _dp(out)), self._h, "crcl_bench_transform")
        return dict(ms_dfma=out[0], ms_dmma=out[1], max_rel_diff=out[2], flops=out[3], max_dq=out[4])


# ---- egrad_<pes>(q,Natoms,Nbeads,V,dVdq,info): the reference's PES plug-in signature ---------
_PES_MASS = {_l.PES_H3: ["H"] * 3, _l.PES_OH3: ["O", "H", "H", "H"], _l.PES_CH4H: ["H", "C", "H", "H", "H", "H"],
             _l.PES_BRH2: ["H", "BR", "H"], _l.PES_O3: ["O", "O", "O"],
             _l.PES_CH4OH: ["H", "C", "H", "H", "H", "O", "H"], _l.PES_GEH4OH: ["H", "GE", "H", "H", "H", "O", "H"],
             _l.PES_CH4CN: ["H", "C", "H", "H", "H", "C", "N"],
             _l.PES_CLNH3: ["H", "N", "H", "H", "CL"], _l.PES_NH3OH: ["H", "N", "H", "H", "O", "H"], _l.PES_H2CO: ["C", "O", "H", "H"]}
_egrad_handles = {}


def egrad(pes, q, natoms=None, nbeads=None, device=0):
    pes_id = _l.PES_IDS[pes] if isinstance(pes, str) else int(pes)
    key = (pes_id, device)
    if key not in _egrad_handles:
        m = [atomic_mass_au(s) for s in _PES_MASS[pes_id]]
        _egrad_handles[key] = RPMD(pes_id, 1, m, 1.0, 1.0, device=device)
    q = _f64(q)
    if natoms is not None and natoms != _l.PES_NATOMS[pes_id]:
        raise ValueError("Natoms does not match the surface")
    V, g, info = _egrad_handles[key].egrad(q)
    return V, g.reshape(q.shape), info


def egrad_h3(q, Natoms=3, Nbeads=None):
    return egrad(_l.PES_H3, q, Natoms, Nbeads)


def egrad_oh3(q, Natoms=4, Nbeads=None):
    return egrad(_l.PES_OH3, q, Natoms, Nbeads)


def egrad_brh2(q, Natoms=3, Nbeads=None):
    return egrad(_l.PES_BRH2, q, Natoms, Nbeads)


def egrad_o3(q, Natoms=3, Nbeads=None):
    return egrad(_l.PES_O3, q, Natoms, Nbeads)


def egrad_ch4oh(q, Natoms=7, Nbeads=None):
    return egrad(_l.PES_CH4OH, q, Natoms, Nbeads)


def egrad_geh4oh(q, Natoms=7, Nbeads=None):
    return egrad(_l.PES_GEH4OH, q, Natoms, Nbeads)


def egrad_ch4cn(q, Natoms=7, Nbeads=None):
    return egrad(_l.PES_CH4CN, q, Natoms, Nbeads)


def egrad_clnh3(q, Natoms=5, Nbeads=None):
    return egrad(_l.PES_CLNH3, q, Natoms, Nbeads)


def egrad_nh3oh(q, Natoms=6, Nbeads=None):
    return egrad(_l.PES_NH3OH, q, Natoms, Nbeads)


def egrad_h2co(q, Natoms=4, Nbeads=None):
    """egrad_h2co(cood,natoms,e_evb,pot_grad,info) of main_h2co.f90:3170 takes one structure; here a batch of images"""
    return egrad(_l.PES_H2CO, q, Natoms, Nbeads)


def egrad_ch4h(q, Natoms=6, Nbeads=None):
    return egrad(_l.PES_CH4H, q, Natoms, Nbeads)
