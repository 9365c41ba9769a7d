"""ctypes wrapper of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
never by caracal_b200.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
dp = ctypes.POINTER(ctypes.c_double)
ip = ctypes.POINTER(ctypes.c_int)
PES = {"h3": 1, "oh3": 2, "ch4h": 3, "brh2": 4, "o3": 5, "ch4oh": 6, "geh4oh": 7, "ch4cn": 8, "clnh3": 9, "nh3oh": 13, "h2co": 14}


def build(force=False):
    """Compile the C restatement (gcc; see oracle/Makefile).  Serialised by a file lock: several ranks may call it."""
    import fcntl
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force):
    need = force or not all(os.path.exists(os.path.join(HERE, f)) for f in ("liboracle.so", "liboracle_exact.so"))
    if not need:
        so = os.path.getmtime(os.path.join(HERE, "liboracle.so"))
        need = any(os.path.getmtime(os.path.join(HERE, f)) > so for f in os.listdir(HERE)
                   if f.endswith((".c", ".h")) or f == "Makefile")
    if need:
        subprocess.run(["make", "-C", HERE, "-s", "all"], check=True, capture_output=True)
    return os.path.join(HERE, "liboracle.so")


_libs = {}


def lib(exact=False):
    name = "liboracle_exact.so" if exact else "liboracle.so"
    if name not in _libs:
        build()
        L = ctypes.CDLL(os.path.join(HERE, name))
        L.oracle_sys_create.restype = ctypes.c_void_p
        L.oracle_sys_create.argtypes = [ctypes.c_int, ctypes.c_int, dp, ip, ctypes.c_double, ctypes.c_double,
                                        ctypes.c_int]
        L.oracle_sys_q.restype = dp
        L.oracle_sys_p.restype = dp
        for f in ("oracle_sys_q", "oracle_sys_p", "oracle_sys_free"):
            getattr(L, f).argtypes = [ctypes.c_void_p]
        L.oracle_sys_set_mecha.argtypes = [ctypes.c_void_p, ctypes.c_int, ip, ctypes.c_int, ip, dp, dp, ctypes.c_int,
                                           ip, ip, ctypes.c_double]
        L.oracle_sys_set_thermostat.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                                ctypes.c_double]
        L.oracle_sys_set_kforce.argtypes = [ctypes.c_void_p, ctypes.c_double]
        L.oracle_sys_set_rng.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32]
        L.oracle_sys_inject_normals.argtypes = [ctypes.c_void_p, dp, ctypes.c_long]
        L.oracle_sys_get_nhc.argtypes = [ctypes.c_void_p, dp]
        L.orc_verlet.restype = ctypes.c_int
        L.orc_verlet.argtypes = [ctypes.c_void_p, ctypes.c_int, dp, dp, ctypes.c_double, dp, dp, ctypes.c_int]
        L.orc_mdinit.argtypes = [ctypes.c_void_p, dp, ctypes.c_double, dp, ctypes.c_int]
        L.orc_calc_xi.argtypes = [ctypes.c_void_p, dp, ctypes.c_double, dp, dp, dp, ctypes.c_int]
        L.orc_umbrella.argtypes = [ctypes.c_void_p, dp, ctypes.c_double, dp, dp, dp, ctypes.c_int]
        L.orc_get_centroid.argtypes = [ctypes.c_void_p, dp]
        L.orc_gradient.argtypes = [ctypes.c_void_p, dp, dp, dp]
        L.oracle_sys_get_event.argtypes = [ctypes.c_void_p]
        L.oracle_sys_get_event.restype = ctypes.c_uint32
        L.orc_andersen.argtypes = [ctypes.c_void_p]
        L.orc_transrot.argtypes = [ctypes.c_void_p]
        L.orc_transrot.restype = ctypes.c_int
        L.orc_recross_pair.restype = ctypes.c_int
        L.orc_recross_pair.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_int, dp, dp]
        L.oracle_recross_children.restype = ctypes.c_int
        L.oracle_recross_children.argtypes = [ctypes.c_void_p, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_double, ctypes.c_uint64, ctypes.c_int, dp, dp]
        L.oracle_egrad.restype = ctypes.c_int
        L.oracle_egrad.argtypes = [ctypes.c_int, dp, ctypes.c_int, ctypes.c_int, dp, dp]
        L.oracle_rng_normal_pair.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                             ctypes.c_uint32, dp]
        L.oracle_philox4x32_10.argtypes = [ctypes.POINTER(ctypes.c_uint32)] * 3
        L.oracle_h3_pote.argtypes = [dp, dp, dp]
        L.oracle_oh3_pot.argtypes = [dp, dp, dp]
        L.oracle_ch4h_parts.argtypes = [dp, dp, dp]
        L.oracle_brh2_pot.argtypes = [dp, dp, dp, ip]
        _libs[name] = L
    return _libs[name]


def _f64c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _d(a):
    return a.ctypes.data_as(dp)


def _i(a):
    return a.ctypes.data_as(ip)


def egrad(pes, q, exact=False):
    """Oracle of egrad_<pes> on [nimg][natoms][3]; loops images the way gradient.f90 does."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    pid = PES[pes] if isinstance(pes, str) else pes
    nat = {1: 3, 2: 4, 3: 6, 4: 3, 5: 3, 6: 7, 7: 7, 8: 7, 9: 5, 13: 6, 14: 4}[pid]
    qq = q.reshape(-1, nat, 3)
    V = np.zeros(qq.shape[0])
    g = np.zeros_like(qq)
    info = lib(exact).oracle_egrad(pid, _d(qq), nat, qq.shape[0], _d(V), _d(g))
    return V, g.reshape(q.shape), info


def normals(seed, traj, event, bead, n):
    out = np.empty(n)
    z = (ctypes.c_double * 2)()
    L = lib()
    for pr in range((n + 1) // 2):
        L.oracle_rng_normal_pair(seed, traj, event, bead, pr, z)
        out[2 * pr] = z[0]
        if 2 * pr + 1 < n:
            out[2 * pr + 1] = z[1]
    return out


def philox(ctr, key):
    c = (ctypes.c_uint32 * 4)(*ctr)
    k = (ctypes.c_uint32 * 2)(*key)
    o = (ctypes.c_uint32 * 4)()
    lib().oracle_philox4x32_10(c, k, o)
    return list(o)


class System:
    """One ring polymer with the oracle's explicit copy of the reference's module state."""

    def __init__(self, pes, nbeads, mass, beta, dt, at_move=None, exact=False):
        self.L = lib(exact)
        self.mass = np.ascontiguousarray(mass, dtype=np.float64)
        self.natoms = len(self.mass)
        self.nbeads = nbeads
        am = np.ones(self.natoms, dtype=np.int32) if at_move is None else np.ascontiguousarray(at_move, np.int32)
        pid = PES[pes] if isinstance(pes, str) else pes
        self.h = ctypes.c_void_p(self.L.oracle_sys_create(self.natoms, nbeads, _d(self.mass), _i(am), beta, dt, pid))
        n = 3 * self.natoms * nbeads
        self.q = np.ctypeslib.as_array(self.L.oracle_sys_q(self.h), shape=(n,)).reshape(nbeads, self.natoms, 3)
        self.p = np.ctypeslib.as_array(self.L.oracle_sys_p(self.h), shape=(n,)).reshape(nbeads, self.natoms, 3)
        self.derivs = np.zeros((nbeads, self.natoms, 3))
        self.dxi = np.zeros((self.natoms, 3))
        self._keep = []

    def __del__(self):
        try:
            self.L.oracle_sys_free(self.h)
        except Exception:
            pass

    def set_mechanism(self, m):
        """m: caracal_b200.api.Mechanism-like (1-based indices) -> 0-based for the oracle."""
        kind = getattr(m, "kind", "bimolec")
        if kind == "atom_shift":
            self.L.oracle_sys_set_atom_shift.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_double] * 4
            self.L.oracle_sys_set_atom_shift(self.h, m.shift_atom - 1, m.shift_coord, m.shift_lo, m.shift_hi, m.shift2_lo,
                                             m.shift2_hi)
            return
        if kind == "unimol":
            bf = np.ascontiguousarray(m.bond_form - 1, dtype=np.int32)
            bb = np.ascontiguousarray(m.bond_break - 1, dtype=np.int32)
            fr, br = _f64c(m.form_ref), _f64c(m.break_ref)
            nr, ar = np.zeros(1, dtype=np.int32), np.zeros(1, dtype=np.int32)
            self.L.oracle_sys_set_mecha(self.h, len(bf), _i(bf), len(bb), _i(bb), _d(fr), _d(br), 0, _i(nr), _i(ar), 0.0)
            self.L.oracle_sys_set_unimol.argtypes = [ctypes.c_void_p, dp, dp]
            self.L.oracle_sys_set_unimol(self.h, _d(_f64c(m.form_reac)), _d(_f64c(m.break_reac)))
            return
        bf = np.ascontiguousarray(m.bond_form - 1, dtype=np.int32)
        bb = np.ascontiguousarray(m.bond_break - 1, dtype=np.int32)
        nr = np.array([len(r) for r in m.reactants], dtype=np.int32)
        ar = np.ascontiguousarray(np.concatenate(m.reactants) - 1, dtype=np.int32)
        fr = np.ascontiguousarray(m.form_ref, dtype=np.float64)
        br = np.ascontiguousarray(m.break_ref, dtype=np.float64)
        self.L.oracle_sys_set_mecha(self.h, len(bf), _i(bf), len(bb), _i(bb), _d(fr), _d(br), len(nr), _i(nr), _i(ar),
                                    m.R_inf)

    def set_thermostat(self, thermostat, andersen_step=0, kelvin=0.0, nose_q=0.0):
        self.L.oracle_sys_set_thermostat(self.h, thermostat, andersen_step, kelvin, nose_q)

    def set_custom_grad(self, fn):
        """fn(xyz[natoms,3]) -> (e, g[natoms,3]); the custom_grad.f90:35 plug-in slot"""
        natoms = self.natoms
        CB = ctypes.CFUNCTYPE(None, dp, dp, dp, ctypes.c_int)

        def tramp(xyz, e, g, n):
            x = np.ctypeslib.as_array(xyz, shape=(natoms, 3))
            ev, gv = fn(x)
            e[0] = float(ev)
            np.ctypeslib.as_array(g, shape=(natoms, 3))[:] = gv
        self._cb = CB(tramp)
        self.L.oracle_sys_set_custom_grad.argtypes = [ctypes.c_void_p, CB]
        self.L.oracle_sys_set_custom_grad(self.h, self._cb)

    def set_box(self, periodic, box):
        """pbc_mod: periodic, boxlen_x/y/z (bohr) -> the wrap of verlet.f90:591-641"""
        self.L.oracle_sys_set_box.argtypes = [ctypes.c_void_p, ctypes.c_int, dp]
        self.L.oracle_sys_set_box(self.h, int(periodic), _d(_f64c(box)))

    def set_rpmd_check(self, on, energy_ts=0.0, energy_tol=0.0, xi_tol=0.0):
        """rpmd_check.f90:100-116 after every biased / constrained step"""
        self.L.oracle_sys_set_rpmd_check.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_double] * 3
        self.L.oracle_sys_set_rpmd_check(self.h, int(on), energy_ts, energy_tol, xi_tol)

    def set_kforce(self, k):
        self.L.oracle_sys_set_kforce(self.h, k)

    def set_rng(self, seed, traj, event=0):
        self.L.oracle_sys_set_rng(self.h, seed, traj, event)

    def event(self):
        """number of momentum-resampling events drawn so far on this stream"""
        return int(self.L.oracle_sys_get_event(self.h))

    def inject_normals(self, z):
        z = np.ascontiguousarray(z, dtype=np.float64)
        self._keep.append(z)
        self.L.oracle_sys_inject_normals(self.h, _d(z), z.size)

    def nhc(self):
        out = np.zeros(8)
        self.L.oracle_sys_get_nhc(self.h, _d(out))
        return out

    def mdinit(self, xi_ideal=0.0, bias_mode=0):
        self.L.orc_mdinit(self.h, _d(self.derivs), xi_ideal, _d(self.dxi), bias_mode)

    def gradient_all(self):
        """derivs <- plain PES gradient of every bead, no bias (calc_rate.f90:1619-1623)"""
        e = ctypes.c_double(0.0)
        for b in range(self.nbeads):
            self.L.orc_gradient(self.h, _d(self.q[b]), ctypes.byref(e), _d(self.derivs[b]))

    def verlet(self, istep, xi_ideal=0.0, constrain=-1):
        epot = ctypes.c_double(0.0)
        xr = ctypes.c_double(0.0)
        st = self.L.orc_verlet(self.h, istep, _d(self.derivs), ctypes.byref(epot), xi_ideal, ctypes.byref(xr),
                               _d(self.dxi), constrain)
        return epot.value, xr.value, st

    def calc_xi(self, coords, xi_ideal, mode, hessian=False):
        c = np.ascontiguousarray(coords, dtype=np.float64)
        xi = ctypes.c_double(0.0)
        dxi = np.zeros((self.natoms, 3))
        d2 = np.zeros((self.natoms, 3, self.natoms, 3)) if hessian else None
        self.L.orc_calc_xi(self.h, _d(c), xi_ideal, ctypes.byref(xi), _d(dxi), _d(d2) if hessian else None, mode)
        return (xi.value, dxi, d2) if hessian else (xi.value, dxi)

    def umbrella(self, centroid, xi_ideal, grad, mode):
        c = np.ascontiguousarray(centroid, dtype=np.float64)
        xr = ctypes.c_double(0.0)
        dxi = np.zeros((self.natoms, 3))
        self.L.orc_umbrella(self.h, _d(c), xi_ideal, ctypes.byref(xr), _d(dxi), _d(grad), mode)
        return xr.value, dxi

    def recross_pair(self, xi_ideal, child_evol):
        num = np.zeros(child_evol)
        den = ctypes.c_double(0.0)
        st = self.L.orc_recross_pair(self.h, xi_ideal, child_evol, _d(num), ctypes.byref(den))
        return num, den.value, st

    def recross_children(self, q_parents, pair0, npairs, child_evol, xi_ideal, seed, nthreads=1):
        qp = np.ascontiguousarray(q_parents, dtype=np.float64).reshape(-1, self.nbeads, self.natoms, 3)
        num = np.zeros(child_evol)
        den = ctypes.c_double(0.0)
        st = self.L.oracle_recross_children(self.h, _d(qp), qp.shape[0], pair0, npairs, child_evol, xi_ideal, seed,
                                            nthreads, _d(num), ctypes.byref(den))
        return num, den.value, st


class _QmdffStruct(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int), ("at", ip), ("q", dp),
                ("nbond", ctypes.c_int), ("nangl", ctypes.c_int), ("ntors", ctypes.c_int), ("nnci", ctypes.c_int),
                ("ldvt", ctypes.c_int), ("nmols", ctypes.c_int),
                ("bond", ip), ("vbond", dp), ("angl", ip), ("vangl", dp), ("tors", ip), ("vtors", dp), ("nci", ip),
                ("molnum", ip), ("c6xy", dp), ("r0ab", dp), ("zab", dp), ("r094", dp), ("sr42", dp), ("rad", dp),
                ("eps1", ctypes.c_double * 6), ("eps2", ctypes.c_double * 6),
                ("periodic", ctypes.c_int), ("zahn", ctypes.c_int),
                ("box", ctypes.c_double * 3), ("coul_cut", ctypes.c_double), ("vdw_cut", ctypes.c_double),
                ("cut_low", ctypes.c_double), ("zahn_a", ctypes.c_double), ("zahn_par", ctypes.c_double),
                ("e_zero", ctypes.c_double),
                ("nhb", ctypes.c_int), ("hb", ip), ("vhb", dp), ("scalehb", dp), ("scalexb", dp), ("q_glob", dp)]


class Qmdff:
    """Oracle of gradient.f90:341-362 (nqmdff = 1): ff_eg + ff_nonb + E_zero1 on the given tables."""

    def __init__(self, T):
        self.L = lib()
        self.keep = {}

        def arr(key, dtype):
            order = "F" if key in ("c6xy", "r0ab", "zab", "r094", "sr42") else "C"
            a = np.array(T[key], dtype=dtype, order=order)
            self.keep[key] = a
            return a
        S = _QmdffStruct()
        S.n = int(T["n"])
        S.at = _i(arr("at", np.int32))
        S.q = _d(arr("q", np.float64))
        S.nbond, S.nangl, S.ntors, S.nnci = len(T["bond"]), len(T["angl"]), len(T["tors"]), len(T["nci"])
        S.ldvt, S.nmols = int(T["ldvt"]), int(T["nmols"])
        S.bond, S.vbond = _i(arr("bond", np.int32)), _d(arr("vbond", np.float64))
        S.angl, S.vangl = _i(arr("angl", np.int32)), _d(arr("vangl", np.float64))
        S.tors, S.vtors = _i(arr("tors", np.int32)), _d(arr("vtors", np.float64))
        S.nci, S.molnum = _i(arr("nci", np.int32)), _i(arr("molnum", np.int32))
        for k in ("c6xy", "r0ab", "zab", "r094", "sr42", "rad"):
            setattr(S, k, _d(arr(k, np.float64)))
        S.eps1 = (ctypes.c_double * 6)(*T["eps1"])
        S.eps2 = (ctypes.c_double * 6)(*T["eps2"])
        S.periodic, S.zahn = int(T["periodic"]), int(T["zahn"])
        S.box = (ctypes.c_double * 3)(*T["box"])
        S.coul_cut, S.vdw_cut, S.cut_low = float(T["coul_cut"]), float(T["vdw_cut"]), float(T["cut_low"])
        S.zahn_a, S.zahn_par, S.e_zero = float(T["zahn_a"]), float(T["zahn_par"]), float(T["e_zero"])
        S.nhb = int(T.get("nhb", 0))
        if "scalehb" in T:
            S.hb, S.vhb = _i(arr("hb", np.int32)), _d(arr("vhb", np.float64))
            S.scalehb, S.scalexb, S.q_glob = _d(arr("scalehb", np.float64)), _d(arr("scalexb", np.float64)), \
                _d(arr("q_glob", np.float64))
        self.S = S
        self.n = S.n
        self.L.orc_qmdff_egrad.argtypes = [ctypes.POINTER(_QmdffStruct), dp, ctypes.c_int, dp, dp]
        self.L.orc_ff_eg.argtypes = [ctypes.POINTER(_QmdffStruct), dp, dp, dp]

    def egrad(self, xyz):
        x = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, self.n, 3)
        V = np.zeros(x.shape[0])
        g = np.zeros_like(x)
        self.L.orc_qmdff_egrad(ctypes.byref(self.S), _d(x), x.shape[0], _d(V), _d(g))
        return V, g

    def ff_eg(self, xyz):
        x = np.ascontiguousarray(xyz, dtype=np.float64).reshape(self.n, 3)
        e = ctypes.c_double(0.0)
        g = np.zeros_like(x)
        self.L.orc_ff_eg(ctypes.byref(self.S), _d(x), ctypes.byref(e), _d(g))
        return e.value, g


class _DgevbStruct(ctypes.Structure):
    _fields_ = [("mode", ctypes.c_int), ("npoints", ctypes.c_int), ("nat6", ctypes.c_int), ("natoms", ctypes.c_int),
                ("coord_def", ip), ("point_int", dp), ("alph", dp), ("b_vec", dp), ("g_thres", ctypes.c_double)]


class Dgevb:
    """Oracle of the DG-EVB branch of gradient.f90:365-537 on two QMDFF table sets."""

    def __init__(self, T1, T2, E):
        self.Q1, self.Q2 = Qmdff(T1), Qmdff(T2)
        self.L = self.Q1.L
        self.keep = dict(cd=np.ascontiguousarray(E["coord_def"], dtype=np.int32),
                         pi=np.ascontiguousarray(E["point_int"], dtype=np.float64),
                         al=np.ascontiguousarray(E["alph"], dtype=np.float64),
                         bv=np.ascontiguousarray(E["b_vec"], dtype=np.float64))
        S = _DgevbStruct()
        S.mode, S.npoints, S.nat6, S.natoms = int(E["mode"]), len(E["alph"]), len(E["coord_def"]), self.Q1.n
        S.coord_def, S.point_int, S.alph, S.b_vec = _i(self.keep["cd"]), _d(self.keep["pi"]), _d(self.keep["al"]), \
            _d(self.keep["bv"])
        S.g_thres = float(E.get("g_thres", 1e-10))
        self.S = S
        self.L.orc_dgevb_egrad.argtypes = [ctypes.POINTER(_QmdffStruct), ctypes.POINTER(_QmdffStruct),
                                           ctypes.POINTER(_DgevbStruct), dp, ctypes.c_int, dp, dp]
        self.L.orc_xyz_2int.argtypes = [ctypes.POINTER(_DgevbStruct), dp, dp]
        self.L.orc_qmdff_two_one.argtypes = [ctypes.POINTER(_QmdffStruct), dp, ctypes.POINTER(ctypes.c_double), dp]

    def second_state(self, xyz):
        """E2 + E_zero2 and g2 of one structure (ff_eg_two + ff_nonb_two + ff_hb_two)"""
        x = np.ascontiguousarray(xyz, dtype=np.float64)
        e = ctypes.c_double(0.0)
        g = np.zeros_like(x)
        self.L.orc_qmdff_two_one(ctypes.byref(self.Q2.S), _d(x), ctypes.byref(e), _d(g))
        return e.value + self.Q2.S.e_zero, g

    def egrad(self, xyz):
        n = self.Q1.n
        x = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, n, 3)
        V = np.zeros(x.shape[0])
        g = np.zeros_like(x)
        self.L.orc_dgevb_egrad(ctypes.byref(self.Q1.S), ctypes.byref(self.Q2.S), ctypes.byref(self.S), _d(x),
                               x.shape[0], _d(V), _d(g))
        return V, g

    def internals(self, xyz):
        x = np.ascontiguousarray(xyz, dtype=np.float64)
        out = np.zeros(self.S.nat6)
        self.L.orc_xyz_2int(ctypes.byref(self.S), _d(x), _d(out))
        return out


class Ewald:
    """Oracle of set_periodic.f90:114-231 (SPME set-up) and ewald_recip.f90 for an orthorhombic box."""

    def __init__(self, box):
        self.L = lib()
        self.box = np.ascontiguousarray(box, dtype=np.float64)
        self.L.orc_ewald_setup.restype = ctypes.c_void_p
        self.L.orc_ewald_setup.argtypes = [dp]
        self.L.orc_ewald_free.argtypes = [ctypes.c_void_p]
        self.L.orc_ewald_nfft.argtypes = [ctypes.c_void_p]
        self.L.orc_ewald_alpha.argtypes = [ctypes.c_void_p]
        self.L.orc_ewald_alpha.restype = ctypes.c_double
        self.L.orc_ewald_bsmod.argtypes = [ctypes.c_void_p, dp]
        self.L.orc_ewald_recip.argtypes = [ctypes.c_void_p, ctypes.c_int, dp, dp, dp, dp]
        self.L.orc_ewald_direct_recip.argtypes = [dp, ctypes.c_double, ctypes.c_int, ctypes.c_int, dp, dp, dp, dp]
        self.h = ctypes.c_void_p(self.L.orc_ewald_setup(_d(self.box)))
        self.nfft = int(self.L.orc_ewald_nfft(self.h))
        self.a_ewald = float(self.L.orc_ewald_alpha(self.h))
        self.bsorder = 5
        bs = np.zeros(3 * self.nfft)
        self.L.orc_ewald_bsmod(self.h, _d(bs))
        self.bsmod = bs.reshape(3, self.nfft)

    def __del__(self):
        try:
            self.L.orc_ewald_free(self.h)
        except Exception:
            pass

    def recip(self, xyz, q):
        x = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        qq = np.ascontiguousarray(q, dtype=np.float64)
        e = ctypes.c_double(0.0)
        g = np.zeros_like(x)
        self.L.orc_ewald_recip(self.h, len(qq), _d(x), _d(qq), ctypes.byref(e), _d(g))
        return e.value, g

    def direct_recip(self, xyz, q, mmax):
        x = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        qq = np.ascontiguousarray(q, dtype=np.float64)
        e = ctypes.c_double(0.0)
        g = np.zeros_like(x)
        self.L.orc_ewald_direct_recip(_d(self.box), self.a_ewald, int(mmax), len(qq), _d(x), _d(qq), ctypes.byref(e), _d(g))
        return e.value, g


class _WaterStruct(ctypes.Structure):
    _fields_ = [("natoms", ctypes.c_int), ("periodic", ctypes.c_int), ("zahn", ctypes.c_int),
                ("box", ctypes.c_double * 3), ("coul_cut", ctypes.c_double), ("zahn_a", ctypes.c_double),
                ("zahn_par", ctypes.c_double), ("pars", ctypes.c_double * 11), ("q", dp), ("is_O", ip)]


def water_default_pars():
    """water_init.f90:75-101: the eleven parameters in atomic units"""
    p = (ctypes.c_double * 11)()
    lib().orc_water_default_pars(p)
    return np.array(list(p))


class Water:
    """Oracle of egrad_water.f90 (pes WATER_SPC).  W: dict(n, periodic, zahn, box, coul_cut, zahn_a, zahn_par, pars,
    q, is_O) -- what water_init.f90 and set_periodic.f90 leave in the modules."""

    def __init__(self, W):
        self.L = lib()
        self.n = int(W["n"])
        self.keep = dict(q=np.ascontiguousarray(W["q"], dtype=np.float64),
                         o=np.ascontiguousarray(W["is_O"], dtype=np.int32))
        S = _WaterStruct()
        S.natoms, S.periodic, S.zahn = self.n, int(W["periodic"]), int(W["zahn"])
        S.box = (ctypes.c_double * 3)(*[float(v) for v in W["box"]])
        S.coul_cut, S.zahn_a, S.zahn_par = float(W["coul_cut"]), float(W["zahn_a"]), float(W["zahn_par"])
        S.pars = (ctypes.c_double * 11)(*[float(v) for v in W["pars"]])
        S.q, S.is_O = _d(self.keep["q"]), _i(self.keep["o"])
        self.S = S
        self.L.orc_water_egrad.argtypes = [ctypes.POINTER(_WaterStruct), dp, ctypes.c_int, dp, dp]

    def egrad(self, xyz):
        x = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, self.n, 3)
        V = np.zeros(x.shape[0])
        g = np.zeros_like(x)
        self.L.orc_water_egrad(ctypes.byref(self.S), _d(x), x.shape[0], _d(V), _d(g))
        return V, g
