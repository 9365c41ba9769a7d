"""Second, INDEPENDENT checks of the oracle: transcriptions that do not share a line with
oracle/*.c, written straight from the reference source in numpy, plus finite differences of every CBE sub-term.
The oracle cannot be pinned by the reference's own tests (it has none) nor by the reference binary (no
Fortran compiler here, F1); these tests narrow the blind spot "device == oracle because both restate the same
misreading".  Each test names the reference lines it covers (DESIGN.md section 2 lists them).

  free ring polymer   rfft.f90:43-59, irfft.f90:43-61, verlet.f90:401-461    numpy.fft.fft stands where FFTW stands
  periodic wrap       verlet.f90:591-641                                      plain Python loops
  rpmd_check          rpmd_check.f90:76-116
  CBE sub-terms       egrad_ch4h.f:506-712 (stretch), :713-864 (opbend), :865-985 (ipbend) and the same three routines of
                      egrad_ch4oh.f / egrad_geh4oh.f: d(term)/dq by central differences against the routine's own pdot
  QMDFF pair terms    ff_nonb.f90:88-193 (dispersion + repulsion), :339-417 (Coulomb): numpy on the nci list
"""
import ctypes

import numpy as np
import pytest

from tests import common as C

PI_QMDFF = 3.1415926535897932384626433832795029   # qmdff.f90:44


# ---- rfft -> poly -> irfft with numpy.fft ---------------------------------------------------------------------
def rfft_np(x):
    """rfft.f90:43-59 == irfft.f90:43-61: ain = x (complex), FFTW_FORWARD plan, x = sqrt(1/N) * real(aout)"""
    n = len(x)
    return np.sqrt(1.0 / n) * np.real(np.fft.fft(x.astype(np.complex128)))   # numpy's forward sign = FFTW_FORWARD


def free_rp_np(q, p, mass, beta, dt):
    """verlet.f90:401-461 on q, p [nbeads][natoms][3] (Fortran q_i(i,j,k) == q[k-1][j-1][i-1])"""
    nb, na, _ = q.shape
    q, p = q.copy(), p.copy()
    for i in range(3):
        for j in range(na):
            p[:, j, i] = rfft_np(p[:, j, i])
            q[:, j, i] = rfft_np(q[:, j, i])
    for j in range(na):
        poly = np.zeros((4, nb))
        poly[:, 0] = 1.0, 0.0, dt / mass[j], 1.0
        beta_n = beta / nb
        twown = 2.0 / beta_n
        pi_n = PI_QMDFF / nb
        for k in range(1, nb // 2 + 1):
            wk = twown * np.sin(k * pi_n)
            wt = wk * dt
            wm = wk * mass[j]
            poly[:, k] = np.cos(wt), -wm * np.sin(wt), np.sin(wt) / wm, np.cos(wt)
        for k in range(1, (nb - 1) // 2 + 1):
            poly[:, nb - k] = poly[:, k]
        for k in range(nb):
            for i in range(3):
                p_new = p[k, j, i] * poly[0, k] + q[k, j, i] * poly[1, k]
                q[k, j, i] = p[k, j, i] * poly[2, k] + q[k, j, i] * poly[3, k]
                p[k, j, i] = p_new
    for i in range(3):
        for j in range(na):
            p[:, j, i] = rfft_np(p[:, j, i])
            q[:, j, i] = rfft_np(q[:, j, i])
    return q, p


def _free_system(oracle, nb, natoms=3, kelvin=300.0, dt_fs=0.1):
    """an oracle ring polymer on a zero potential: a child step (constrain 2: no thermostat, no bias, no transrot) is
    then exactly the free ring-polymer block"""
    name = "h3"
    m = C.masses(name)[:natoms] * np.array([1.0, 2.0, 16.0])[:natoms]   # different masses per atom
    o = oracle.System(0, nb, m, C.beta_calc_rate(kelvin), C.dt_au(dt_fs))
    o.set_mechanism(C.mechanism(name))
    o.set_custom_grad(lambda xyz: (0.0, np.zeros_like(xyz)))
    return o, m


@pytest.mark.parametrize("nb", [2, 3, 4, 8, 16, 64])
def test_free_ring_polymer_against_numpy_fft(oracle, nb):
    o, m = _free_system(oracle, nb)
    rng = np.random.default_rng(nb)
    q = C.h3_ts()[None] + rng.normal(0, 0.05, (nb, 3, 3))
    p = rng.normal(0, 3.0, (nb, 3, 3))
    for step in range(3):
        qn, pn = free_rp_np(q, p, m, C.beta_calc_rate(300.0), C.dt_au(0.1))
        o.q[:], o.p[:] = q, p
        o.derivs[:] = 0.0
        _, _, st = o.verlet(step + 1, 0.98, 2)
        assert st == 0
        assert np.abs(o.q - qn).max() < 1e-13 * np.abs(qn).max()
        assert np.abs(o.p - pn).max() < 1e-13 * np.abs(pn).max()
        q, p = qn, pn
    # SURVEY F2, seen from the numpy side as well: the pair is not an inverse pair, beads a and N-a coincide
    if nb > 2:
        assert max(np.abs(q[a] - q[nb - a]).max() for a in range(1, nb)) < 1e-12
        assert np.abs(rfft_np(rfft_np(q[:, 0, 0])) - 0.5 * (q[:, 0, 0] + np.roll(q[::-1, 0, 0], 1))).max() < 1e-13


# ---- periodic wrap ----------------------------------------------------------------------------------------------
def wrap_py(q, box):
    """verlet.f90:591-641, plain-box branch, loops as written (i beads, j atoms, k xyz)"""
    q = q.copy()
    nb, na, _ = q.shape
    fatal = False
    for i in range(nb):
        for j in range(na):
            for k in range(3):
                tries = 0
                while q[i, j, k] < 0:
                    q[:, j, k] = q[:, j, k] + box[k]
                    tries += 1
                    if tries > 100:
                        fatal = True
                        break
                while q[i, j, k] > box[k] and not fatal:
                    q[:, j, k] = q[:, j, k] - box[k]
                    tries += 1
                    if tries > 100:
                        fatal = True
                        break
    return q, fatal


def test_periodic_wrap_against_python_loops(oracle):
    nb = 6
    o, m = _free_system(oracle, nb)
    box = np.array([9.0, 7.5, 11.0])
    o.set_box(True, box)
    rng = np.random.default_rng(8)
    nwrapped = 0
    for trial in range(40):
        # atoms near faces, ring polymers that straddle them, a few far outside (several shifts)
        q = rng.uniform(-0.3, 0.3, (nb, 3, 3)) + rng.choice([0.0, 1.0], (1, 3, 3)) * box + \
            rng.choice([0.0, 0.0, 0.0, 2.0, -3.0], (1, 3, 3)) * box
        p = rng.normal(0, 2.0, (nb, 3, 3))
        qf, pf = free_rp_np(q, p, m, C.beta_calc_rate(300.0), C.dt_au(0.1))
        qw, fatal = wrap_py(qf, box)
        assert not fatal
        nwrapped += int((np.abs(qw - qf) > 1.0).any(axis=0).sum())
        o.q[:], o.p[:] = q, p
        o.derivs[:] = 0.0
        _, _, st = o.verlet(1, 0.98, 2)
        assert st == 0
        assert np.abs(o.q - qw).max() < 1e-12
        # every bead inside [0, L] unless the ring polymer itself is wider than the face distance allows
        assert np.abs(o.p - pf).max() < 1e-12 * np.abs(pf).max()
    assert nwrapped > 100
    # beads shifted together: the wrap never changes bead-to-bead differences
    assert np.abs((qw - qw[0]) - (qf - qf[0])).max() < 1e-12
    # the give-up path (more than 100 shifts of one coordinate) is a status bit, not an abort
    q[:, 0, 0] -= 200 * box[0]
    o.q[:], o.p[:] = q, 0.0
    assert o.verlet(1, 0.98, 2)[2] & 64


def test_rpmd_check_bits(oracle):
    """rpmd_check.f90:76-116 restated as status bits: energy above (E_TS + tol) * nbeads, xi out of tolerance"""
    name, nb = "h3", 4
    o = oracle.System(name, nb, C.masses(name), C.beta_calc_rate(300.0), C.dt_au(0.1))
    o.set_mechanism(C.mechanism(name))
    o.set_thermostat(1, 50, 300.0)
    o.set_kforce(15.0)
    o.q[:] = C.ring_polymer(name, nb, np.random.default_rng(0), 0.01)
    o.set_rng(1, 0)
    o.mdinit(0.98, 2)
    e, xr, st = o.verlet(1, 0.98, 0)
    assert st == 0
    e_ts = e / nb
    o.set_rpmd_check(True, e_ts, 0.1, 0.1)
    assert o.verlet(2, 0.98, 0)[2] == 0
    o.set_rpmd_check(True, e_ts, -0.05, 0.1)               # tolerance below the actual energy
    assert o.verlet(3, 0.98, 0)[2] == 8
    o.set_rpmd_check(True, e_ts, 0.1, 1e-6)
    assert o.verlet(4, 0.98, 0)[2] == 32
    assert o.verlet(5, 0.98, 1)[2] & 32 == 0               # recross.f90:275 passes xi_ideal twice: never trips
    assert o.verlet(6, 0.98, 2)[2] == 0                    # children are not checked


# ---- CBE sub-terms by finite differences ------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["ch4h", "ch4oh", "geh4oh", "ch4cn"])
def test_cbe_subterm_gradients_by_finite_differences(oracle, name):
    """stretch / opbend / ipbend separately: the gradient each routine adds to pdot against central differences of
    the energy the same routine returns (egrad_ch4h.f:506,713,865 and the twins in egrad_ch4oh.f, egrad_geh4oh.f).
    A transcription slip in one routine's force part cannot hide behind the other two this way."""
    L = oracle.lib()
    fn = getattr(L, "oracle_%s_parts_grad" % name)
    dp = ctypes.POINTER(ctypes.c_double)
    fn.argtypes = [dp, dp, dp]
    fn.restype = None
    rng = np.random.default_rng(11)
    qs = C.ts_cloud(name, 6, 0.08, rng)
    nc = qs.shape[1] * 3
    bohr2ang = 0.52918                                      # egrad_ch4h.f:231
    worst = np.zeros(3)

    def parts(x):
        x = np.ascontiguousarray(x.reshape(-1))
        out, g = np.zeros(3), np.zeros(3 * nc)
        fn(x.ctypes.data_as(dp), out.ctypes.data_as(dp), g.ctypes.data_as(dp))
        return out, g.reshape(3, nc)
    for q in qs:
        e0, g0 = parts(q)
        h = 1e-5
        fd = np.zeros((3, nc))
        for c in range(nc):
            xp, xm = q.reshape(-1).copy(), q.reshape(-1).copy()
            xp[c] += h
            xm[c] -= h
            fd[:, c] = (parts(xp)[0] - parts(xm)[0]) / (2 * h) / bohr2ang   # per Angstrom, as pdot is
        for k in range(3):
            scale = max(np.abs(g0[k]).max(), 1e-3)
            worst[k] = max(worst[k], np.abs(fd[k] - g0[k]).max() / scale)
    assert (worst < 2e-6).all(), worst


# ---- QMDFF non-covalent pair terms in numpy -----------------------------------------------------------------------
def nci_energy_np(T, x):
    """ff_nonb.f90:88-193 + :339-417 for a non-periodic, single-molecule table set: BJ-type dispersion
    -eps2 (c6/(r^6+R0^6) + sr42 c6/(r^8+R0^8)), repulsion eps2 zab exp(-r0ab r)/r for r < 25, Coulomb
    q_i q_j eps1 / r inside coul_cut, all over the nci list (1-based (i, j, class) rows)."""
    e = 0.0
    at, q = T["at"], T["q"]
    for (i1, i2, nk) in np.asarray(T["nci"]).reshape(-1, 3):
        vab = x[i1 - 1] - x[i2 - 1]
        r2 = vab @ vab
        r = np.sqrt(r2)
        iz1, iz2 = at[i1 - 1] - 1, at[i2 - 1] - 1
        R0 = T["r094"][iz1, iz2]
        c6 = T["c6xy"][i2 - 1, i1 - 1]
        r6 = r2 * r2 * r2
        r06 = R0 ** 6
        t6, t8 = r6 + r06, r6 * r2 + r06 * R0 * R0
        e -= (c6 / t6 + T["sr42"][iz1, iz2] * c6 / t8) * T["eps2"][nk - 1]
        if r < 25:
            e += T["zab"][iz1, iz2] * np.exp(-T["r0ab"][iz1, iz2] * r) / r * T["eps2"][nk - 1]
        if r <= T["coul_cut"]:
            e += q[i1 - 1] * q[i2 - 1] / r * T["eps1"][nk - 1]
    return e


def test_qmdff_nci_pair_terms_against_numpy(oracle):
    """the non-covalent list terms of one QMDFF summed in numpy straight from ff_nonb.f90, against the oracle's
    ff_nonb restatement with every bonded list emptied; gradient by finite differences of the numpy energy"""
    from tests.qmdff_synth import make_system
    T = dict(make_system(nmol=1, seed=5, periodic=False, zahn=False, hb=False))
    for k, w in (("bond", 2), ("angl", 3), ("tors", 6)):
        T["n" + k] = 0
        T[k] = np.zeros((0, w), dtype=np.int32)
    T.update(vbond=np.zeros((0, 3)), vangl=np.zeros((0, 2)), vtors=np.zeros((0, np.asarray(T["vtors"]).shape[-1])))
    Q = oracle.Qmdff(T)
    rng = np.random.default_rng(2)
    x = T["xyz"] + rng.normal(0, 0.03, T["xyz"].shape)
    V, g = Q.egrad(x[None])
    ref = nci_energy_np(T, x) + T.get("e_zero", 0.0)
    assert len(np.asarray(T["nci"]).reshape(-1, 3)) > 5
    assert abs(V[0] - ref) < 1e-12 * max(1.0, abs(ref)), (V[0], ref)
    h = 1e-5
    for c in [(0, 0), (2, 1), (T["n"] - 1, 2)]:
        xp, xm = x.copy(), x.copy()
        xp[c] += h
        xm[c] -= h
        fd = (nci_energy_np(T, xp) - nci_energy_np(T, xm)) / (2 * h)
        assert abs(fd - g[0][c]) < 1e-7 * max(1.0, abs(fd)), (c, fd, g[0][c])
