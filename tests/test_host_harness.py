"""CPU checks of the PRODUCT's device code: the __host__ __device__ PES functors, the reaction
coordinate / umbrella-hams routine and the counter-based RNG are compiled for the host
(tests/host_harness/pes_host.cu) and compared with the literal oracle.  These are independent
derivations (see the headers of caracal_b200/csrc/pes_*.cuh, xi.cuh), so agreement to rounding
validates both; the same comparisons run on the GPU under -m gpu."""
import ctypes

import numpy as np
import pytest

from tests import common as C

dp = ctypes.POINTER(ctypes.c_double)
ip = ctypes.POINTER(ctypes.c_int)
PID = {"h3": 1, "oh3": 2, "ch4h": 3, "brh2": 4, "o3": 5, "ch4oh": 6, "geh4oh": 7, "ch4cn": 8, "clnh3": 9, "nh3oh": 13, "h2co": 14}


def hh_egrad(H, name, q):
    q = np.ascontiguousarray(q, dtype=np.float64)
    V = np.zeros(q.shape[0])
    g = np.zeros_like(q)
    H.hh_egrad(PID[name], q.ctypes.data_as(dp), q.shape[0], V.ctypes.data_as(dp), g.ctypes.data_as(dp))
    return V, g


@pytest.mark.parametrize("name,sigma", [("h3", 0.15), ("h3", 0.5), ("oh3", 0.15), ("oh3", 0.5),
                                        ("ch4h", 0.15), ("ch4h", 0.4), ("brh2", 0.15), ("brh2", 0.5), ("o3", 0.15), ("o3", 0.4),
                                        ("ch4oh", 0.15), ("ch4oh", 0.4), ("geh4oh", 0.15), ("geh4oh", 0.4),
                                        ("ch4cn", 0.15), ("ch4cn", 0.4), ("clnh3", 0.15), ("clnh3", 0.4),
                                        ("nh3oh", 0.15), ("nh3oh", 0.4), ("h2co", 0.1), ("h2co", 0.3)])
def test_pes_functor_matches_oracle(oracle, host_harness, name, sigma):
    rng = np.random.default_rng(C.SEED)
    q = C.ts_cloud(name, 300 if name == "h2co" else 20000, sigma, rng)   # h2co: 390 000 libm pow calls per oracle gradient
    Vo, go, _ = oracle.egrad(name, q)
    Vd, gd = hh_egrad(host_harness, name, q)
    ok = np.isfinite(Vo)
    assert ok.mean() > 0.999
    assert C.rel_err_E(Vd[ok], Vo[ok]).max() < C.tol_energy(name)
    assert C.rel_err_G(gd[ok], go[ok]).max() < C.tol_grad(name)


@pytest.mark.parametrize("name", ["h3", "oh3", "ch4h", "brh2", "o3", "ch4oh", "geh4oh", "ch4cn", "clnh3", "nh3oh", "h2co"])
def test_pes_functor_far_apart(oracle, host_harness, name):
    """reactants 8 ... 45 bohr apart (the umbrella windows of a rate calculation reach DIST_INF): the curves are evaluated
    far outside the region the saddle-point clouds sample -- BKMP2's H2 singlet curve, for one, calls exp(-2e12) at
    30 bohr, which the branch-free exp of the device code has to survive"""
    rng = np.random.default_rng(17)
    q = C.ts_cloud(name, 200 if name == "h2co" else 4000, 0.1, rng)
    frag = [i - 1 for i in C.SYSTEMS[name]["mecha"]["reactants"][-1]]
    rest = [i for i in range(q.shape[1]) if i not in frag]
    d = q[:, frag].mean(axis=1) - q[:, rest].mean(axis=1)
    d /= np.linalg.norm(d, axis=1)[:, None]
    # Br + H2 only to 35 bohr: beyond r(H-Br) = 39 a0 its HBr curves (a product of two exponentials, VHX
    # egrad_brh2.f:446-450) are denormal numbers, and whether TRBAK3's (s/h)/h (:1047) then overflows hangs on the last
    # bit of the matrix elements -- in the reference's own arithmetic as much as in any restatement of it
    far = 35.0 if name == "brh2" else 45.0
    q[:, frag] += (d * rng.uniform(8.0, far, (len(q), 1)))[:, None, :]
    Vo, go, _ = oracle.egrad(name, q)
    Vd, gd = hh_egrad(host_harness, name, q)
    ok = np.isfinite(Vo) & np.isfinite(go.reshape(len(q), -1)).all(axis=1)
    assert ok.mean() > 0.99 and np.isfinite(Vd[ok]).all()
    assert C.rel_err_E(Vd[ok], Vo[ok]).max() < C.tol_energy(name)
    assert C.rel_err_G(gd[ok], go[ok]).max() < C.tol_grad(name)


def test_h3_compact_geometries(oracle, host_harness):
    rng = np.random.default_rng(5)
    q = rng.uniform(-1.6, 1.6, (40000, 3, 3))
    d = np.linalg.norm(q[:, [0, 0, 1]] - q[:, [1, 2, 2]], axis=-1)
    q = q[(d.min(axis=1) > 0.6) & (d.min(axis=1) < 1.15)][:5000]   # compact branch (R < 1.15 a0)
    assert len(q) > 1000
    Vo, go, _ = oracle.egrad("h3", q)
    Vd, gd = hh_egrad(host_harness, "h3", q)
    assert C.rel_err_E(Vd, Vo).max() < C.TOL_EG
    assert C.rel_err_G(gd, go).max() < C.TOL_EG


def test_oh3_stale_dedr_region(oracle, host_harness):
    """VH2O_oh3 (egrad_oh3.f:549-583): for 0.5 gamma (R - Re) >= 43 (an O-H or H-H distance beyond ~37 a0) the routine
    sets Q(I) = 0, skips DEDR(I) and the value the previous routine left in COMMON /POT2CM_oh3/ is swapped and added to
    the gradient.  The device functor reproduces that (no longer an unsampled divergence): separated
    fragments OH + H2, OH2 ... H, O ... H3 at 40-80 a0, energies and gradients against the literal restatement."""
    rng = np.random.default_rng(9)
    base = C.oh3_ts()
    qs = []
    for far in ([3], [2, 3], [1], [1, 2, 3], [2]):
        for dist in (40.0, 55.0, 80.0):
            q = base + rng.normal(0, 0.05, base.shape)
            for a in far:
                q[a] += rng.normal(0, 1.0, 3) / np.sqrt(3) * 0.0 + np.array([dist, 0.3 * a, -0.2 * a])
            qs.append(q)
    q = np.array(qs)
    Vo, go, _ = oracle.egrad("oh3", q)
    Vd, gd = hh_egrad(host_harness, "oh3", q)
    assert np.isfinite(Vo).all() and np.isfinite(go).all()
    assert C.rel_err_E(Vd, Vo).max() < C.TOL_EG
    assert (np.abs(gd - go).max(axis=(1, 2)) < 1e-10 * np.maximum(np.abs(go).max(axis=(1, 2)), 1e-6)).all()
    # the region is really the stale one: a functor that returned 0 there would miss a finite contribution
    assert np.abs(go).max() > 1e-4


def _hh_xi(H, name, x, xi_ideal, mode, beta, hams):
    m = C.masses(name)
    mech = C.mechanism(name)
    bf = np.ascontiguousarray(mech.bond_form, dtype=np.int32)
    bb = np.ascontiguousarray(mech.bond_break, dtype=np.int32)
    nr = np.array([len(r) for r in mech.reactants], dtype=np.int32)
    ar = np.ascontiguousarray(np.concatenate(mech.reactants), dtype=np.int32)
    xi = ctypes.c_double(0.0)
    dxi = np.zeros_like(x)
    hh = np.zeros_like(x)
    rc = H.hh_calc_xi(len(m), m.ctypes.data_as(dp), len(bf), bf.ctypes.data_as(ip), len(bb), bb.ctypes.data_as(ip),
                      mech.form_ref.ctypes.data_as(dp), mech.break_ref.ctypes.data_as(dp), len(nr),
                      nr.ctypes.data_as(ip), ar.ctypes.data_as(ip), ctypes.c_double(mech.R_inf),
                      x.ctypes.data_as(dp), ctypes.c_double(xi_ideal), mode, ctypes.c_double(beta),
                      ctypes.byref(xi), dxi.ctypes.data_as(dp), hh.ctypes.data_as(dp) if hams else None)
    assert rc == 0
    return xi.value, dxi, hh


@pytest.mark.parametrize("name", ["h3", "oh3", "ch4h"])
def test_xi_and_hams_match_oracle(oracle, host_harness, name):
    rng = np.random.default_rng(6)
    beta = C.beta_calc_rate(300.0)
    s = oracle.System(name, 2, C.masses(name), beta, C.dt_au(0.1))
    s.set_mechanism(C.mechanism(name))
    kf = 0.05 * 300.0
    s.set_kforce(kf)
    nat = len(C.masses(name))
    for x in C.ts_cloud(name, 50, 0.1, rng):
        x = np.ascontiguousarray(x)
        for mode in (1, 2):
            xo, dxo = s.calc_xi(x, 0.93, mode)
            xd, dxd, _ = _hh_xi(host_harness, name, x, 0.93, mode, beta, False)
            assert abs(xd - xo) < 1e-13 * max(1, abs(xo))
            assert np.abs(dxd - dxo).max() < 1e-13
        # umbrella mode 0 adds k*(xi-xi0)*dxi + hams to every bead's gradient
        grad = np.zeros((2, nat, 3))
        xr, dxo = s.umbrella(x, 0.93, grad, 0)
        xd, dxd, hams = _hh_xi(host_harness, name, x, 0.93, 1, beta, True)
        want = kf * (xd - 0.93) * dxd + hams
        assert np.abs(want - grad[0]).max() < 1e-10 * max(1.0, np.abs(grad).max())
        assert np.array_equal(grad[0], grad[1])


def test_rng_matches_oracle(oracle, host_harness):
    z = (ctypes.c_double * 2)()
    for traj, event, bead, pair in [(0, 0, 0, 0), (5, 2, 3, 7), (2 ** 31, 9, 15, 8), (123456, 1, 63, 0)]:
        host_harness.hh_normal_pair(ctypes.c_ulonglong(C.SEED), traj, event, bead, pair, z)
        ref = oracle.normals(C.SEED, traj, event, bead, 2 * pair + 2)[-2:]
        assert abs(z[0] - ref[0]) < 1e-14 and abs(z[1] - ref[1]) < 1e-14
