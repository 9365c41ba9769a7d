"""CPU tests of the NH3 + Cl and NH3 + OH oracles (oracle/pes_clnh3.c, pes_nh3oh.c <- egrad_clnh3.f, egrad_nh3oh.f; SURVEY.md
8f row N4).  The reference ships no outputs for these surfaces and cannot be compiled here, so the restatement is pinned by
what does not share its lines: finite differences per term (with the geometry-dependent reference length of the source
held fixed, since the source's analytic gradient treats it as a constant), the numeric-gradient loop of egrad_nh3oh.f redone
in numpy, permutation / rigid-motion invariance, and known answers for the constants (fragment geometries and reaction
energies from experiment)."""
import ctypes

import numpy as np
import pytest
from scipy.optimize import minimize

from tests import common as C

KCAL = 627.509474
ANG = 0.52918            # the surface's own bohr -> Angstrom factor (egrad_clnh3.f:155)
dp = ctypes.POINTER(ctypes.c_double)


def _d(a):
    return a.ctypes.data_as(dp)


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as O
    O.build()
    L = O.lib()
    L.oracle_clnh3_parts.argtypes = [dp, dp, dp]
    L.oracle_clnh3_parts_grad.argtypes = [dp, dp, dp]
    L.oracle_clnh3_parts_frozen.argtypes = [dp, ctypes.c_double, dp, dp]
    L.oracle_nh3oh_parts.argtypes = [dp, dp, dp]
    return O


def energy(O, name, q):
    return O.egrad(name, np.asarray(q)[None])[0][0]


def central(f, q, h=1e-4):
    g = np.zeros_like(q)
    for idx in np.ndindex(q.shape):
        qp, qm = q.copy(), q.copy()
        qp[idx] += h
        qm[idx] -= h
        g[idx] = (f(qp) - f(qm)) / (2 * h)
    return g


def test_clnh3_terms_have_consistent_gradients_at_fixed_r0ch(oracle):
    """stretch and in-plane bend separately: analytic d(part)/dq of the restatement against central differences of the
    same part with r0ch held at the value the geometry gives (egrad_clnh3.f:287-299 makes r0ch a function of the three
    N-H lengths; stretch :688-783, ipbend :990-1050 and ipforce :1478-1540 differentiate as if it were a constant)"""
    L = oracle.lib()
    q = C.ts_cloud("clnh3", 5, 0.1, np.random.default_rng(11))
    for x in q:
        x = np.ascontiguousarray(x.reshape(15))
        parts, gp, V = np.zeros(3), np.zeros(30), np.zeros(1)
        L.oracle_clnh3_parts_grad(_d(x), _d(parts), _d(gp))
        r0 = parts[1]
        assert 1.0141 <= r0 <= 1.027

        def part(y, k):
            pp, vv = np.zeros(3), np.zeros(1)
            L.oracle_clnh3_parts_frozen(_d(np.ascontiguousarray(y)), r0, _d(pp), _d(vv))
            return pp[k]
        for k, row in ((0, 0), (2, 1)):
            gn = central(lambda y: part(y, k), x, h=1e-4) / ANG          # per Angstrom
            ga = gp[15 * row:15 * row + 15]
            assert np.abs(gn - ga).max() < 2e-6 * max(np.abs(ga).max(), 1.0), (k, np.abs(gn - ga).max())
        # the returned energy and gradient are the unit-converted sums (:185-202)
        Vt, g, _ = oracle.egrad("clnh3", x.reshape(1, 5, 3))
        assert abs(Vt[0] - (parts[0] + parts[2]) * 0.03812) < 1e-15
        assert np.abs(g.reshape(15) - (gp[:15] + gp[15:]) * 0.0201723).max() < 1e-16


def test_clnh3_gradient_misses_the_r0ch_term_as_in_the_source(oracle):
    """the gradient returned is NOT the derivative of the energy returned: the difference is d r0ch / d rch, largest where
    an N-H bond is stretched; it vanishes for displacements of the chlorine, which r0ch does not depend on"""
    ts = C.SYSTEMS["clnh3"]["ts"]()
    V, g, _ = oracle.egrad("clnh3", ts[None])
    gn = central(lambda y: energy(oracle, "clnh3", y), ts)
    miss = np.abs(gn - g[0])
    assert miss[:4].max() > 1e-5                                       # N and H rows carry the missing term
    assert miss[4].max() < 1e-5 * np.abs(g).max() + 2e-8               # Cl row: unit-factor mismatch only


@pytest.mark.parametrize("name,perm", [("clnh3", [2, 1, 0, 3, 4]), ("clnh3", [3, 1, 2, 0, 4]), ("clnh3", [0, 1, 3, 2, 4]),
                                       ("nh3oh", [2, 1, 0, 3, 4, 5]), ("nh3oh", [0, 1, 3, 2, 4, 5])])
def test_equivalent_hydrogens_permute(oracle, name, perm):
    q = C.ts_cloud(name, 4, 0.1, np.random.default_rng(5))
    V, g, _ = oracle.egrad(name, q)
    Vp, gpm, _ = oracle.egrad(name, q[:, perm])
    assert np.abs(V - Vp).max() < 1e-13
    tol = 1e-12 if name == "clnh3" else 5e-9       # nh3oh: forward differences, one ulp of E is 6e-12 Eh/bohr
    assert np.abs(g[:, perm] - gpm).max() < tol


@pytest.mark.parametrize("name", ["clnh3", "nh3oh"])
def test_rigid_motions_leave_the_energy(oracle, name):
    rng = np.random.default_rng(2)
    q = C.ts_cloud(name, 3, 0.1, rng)
    A, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    V, g, _ = oracle.egrad(name, q)
    V2, g2, _ = oracle.egrad(name, q @ A.T + rng.normal(size=3))
    assert np.abs(V - V2).max() < 1e-13
    # nh3oh: a forward difference along rotated axes differs in its O(h) truncation term (h * second derivative / 2)
    tol = 1e-12 if name == "clnh3" else 5e-5
    assert np.abs(g @ A.T - g2).max() < tol
    if name == "clnh3":
        assert np.abs(g.sum(axis=1)).max() < 1e-15                      # pair terms: no net force


def test_nh3oh_gradient_is_the_forward_difference_loop(oracle):
    """POT_nh3oh :283-296 redone outside the restatement: E from the oracle's energy entry at coordinates displaced by
    PASO = 1e-5 A one at a time, every earlier coordinate left at (q + h) - h.  The displaced geometry has to pass through
    the bohr interface (one rounding of the coordinate = 1e-16 g in E = 1e-11 in the quotient)."""
    L = oracle.lib()
    h = 1.0e-5
    q = C.ts_cloud("nh3oh", 3, 0.1, np.random.default_rng(8))
    for x in q:
        V, g, _ = oracle.egrad("nh3oh", x[None])
        qa = x.reshape(18) * ANG
        grad = np.zeros(18)
        for i in range(18):
            qa[i] = qa[i] + h
            pp, vv = np.zeros(3), np.zeros(1)
            L.oracle_nh3oh_parts(_d(np.ascontiguousarray(qa / ANG)), _d(pp), _d(vv))
            grad[i] = (vv[0] - V[0]) / h * ANG
            qa[i] = qa[i] - h
        assert np.abs(grad - g.reshape(18)).max() < 1e-8
        # and it is the derivative of the energy to first order in the step
        gc = central(lambda y: energy(oracle, "nh3oh", y), x, h=1e-4)
        assert np.abs(gc - g[0]).max() < 3e-4 * np.abs(g).max()


def _relax(O, name, x0, free):
    """minimise the oracle energy over the Cartesian rows `free` (others fixed), energy only"""
    x0 = np.array(x0, dtype=float)

    def f(v):
        x = x0.copy()
        x[free] = v.reshape(-1, 3)
        return energy(O, name, x) * KCAL
    res = minimize(f, x0[free].ravel(), method="BFGS", options=dict(gtol=1e-7))
    x = x0.copy()
    x[free] = res.x.reshape(-1, 3)
    return x, res.fun


def _angle(a, b, c):
    u, v = a - b, c - b
    return np.degrees(np.arccos(u @ v / np.linalg.norm(u) / np.linalg.norm(v)))


def _far_reactants(name):
    ts = C.SYSTEMS[name]["ts"]()
    x = ts.copy()
    u = ts[0] / np.linalg.norm(ts[0])
    x[0] = u * 1.014 / C.BOHR
    shift = u * 30.0
    x[4:] = ts[4:] + shift                      # Cl, or O-H, 30 bohr further out
    return x


def _far_products(name):
    ts = C.SYSTEMS[name]["ts"]()
    x = ts.copy()
    u = ts[0] / np.linalg.norm(ts[0])
    shift = u * 30.0
    x[0] = ts[0] + shift
    x[4:] = ts[4:] + shift
    return x


def test_clnh3_fragments_and_reaction_energy(oracle):
    """known answers for the BLOCK DATA (:1846-1898): ammonia (r0chr, tau), the amino radical (r0chp, taunh2), hydrogen
    chloride (r0hh, d1hh) and the reaction energy d1ch - d1hh.  Experiment: NH3 r = 1.012 A, 106.7 deg; NH2 r = 1.024 A,
    103.4 deg; HCl r = 1.275 A; Cl + NH3 -> HCl + NH2 classical endothermicity D_e(H-NH2) - D_e(HCl) ~ 9-10 kcal/mol
    (D_0 106.7 - 102.2 plus 5 kcal/mol of zero-point energy)."""
    xr, er = _relax(oracle, "clnh3", _far_reactants("clnh3"), [0, 2, 3])
    r = [np.linalg.norm(xr[i] - xr[1]) * C.BOHR for i in (0, 2, 3)]
    assert np.abs(np.array(r) - 1.012).max() < 0.005
    assert abs(_angle(xr[0], xr[1], xr[2]) - 106.7) < 3.0
    xp, ep = _relax(oracle, "clnh3", _far_products("clnh3"), [0, 2, 3])
    assert abs(np.linalg.norm(xp[0] - xp[4]) * C.BOHR - 1.275) < 0.005
    assert np.abs(np.array([np.linalg.norm(xp[i] - xp[1]) * C.BOHR for i in (2, 3)]) - 1.024).max() < 0.005
    assert abs(_angle(xp[2], xp[1], xp[3]) - 103.4) < 0.5
    assert 8.0 < ep - er < 11.0
    # the saddle-region start geometry lies above the reactants (the surface has a late barrier near the product energy)
    ets = energy(oracle, "clnh3", C.SYSTEMS["clnh3"]["ts"]()) * KCAL
    assert ets > er + 5.0


def test_nh3oh_fragments_and_reaction_energy(oracle):
    """known answers for the constants egrad_nh3oh.f adds (BLOCK DATA :2081-2131): water (r0hhp = 0.9595 A, angh2oeq =
    103.6 deg; experiment 0.957 A, 104.5 deg), hydroxyl (r0hhr = 0.971 A; experiment 0.970 A), and the classical
    exothermicity d1ch - d1hh = -10.0 kcal/mol (experiment: D_0(H-NH2) - D_0(H-OH) = 106.7 - 117.6 = -10.9, about -9.5
    after zero-point energy)."""
    xr, er = _relax(oracle, "nh3oh", _far_reactants("nh3oh"), [0, 2, 3, 5])
    assert abs(np.linalg.norm(xr[5] - xr[4]) * C.BOHR - 0.970) < 0.003
    assert np.abs(np.array([np.linalg.norm(xr[i] - xr[1]) * C.BOHR for i in (0, 2, 3)]) - 1.012).max() < 0.005
    xp, ep = _relax(oracle, "nh3oh", _far_products("nh3oh"), [0, 2, 3, 5])
    # the blend of r0hh (:414-424) is driven by the SPECTATOR O-H length alone, which stays near w4 = 0.973 A in both
    # channels: P2 = 1 - tanh(rno - 0.973) ~ 1, so the water of this surface keeps the hydroxyl length r0hhr = 0.971 A
    # (the product value r0hhp = 0.9595 A is reached only if the spectator bond is stretched)
    roh = [np.linalg.norm(xp[i] - xp[4]) * C.BOHR for i in (0, 5)]
    assert np.abs(np.array(roh) - 0.971).max() < 0.002
    assert abs(_angle(xp[0], xp[4], xp[5]) - 104.5) < 1.5
    assert abs(_angle(xp[2], xp[1], xp[3]) - 103.4) < 0.5
    assert -11.5 < ep - er < -8.5
