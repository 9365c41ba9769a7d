"""CPU tests of the H2CO oracle (oracle/pes_h2co.c <- main_h2co.f90).  The reference ships no outputs for
this surface; the restatement is pinned by the tables re-read from the source text, by the central-difference loop redone
outside it (bit for bit), by the symmetry of the fit, and by known answers from experiment: the fit's zero of energy is its
formaldehyde minimum, whose geometry is the experimental one; far along the molecular channel the fragments relax to H2
and CO."""
import ctypes
import os

import numpy as np
import pytest
from scipy.optimize import minimize

from tests import common as C

KCAL = 627.509474
dp = ctypes.POINTER(ctypes.c_double)
SRC = "/root/reference/src/main_h2co.f90"


def _d(a):
    return a.ctypes.data_as(dp)


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as O
    O.build()
    O.lib().oracle_h2co_energy.argtypes = [dp, dp, ctypes.POINTER(ctypes.c_int)]
    return O


def energy(O, q):
    q = np.ascontiguousarray(q, dtype=np.float64).reshape(12)
    v, far = np.zeros(1), ctypes.c_int(0)
    O.lib().oracle_h2co_energy(_d(q), _d(v), ctypes.byref(far))
    return v[0], far.value


def h2co_geometry(rco=1.205, rch=1.111, hch=116.1):
    th = np.deg2rad(hch / 2)
    return np.array([[0, 0, 0], [0, 0, rco], [rch * np.sin(th), 0, -rch * np.cos(th)],
                     [-rch * np.sin(th), 0, -rch * np.cos(th)]]) / C.BOHR


@pytest.mark.skipif(not os.path.exists(SRC), reason="the reference source is only in the build container")
def test_tables_against_the_source_text():
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(os.path.dirname(__file__), "..", "oracle", "make_h2co_tables.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    cof, pw = mk.parse(SRC)
    for name in ("oracle/h2co_tables.h", "caracal_b200/csrc/pes_h2co_tables.cuh"):
        txt = open(os.path.join(os.path.dirname(__file__), "..", name)).read()
        import re
        got = re.findall(r"F32\(([^)]*)\)", txt)
        assert got == cof
        rows = re.findall(r"\{(\d+), (\d+), (\d+), (\d+), (\d+), (\d+)\}", txt)
        assert [[int(t) for t in r] for r in rows] == pw
    # the symmetrised basis makes (p2..p5) and its mirror the same term: no row may list its own mirror image twice
    assert len({tuple(p) for p in pw}) == 1561


def test_zero_of_energy_is_the_formaldehyde_minimum(oracle):
    """experiment: r(CO) = 1.205 A, r(CH) = 1.111 A, H-C-H = 116.1 deg; the fit adds 114.3329... Eh so that its minimum is 0"""
    x0 = h2co_geometry()
    e0, far = energy(oracle, x0)
    assert far == 0 and abs(e0) * KCAL < 0.2
    res = minimize(lambda v: energy(oracle, v)[0] * KCAL, x0.ravel(), method="BFGS", options=dict(gtol=1e-6))
    x = res.x.reshape(4, 3)
    assert abs(res.fun) < 0.01                                           # kcal/mol
    assert abs(np.linalg.norm(x[1] - x[0]) * C.BOHR - 1.205) < 0.01
    assert abs(np.linalg.norm(x[2] - x[0]) * C.BOHR - 1.111) < 0.015 and abs(np.linalg.norm(x[3] - x[0]) * C.BOHR - 1.111) < 0.015
    u, v = x[2] - x[0], x[3] - x[0]
    assert abs(np.degrees(np.arccos(u @ v / np.linalg.norm(u) / np.linalg.norm(v))) - 116.1) < 1.5


def test_molecular_channel_fragments_are_h2_and_co(oracle):
    """H2 ... CO 6 A apart (r(H-H) stays far below the 8 bohr switch): the fragments relax to r(H-H) = 0.741 A and
    r(C-O) = 1.128 A (experiment), a few kcal/mol from the formaldehyde minimum (the reaction is nearly thermoneutral)"""
    x0 = np.array([[0, 0, 0], [0, 0, 1.13], [0.37, 0, -6.0], [-0.37, 0, -6.0]]) / C.BOHR

    def f(v):
        x = x0.copy()
        x[1, 2] = v[0]
        x[2, 0], x[3, 0] = v[1], -v[1]
        return energy(oracle, x)[0] * KCAL
    res = minimize(f, [x0[1, 2], x0[2, 0]], method="Nelder-Mead", options=dict(xatol=1e-5, fatol=1e-7))
    assert abs(res.x[0] * C.BOHR - 1.128) < 0.01
    assert abs(2 * res.x[1] * C.BOHR - 0.741) < 0.01
    assert -5.0 < res.fun < 12.0


def test_gradient_is_the_in_place_central_difference_loop(oracle):
    """egrad_h2co :3231-3304 redone outside the restatement from its energy entry: x + h, (x + h) - 2h, ((x + h) - 2h) + h
    in place, coordinate after coordinate -- identical bits"""
    h = 0.001
    for x in C.ts_cloud("h2co", 4, 0.1, np.random.default_rng(3)):
        V, g, info = oracle.egrad("h2co", x[None])
        c = x.reshape(12).copy()
        grad = np.zeros(12)
        assert V[0] == energy(oracle, c)[0]
        for i in range(12):
            c[i] = c[i] + h
            eu = energy(oracle, c)[0]
            c[i] = c[i] - 2.0 * h
            el = energy(oracle, c)[0]
            c[i] = c[i] + h
            grad[i] = (eu - el) / (2.0 * h)
        assert np.array_equal(grad, g.reshape(12))


def test_hydrogens_permute_and_rigid_motions(oracle):
    rng = np.random.default_rng(7)
    q = C.ts_cloud("h2co", 6, 0.15, rng)
    V, g, _ = oracle.egrad("h2co", q)
    Vp, gp, _ = oracle.egrad("h2co", q[:, [0, 1, 3, 2]])
    assert np.abs(V - Vp).max() < 1e-12                       # the two orders of the symmetrised product differ by rounding
    assert np.abs(g[:, [0, 1, 3, 2]] - gp).max() < 1e-8       # difference quotient of step 1e-3: 1e-13 / 2e-3
    A, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    V2, g2, _ = oracle.egrad("h2co", q @ A.T + rng.normal(size=3))
    assert np.abs(V - V2).max() < 1e-11
    assert np.abs(g @ A.T - g2).max() < 1e-5                  # O(h^2) truncation of a central difference along rotated axes


def test_radical_channel_is_flagged(oracle):
    """r(H-H) >= 8 bohr: the reference calls hcopot, which opens a parameter file it does not ship; info = 1"""
    x = h2co_geometry()
    x[3] = x[0] + np.array([-9.0, 0.0, -2.0])
    V, g, info = oracle.egrad("h2co", x[None])
    assert info == 1 and np.isfinite(V).all()
    assert oracle.egrad("h2co", h2co_geometry()[None])[2] == 0


def test_single_precision_literals_matter_at_the_microhartree_level(oracle):
    """the 1561 coefficients are REAL*4 literals in the source (no D exponent): the surface as compiled differs
    from the published fit by ~1e-6 Eh, and the oracle's default mode is the compiled one"""
    x = h2co_geometry()
    a = oracle.egrad("h2co", x[None])[0][0]
    b = oracle.egrad("h2co", x[None], exact=True)[0][0]
    assert 1e-8 < abs(a - b) < 1e-4
