"""CPU tests of the CH4 + OH, GeH4 + OH and CH4 + CN oracles (oracle/pes_ch4oh.c, pes_geh4oh.c, pes_ch4cn.c <- egrad_ch4oh.f,
egrad_geh4oh.f, egrad_ch4cn.f;
SURVEY.md 8f row N4).  The reference ships no
outputs for this surface; what it does ship is the start structure of its own saddle search
(examples/explore/ts_irc_ch4oh/ts_start.xyz), used here as the geometry the clouds are drawn around."""
import numpy as np
import pytest

from tests import common as C

KCAL = 627.509474


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


NAMES = ["ch4oh", "geh4oh", "ch4cn"]


def fd_gradient(O, q, h=1e-4, name="ch4oh"):
    g = np.zeros_like(q)
    for a in range(q.shape[0]):
        for d in range(3):
            qp, qm = q.copy(), q.copy()
            qp[a, d] += h
            qm[a, d] -= h
            g[a, d] = (O.egrad(name, qp[None])[0][0] - O.egrad(name, qm[None])[0][0]) / (2 * h)
    return g


@pytest.mark.parametrize("name", NAMES)
def test_gradient_is_the_derivative_of_the_energy(oracle, name):
    """central differences; the floor is the reference's own 2e-6 mismatch between its energy and gradient unit
    factors (0.03812 * 0.52918 against 0.0201723, egrad_ch4oh.f:267,:276), as on the CH4 + H surface"""
    q = C.ts_cloud(name, 6, 0.1, np.random.default_rng(3))
    V, g, info = oracle.egrad(name, q)
    assert info == 0 and np.isfinite(V).all()
    for im in range(len(q)):
        gn = fd_gradient(oracle, q[im], name=name)
        assert np.abs(gn - g[im].reshape(7, 3)).max() < 1e-5 * np.abs(g[im]).max()


@pytest.mark.parametrize("name", NAMES)
def test_each_added_term_has_a_consistent_gradient(oracle, name):
    """the three energy parts (stretch incl. the O-H Morse bond, out-of-plane, in-plane incl. the H-O-H bends) move
    when the atoms they depend on move: H(O) enters only through the added terms, so its finite-difference force
    checks them in isolation"""
    q = C.ts_cloud(name, 4, 0.1, np.random.default_rng(5))
    _, g, _ = oracle.egrad(name, q)
    for im in range(len(q)):
        gn = fd_gradient(oracle, q[im], name=name)
        assert np.abs(g[im].reshape(7, 3)[6]).max() > 1e-4            # the added terms act on H(O)
        assert np.abs(gn[6] - g[im].reshape(7, 3)[6]).max() < 1e-5 * np.abs(g[im]).max()


@pytest.mark.parametrize("name", NAMES)
def test_invariances(oracle, name):
    rng = np.random.default_rng(7)
    q = C.ts_cloud(name, 50, 0.15, rng)
    V, g, _ = oracle.egrad(name, q)
    g = g.reshape(q.shape)
    assert np.abs(g.sum(axis=1)).max() < 1e-12                         # no net force
    assert np.abs(np.cross(q, g).sum(axis=1)).max() < 1e-10            # no net torque
    A = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    V2, g2, _ = oracle.egrad(name, q @ A.T + 1.5)
    assert np.abs(V2 - V).max() < 1e-11
    assert np.abs(g2.reshape(q.shape) - g @ A.T).max() < 1e-10
    # the four methane hydrogens are equivalent (atoms 1, 3, 4, 5)
    for perm in ([3, 1, 2, 0, 4, 5, 6], [0, 1, 4, 2, 3, 5, 6], [2, 1, 0, 4, 3, 5, 6]):
        V3, g3, _ = oracle.egrad(name, q[:, perm])
        assert np.abs(V3 - V).max() < 1e-10
        assert np.abs(g3.reshape(q.shape) - g[:, perm]).max() < 1e-9


def test_shipped_saddle_search_start_is_close_to_stationary(oracle):
    """ts_start.xyz is where the reference starts `job opt_ts` on this surface: forces there are a small fraction of
    those a 0.1 bohr displacement produces"""
    V0, g0, _ = oracle.egrad("ch4oh", C.ch4oh_ts()[None])
    _, g, _ = oracle.egrad("ch4oh", C.ts_cloud("ch4oh", 20, 0.1, np.random.default_rng(1)))
    assert np.abs(g0).max() < 0.01
    assert np.abs(g0).max() < 0.2 * np.median(np.abs(g).max(axis=(1,) if g.ndim == 2 else (1, 2)))


def water_far_from_xh3(name, rxh, dth=0.0, scale=1.0):
    """products far apart (XH3 ... H2O, 12 A), the water at the BLOCK DATA constants times `scale`, bend opened by dth"""
    r, th = scale * 0.9706 / C.BOHR, np.deg2rad(104.7132) + dth
    t = np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]]) / np.sqrt(3)
    q = np.zeros((7, 3))
    q[2], q[3], q[4] = t[1] * rxh, t[2] * rxh, t[3] * rxh              # CH3 / GeH3
    q[5] = t[0] * 12.0 / C.BOHR                                        # O
    q[0] = q[5] - t[0] * r                                             # H taken from XH4, now on the oxygen
    e2 = t[1] - (t[1] @ t[0]) * t[0]
    e2 /= np.linalg.norm(e2)
    q[6] = q[5] + r * (np.cos(th) * (-t[0]) + np.sin(th) * e2)
    return q


def test_water_fragment_relaxes_to_the_constants_of_the_block_data(oracle):
    """CH3 ... H2O: the water O-H bonds relax to r0hh = 0.9706 A and the bend to anh2oeq = 104.7132 deg
    (egrad_ch4oh.f:2078,:2103) -- the added Morse bond and bends carry exactly these minima"""
    from scipy.optimize import minimize
    q = water_far_from_xh3("ch4oh", 2.05, dth=0.1, scale=1.05)

    def f(x):
        V, g, _ = oracle.egrad("ch4oh", x.reshape(1, 7, 3))
        return V[0], g.reshape(-1)
    res = minimize(f, q.reshape(-1), jac=True, method="BFGS", options=dict(gtol=1e-7, maxiter=2000))
    x = res.x.reshape(7, 3)
    a, b = x[0] - x[5], x[6] - x[5]
    assert abs(np.linalg.norm(a) * C.BOHR - 0.9706) < 2e-3
    assert abs(np.linalg.norm(b) * C.BOHR - 0.9706) < 2e-3
    ang = np.degrees(np.arccos(a @ b / np.linalg.norm(a) / np.linalg.norm(b)))
    assert abs(ang - 104.7132) < 0.3


@pytest.mark.parametrize("name,rxh", [("ch4oh", 2.05), ("geh4oh", 2.88)])
def test_bend_force_constant_is_the_water_value(oracle, name, rxh):
    """known answer: fkh2oeq = 0.73 mdyn A / rad^2 (egrad_ch4oh.f:2101, egrad_geh4oh.f:2037), the experimental water
    bending constant.  PREPOT scales it by fact2 to 1e5 J/mol (:2001) and POT by 0.03812 to hartree:
    0.73e-18 J / 4.35974e-18 J = 0.1674 Eh / rad^2.  (A restatement without the fact2 line is off by a factor 6.)"""
    d = 1e-2
    e = [oracle.egrad(name, water_far_from_xh3(name, rxh, dth=x)[None])[0][0] for x in (-d, 0.0, d)]
    k = (e[0] - 2 * e[1] + e[2]) / d ** 2
    assert abs(k - 0.73e-18 / 4.35974e-18) < 0.005 * 0.1674
    # and the O-H(O) Morse bond has its minimum at r0hh with the depth d1hh (:2079 / :2015)
    x = water_far_from_xh3(name, rxh)
    _, g, _ = oracle.egrad(name, x[None])
    assert np.abs(g.reshape(7, 3)[6]).max() < 2e-5     # 0.52918 (POT) against 0.52917721 (the test) in r0hh


def hcn_far_from_ch3(rch=1.06497, rcn=1.172, bend=0.0):
    """products far apart (CH3 ... HCN, 12 A): H on the carbon of CN at rch, N at rcn, H-C-N opened by `bend` from 180 deg"""
    t = np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]]) / np.sqrt(3)
    q = np.zeros((7, 3))
    q[2], q[3], q[4] = t[1] * 2.05, t[2] * 2.05, t[3] * 2.05           # CH3
    q[5] = t[0] * 12.0 / C.BOHR                                        # C of CN
    q[0] = q[5] - t[0] * rch / C.BOHR                                  # H taken from CH4, now on that carbon
    e2 = t[1] - (t[1] @ t[0]) * t[0]
    e2 /= np.linalg.norm(e2)
    q[6] = q[5] + rcn / C.BOHR * (np.cos(bend) * t[0] + np.sin(bend) * e2)
    return q


def test_ch4cn_product_fragment_is_hydrogen_cyanide(oracle):
    """known answers for the constants egrad_ch4cn.f adds to the CH4 + OH template (BLOCK DATA :2084-2087, :2109-2111 and
    the literals of :625-627): far from CH3 the H-C-N fragment relaxes to the experimental HCN geometry the constants
    encode -- r(C-H) = r0hh = 1.06497 A (experiment 1.0655), r(C-N) = 1.172 A (the CN radical's 1.1718), linear -- and the
    well of the H-CN bond is d1hh = 132.17 kcal/mol deep"""
    from scipy.optimize import minimize
    q = hcn_far_from_ch3(1.10, 1.20, 0.15)

    def f(x):
        V, g, _ = oracle.egrad("ch4cn", x.reshape(1, 7, 3))
        return V[0], g.reshape(-1)
    res = minimize(f, q.reshape(-1), jac=True, method="BFGS", options=dict(gtol=1e-7, maxiter=3000))
    x = res.x.reshape(7, 3)
    a, b = x[0] - x[5], x[6] - x[5]
    assert abs(np.linalg.norm(a) * C.BOHR - 1.06497) < 2e-3
    assert abs(np.linalg.norm(b) * C.BOHR - 1.172) < 2e-3
    ang = np.degrees(np.arccos(a @ b / np.linalg.norm(a) / np.linalg.norm(b)))
    assert ang > 179.0
    # depth of the H-CN well: pull the hydrogen off along the axis
    far = x.copy()
    far[0] = x[5] + (x[0] - x[5]) / np.linalg.norm(a) * 14.0
    dE = (oracle.egrad("ch4cn", far[None])[0][0] - res.fun) * KCAL
    assert abs(dE - 132.17) < 1.5
