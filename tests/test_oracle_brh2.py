"""CPU checks of the BrH2 DIM-3C restatement (oracle/pes_brh2.c <- egrad_brh2.f).
The reference ships no vectors for this surface either (parity unpinned); the restatement is pinned by the
properties the source's own header states (zero of energy, asymptotic valleys EASYAB/BC/AC :146-148),
finite differences, the energy-difference / integrated-gradient identity, symmetry, and (on the GPU) the device
kernel's independent Jacobi diagonalisation of the same Hamiltonian."""
import numpy as np

from oracle import oracle as O
from tests import common as C

DHH, DHX, RHH, RHX = 0.17447, 0.1439, 1.4016, 2.673   # BLOCK DATA PTPARM, egrad_brh2.f:465-483


def pot(R):
    L = O.lib()
    R = np.ascontiguousarray(R, dtype=np.float64)
    V, d, ie = np.zeros(1), np.zeros(3), np.zeros(1, dtype=np.int32)
    L.oracle_brh2_pot(O._d(R), O._d(V), O._d(d), O._i(ie))
    assert ie[0] == 0
    return V[0], d


def test_zero_of_energy_and_valleys():
    # "zero: Br infinitely far from H2 at its equilibrium distance" (:120-123); Rin = (r12, r13, r23), atom 2 = Br
    V, d = pot([40.0, RHH, 40.3])
    assert abs(V) < 1e-12 and np.abs(d).max() < 1e-12
    # H far from HBr(re): EASYAB - EASYBC = DHH - DHX above the zero (:146-148 with the +DHH shift of :422)
    V, d = pot([RHX, 40.0, 41.0])
    assert abs(V - (DHH - DHX)) < 1e-12 and np.abs(d).max() < 1e-12
    V2, _ = pot([41.0, 40.0, RHX])
    assert abs(V2 - V) < 1e-14


def test_cartesian_gradient_is_the_derivative_of_the_energy():
    rng = np.random.default_rng(3)
    q = C.ts_cloud("brh2", 40, 0.25, rng)
    # keep clear of the collinear switch LCOL (sin^2 A < 1e-6, :217), where the surface is discontinuous
    a, b = q[:, 1] - q[:, 0], q[:, 2] - q[:, 0]
    sin2 = 1 - (np.einsum("ij,ij->i", a, b) / np.linalg.norm(a, axis=1) / np.linalg.norm(b, axis=1)) ** 2
    q = q[sin2 > 1e-3]
    assert len(q) > 20
    V, g, info = O.egrad("brh2", q)
    assert info == 0
    h = 1e-5
    for c in range(9):
        dq = np.zeros(9)
        dq[c] = h
        Vp = O.egrad("brh2", q + dq.reshape(3, 3))[0]
        Vm = O.egrad("brh2", q - dq.reshape(3, 3))[0]
        assert np.abs((Vp - Vm) / (2 * h) - g.reshape(-1, 9)[:, c]).max() < 2e-9


def test_symmetry_and_rigid_motions():
    rng = np.random.default_rng(4)
    q = C.ts_cloud("brh2", 200, 0.3, rng)
    V, g, _ = O.egrad("brh2", q)
    V2, g2, _ = O.egrad("brh2", q[:, [2, 1, 0]])        # the two hydrogens are equivalent
    assert np.abs(V2 - V).max() < 1e-12 and np.abs(g2 - g[:, [2, 1, 0]]).max() < 1e-11
    A = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    V3, g3, _ = O.egrad("brh2", q @ A.T + 1.7)
    assert np.abs(V3 - V).max() < 1e-12 and np.abs(g3 - g @ A.T).max() < 1e-11
    assert np.abs(g.sum(axis=1)).max() < 1e-13


def test_saddle_point_of_the_shared_ts_structure():
    ts = C.brh2_ts()
    V, g, _ = O.egrad("brh2", ts[None])
    assert np.abs(g).max() < 1e-7                                  # stationary
    assert abs(V[0] * 627.5095 - 20.996) < 2e-3                    # 21.0 kcal/mol above Br + H2
    # exactly one negative curvature along the collinear antisymmetric stretch
    def g2(x):
        return np.array(pot([x[0], x[1], x[0] + x[1]])[1]) @ np.array([[1, 0], [0, 1], [1, 1]])
    x0, h = np.array([2.72158888, 2.64056088]), 1e-4
    H = np.array([(g2(x0 + h * e) - g2(x0 - h * e)) / (2 * h) for e in np.eye(2)])
    w = np.linalg.eigvalsh(0.5 * (H + H.T))
    assert w[0] < -5e-3 and w[1] > 0.1


def test_gradient_integrates_to_the_energy_difference():
    """The Hellmann-Feynman derivative of the lowest root from the TRED3/TQL2/TRBAK3 chain (whose QL sweep stops at
    a 2**-37 off-diagonal threshold, :810) is accurate far beyond the 1e-10 parity tolerance: a 5-point
    Gauss-Legendre integral of g.dq over a short segment reproduces the energy difference to 1e-13 Eh.  (The
    device kernel diagonalises with an independent Jacobi iteration; tests/test_gpu_pes.py compares the two.)"""
    rng = np.random.default_rng(5)
    for _ in range(10):
        q = C.ts_cloud("brh2", 1, 0.3, rng)[0]
        dq = rng.normal(0, 1.0, (3, 3))
        dq /= np.linalg.norm(dq)
        # 5-point Gauss-Legendre integral of g.dq over a 0.02 bohr segment vs the energy difference
        x, w = np.polynomial.legendre.leggauss(5)
        s = 0.01
        qs = np.array([q + s * xi * dq for xi in x])
        _, g, _ = O.egrad("brh2", qs)
        integral = s * np.sum(w * np.einsum("kij,ij->k", g, dq))
        Va = O.egrad("brh2", (q - s * dq)[None])[0][0]
        Vb = O.egrad("brh2", (q + s * dq)[None])[0][0]
        assert abs((Vb - Va) - integral) < 1e-13
