"""Pins the DG-EVB oracle (gradient.f90:365-537 + sum_v12 / sum_dv12 / xyz_2int / numeric Wilson B and
the *_two force field) by finite differences and limiting cases; the reference ships no expected
outputs for this path and its evb_pars.dat needs evbopt.x (not runnable here)."""
import numpy as np
import pytest

from tests.qmdff_synth import make_dgevb


def _fd(f, x, a, d, h=1e-5):
    xp, xm = x.copy(), x.copy()
    xp[a, d] += h
    xm[a, d] -= h
    return (f(xp) - f(xm)) / (2 * h)


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_gradient_is_consistent_with_energy(oracle, mode):
    # sum_dv12 is the analytic derivative of sum_v12 and B is a central difference with shift 1e-3
    # (calc_wilson.f90:114-178): relative error of B ~ 1e-6 * curvature, so the bound is loose
    T1, T2, E = make_dgevb(seed=1, mode=mode, npoints=5)
    D = oracle.Dgevb(T1, T2, E)
    x = T1["xyz"] + np.random.default_rng(2).normal(0, 0.04, T1["xyz"].shape)
    V, g = D.egrad(x)
    rng = np.random.default_rng(3)
    for _ in range(12):
        a, d = int(rng.integers(0, T1["n"])), int(rng.integers(0, 3))
        assert abs(_fd(lambda y: D.egrad(y)[0][0], x, a, d) - g[0, a, d]) < 5e-8


def test_zero_coupling_gives_lower_diabat(oracle):
    T1, T2, E = make_dgevb(seed=2, mode=2, npoints=4)
    E = dict(E, b_vec=np.zeros_like(E["b_vec"]))
    D = oracle.Dgevb(T1, T2, E)
    x = T1["xyz"] + np.random.default_rng(5).normal(0, 0.04, T1["xyz"].shape)
    V, g = D.egrad(x)
    e1, g1 = oracle.Qmdff(T1).egrad(x)
    e2, g2 = D.second_state(x)
    lo = 0 if e1[0] < e2 else 1
    assert abs(V[0] - min(e1[0], e2)) < 1e-13
    assert np.abs(g[0] - (g1[0], g2)[lo]).max() < 1e-12


def test_coupling_lowers_the_energy_and_gaussians_are_thresholded(oracle):
    T1, T2, E = make_dgevb(seed=3, mode=1, npoints=3)
    E = dict(E, b_vec=np.abs(E["b_vec"]) + 1e-4)
    D = oracle.Dgevb(T1, T2, E)
    x = T1["xyz"] + np.random.default_rng(6).normal(0, 0.02, T1["xyz"].shape)
    V, _ = D.egrad(x)
    e1, _ = oracle.Qmdff(T1).egrad(x)
    e2, _ = D.second_state(x)
    assert V[0] < min(e1[0], e2) - 1e-6
    # g_thres: with a huge exponent every Gaussian is skipped (sum_v12.f90 `cycle`) -> no coupling
    D2 = oracle.Dgevb(T1, T2, dict(E, alph=np.full(3, 1e6)))
    V2, _ = D2.egrad(x)
    assert abs(V2[0] - min(e1[0], e2)) < 1e-13


def test_second_state_has_no_cutoff_and_ignores_the_box(oracle):
    # ff_nonb_two.f90:74: Coulomb q_i q_j eps1 / r for every nci pair; never periodic
    T1, T2, E = make_dgevb(seed=4, mode=1, npoints=3)
    D = oracle.Dgevb(T1, T2, E)
    x = T1["xyz"].copy()
    e_a, _ = D.second_state(x)
    e_b, _ = D.second_state(x + np.array([50.0, -20.0, 3.0]))
    assert abs(e_a - e_b) < 1e-11


def test_internal_coordinates_known_values(oracle):
    T1, T2, E = make_dgevb(seed=1, mode=1, npoints=2)
    cd = np.array([[1, 1, 2, 0, 0], [2, 1, 2, 3, 0], [3, 1, 2, 3, 4], [4, 1, 2, 3, 4]], dtype=np.int32)
    D = oracle.Dgevb(T1, T2, dict(E, coord_def=cd, point_int=np.zeros((2, 4))))
    x = np.zeros_like(T1["xyz"])
    x[:4] = [[1.0, 0, 0], [0, 0, 0], [0, 2.0, 0], [0, 2.0, 1.5]]
    q = D.internals(x)
    assert abs(q[0] - 1.0) < 1e-15 and abs(q[1] - np.pi / 2) < 1e-15 and abs(q[2] - np.pi / 2) < 1e-15
    # oop.f90: v41.(v42 x v43) / |v41 x v42 + v42 x v43 + v43 x v41| on unit vectors
    v = [(x[3] - x[i]) / np.linalg.norm(x[3] - x[i]) for i in range(3)]
    nv = np.cross(v[0], v[1]) + np.cross(v[1], v[2]) + np.cross(v[2], v[0])
    assert abs(q[3] - v[0].dot(np.cross(v[1], v[2])) / np.linalg.norm(nv)) < 1e-15
