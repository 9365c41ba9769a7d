"""Pins the QMDFF oracle (ff_eg + ff_nonb restatement) by finite differences and invariances; the
reference ships no expected outputs for this path either."""
import numpy as np
import pytest

from tests.qmdff_synth import make_system


def fd(f, x, a, d, h=1e-5):
    xp, xm = x.copy(), x.copy()
    xp[a, d] += h
    xm[a, d] -= h
    return (f(xp) - f(xm)) / (2 * h)


@pytest.mark.parametrize("periodic", [True, False])
def test_bonded_gradient_is_consistent(oracle, periodic):
    T = make_system(nmol=6, seed=3, periodic=periodic)
    Q = oracle.Qmdff(T)
    x = T["xyz"] + np.random.default_rng(1).normal(0, 0.05, T["xyz"].shape)
    e, g = Q.ff_eg(x)
    rng = np.random.default_rng(2)
    for _ in range(25):
        a, d = int(rng.integers(0, T["n"])), int(rng.integers(0, 3))
        assert abs(fd(lambda y: Q.ff_eg(y)[0], x, a, d) - g[a, d]) < 2e-9


def test_nonbonded_gradient_plain_coulomb_is_consistent(oracle):
    # dispersion/repulsion and the plain (non-periodic) Coulomb are true gradients; the Zahn and
    # switched forms reuse e0/r^2 (ff_nonb.f90:384,470) and are NOT -- reproduced, not fixed
    T = make_system(nmol=6, seed=4, periodic=False)
    Q = oracle.Qmdff(T)
    x = T["xyz"] + np.random.default_rng(1).normal(0, 0.05, T["xyz"].shape)
    V, g = Q.egrad(x)
    rng = np.random.default_rng(5)
    for _ in range(20):
        a, d = int(rng.integers(0, T["n"])), int(rng.integers(0, 3))
        assert abs(fd(lambda y: Q.egrad(y)[0][0], x, a, d) - g[0, a, d]) < 5e-9


@pytest.mark.parametrize("periodic,zahn", [(True, True), (True, False), (False, False)])
def test_invariances(oracle, periodic, zahn):
    T = make_system(nmol=8, seed=6, periodic=periodic, zahn=zahn)
    Q = oracle.Qmdff(T)
    x = T["xyz"] + np.random.default_rng(1).normal(0, 0.05, T["xyz"].shape)
    V, g = Q.egrad(x)
    assert np.abs(g[0].sum(axis=0)).max() < 1e-12 * T["n"]          # every term is pairwise / internal
    shift = np.array([0.3, -1.1, 2.2])
    V2, g2 = Q.egrad(x + shift)
    assert abs(V2[0] - V[0]) < 1e-11 and np.abs(g2 - g).max() < 1e-11
    if periodic:                                                    # lattice translation of one molecule
        y = x.copy()
        y[T["molnum"] == 3] += np.array([T["box"][0], 0, -T["box"][2]])
        V3, g3 = Q.egrad(y)
        assert abs(V3[0] - V[0]) < 1e-10 and np.abs(g3 - g).max() < 1e-10
