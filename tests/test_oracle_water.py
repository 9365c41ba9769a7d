"""CPU checks of the flexible SPC water restatement (oracle/water.c <- egrad_water.f90, water_init.f90).  No vectors
exist in the reference for this model either (parity unpinned); pinned by the monomer minimum its parameters encode,
finite differences where the source's gradient is a true gradient, invariances, and the product's own independent
parameter set-up (caracal_b200/water.py)."""
import math

import numpy as np
import pytest

from caracal_b200 import water as WT


def test_parameter_setup_agrees_with_the_products_mirror(oracle):
    assert np.array_equal(oracle.water_default_pars(), WT.water_pars())
    p = WT.water_pars()
    assert abs(p[0] * WT.BOHR - 1.0) < 1e-15 and abs(p[2] * WT.HARTREE - 101.9188) < 1e-12
    assert p[6] != 111.70765 / WT.HARTREE * WT.BOHR ** 2          # the REAL*4 literal is visible at 1e-8 relative


def test_monomer_minimum(oracle):
    W = WT.water_box(1)
    p = W["pars"]
    half = math.asin(p[1] / 2 / p[0])
    x = np.array([[0, 0, 0], [p[0] * math.sin(half), p[0] * math.cos(half), 0], [-p[0] * math.sin(half), p[0] * math.cos(half), 0]])
    V, g = oracle.Water(W).egrad(x)
    assert abs(2 * math.degrees(half) - 109.47) < 0.01
    assert V[0] == 0.0 and np.abs(g).max() == 0.0
    V2, g2 = oracle.Water(W).egrad(x + np.random.default_rng(0).normal(0, 0.05, x.shape))
    assert V2[0] > 0


@pytest.mark.parametrize("nwater", [2, 8])
def test_gradient_of_the_gas_phase_cluster_is_consistent(oracle, nwater):
    rng = np.random.default_rng(nwater)
    W = WT.water_box(nwater)
    x = WT.water_lattice(nwater, 3.2 * math.ceil(nwater ** (1 / 3)), rng, jitter=0.08)
    Q = oracle.Water(W)
    V, g = Q.egrad(x)
    h = 1e-5
    for _ in range(12):
        a, d = int(rng.integers(0, 3 * nwater)), int(rng.integers(0, 3))
        xp, xm = x.copy(), x.copy()
        xp[a, d] += h
        xm[a, d] -= h
        assert abs((Q.egrad(xp)[0][0] - Q.egrad(xm)[0][0]) / (2 * h) - g[0, a, d]) < 5e-9
    assert np.abs(g[0].sum(axis=0)).max() < 1e-13
    A = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    V2, g2 = Q.egrad(x @ A.T + 2.0)
    assert abs(V2[0] - V[0]) < 1e-12 and np.abs(g2[0] - g[0] @ A.T).max() < 1e-12


def test_lennard_jones_applies_to_every_pair_led_by_an_oxygen(oracle):
    """egrad_water.f90:295 tests name(i) twice: O(i)-H(j) pairs of different molecules carry the O-O term too."""
    W = WT.water_box(2)
    x = WT.water_lattice(2, 6.4, np.random.default_rng(1), jitter=0.0)
    Q = oracle.Water(W)
    V, _ = Q.egrad(x)
    W0 = dict(W, pars=W["pars"].copy())
    W0["pars"][10] = 0.0                                           # no Lennard-Jones at all
    V0, _ = oracle.Water(W0).egrad(x)
    p, oner = W["pars"], lambda i, j: 1.0 / np.linalg.norm(x[i] - x[j])   # noqa: E731
    lj = lambda i, j: 4 * p[10] * ((p[9] * oner(i, j)) ** 12 - (p[9] * oner(i, j)) ** 6)   # noqa: E731
    assert abs((V[0] - V0[0]) - (lj(0, 3) + lj(0, 4) + lj(0, 5))) < 1e-14


def test_periodic_box_invariances(oracle):
    rng = np.random.default_rng(5)
    nw, L = 64, 12.6
    W = WT.water_box(nw, periodic_angstrom=[L, L, L])
    assert W["zahn"] == 1 and abs(W["coul_cut"] - (0.5 * L / WT.BOHR - 0.1)) < 1e-12   # clamped to half the box
    x = WT.water_lattice(nw, L, rng)
    Q = oracle.Water(W)
    V, g = Q.egrad(x)
    assert np.isfinite(V).all() and np.abs(g[0].sum(axis=0)).max() < 1e-12
    y = x.copy()
    y[3 * 7:3 * 8] += np.array([W["box"][0], 0.0, -W["box"][2]])       # lattice translation of one molecule
    V2, g2 = Q.egrad(y + 0.7)
    assert abs(V2[0] - V[0]) < 1e-11 and np.abs(g2 - g).max() < 1e-11
