"""Pins the SPME oracle (ewald_recip.f90 + set_periodic.f90:114-231 restatement).  The reference cannot
reach this routine (ff_nonb.f90:337) and ships no expected output for it, so the anchor is
the plain Ewald reciprocal-space sum, within the accuracy SPME has at 1.2 grid points per Angstrom and
order-5 splines, plus exact structural properties."""
import numpy as np
import pytest

from caracal_b200.ewald import ewald_setup


def neutral(rng, n, L):
    x = rng.uniform(0, 1, (n, 3)) * np.asarray(L)
    q = rng.normal(0, 0.4, n)
    return x, q - q.mean()


@pytest.mark.parametrize("L", [(30.0, 30.0, 30.0), (36.0, 36.0, 36.0)])
def test_spme_matches_plain_ewald(oracle, L):
    E = oracle.Ewald(L)
    x, q = neutral(np.random.default_rng(0), 60, L)
    e, g = E.recip(x, q)
    e2, g2 = E.direct_recip(x, q, 14)
    assert abs(e - e2) < 1e-4 * abs(e2)
    assert np.abs(g - g2).max() < 1e-2 * np.abs(g2).max()
    assert np.sqrt(((g - g2) ** 2).mean()) < 3e-3 * np.sqrt((g2 ** 2).mean())


def test_grid_translation_and_lattice_invariance(oracle):
    L = (30.0, 30.0, 30.0)
    E = oracle.Ewald(L)
    x, q = neutral(np.random.default_rng(1), 40, L)
    e, g = E.recip(x, q)
    h = L[0] / E.nfft
    e2, g2 = E.recip(x + np.array([3 * h, -2 * h, h]), q)          # shift by whole grid spacings
    assert abs(e2 - e) < 1e-12 * abs(e) and np.abs(g2 - g).max() < 1e-12 * np.abs(g).max()
    y = x.copy()
    y[::3] += np.array([L[0], -2 * L[1], 0.0])                      # lattice translations of some atoms
    e3, g3 = E.recip(y, q)
    assert abs(e3 - e) < 1e-11 * abs(e) and np.abs(g3 - g).max() < 1e-10 * np.abs(g).max()


def test_energy_is_quadratic_in_the_charges_and_gradient_consistent(oracle):
    L = (30.0, 30.0, 30.0)
    E = oracle.Ewald(L)
    x, q = neutral(np.random.default_rng(2), 30, L)
    e, g = E.recip(x, q)
    e2, g2 = E.recip(x, 2.0 * q)
    assert abs(e2 - 4 * e) < 1e-12 * abs(e) and np.abs(g2 - 4 * g).max() < 1e-12 * np.abs(g).max()
    # SPME forces are the analytic derivative of the SPME energy up to the missing self-consistency of
    # the interpolation: central differences agree to the SPME force accuracy, not to round-off
    for a, d in ((0, 0), (7, 1), (19, 2)):
        xp, xm = x.copy(), x.copy()
        xp[a, d] += 1e-4
        xm[a, d] -= 1e-4
        fd = (E.recip(xp, q)[0] - E.recip(xm, q)[0]) / 2e-4
        assert abs(fd - g[a, d]) < 5e-3 * np.abs(g).max()


def test_host_setup_equals_oracle_setup(oracle):
    for Lx in (30.0, 58.7, 78.6):
        P = ewald_setup([Lx, Lx, Lx])
        E = oracle.Ewald([Lx, Lx, Lx])
        assert P["nfft"] == E.nfft and abs(P["a_ewald"] - E.a_ewald) < 1e-15
        ny = E.nfft // 2                 # Nyquist: zeta = sum2/sum1 with sum1 ~ 1e-10 by cancellation
        keep = np.arange(E.nfft) != ny
        assert np.abs(P["bsmod"][0][keep] / E.bsmod[0][keep] - 1.0).max() < 1e-12
        assert P["bsmod"][0][ny] > 1e15 and E.bsmod[0][ny] > 1e15
    assert ewald_setup([78.6] * 3)["nfft"] == 50 and ewald_setup([20.0] * 3)["nfft"] == 16
