"""The other umbr_type families of calc_xi.f90: unimolecular CYCLOREVER / REARRANGE /
DECOM_1BOND / ELIMINATION (:673-938) and ATOM_SHIFT (:523-672).  Oracle restatement pinned by finite
differences (gradient and Hessian, both modes) and closed-form values."""
import numpy as np
import pytest

from caracal_b200.api import AtomShiftMechanism, UnimolMechanism
from tests import common as C


def unimol_system(oracle, rng, nb=1):
    """CH4 + H geometry used as a 6-atom 'rearrangement': bond 2-1 breaks, 1-6 forms"""
    ts = C.ch5_ts()
    reac = ts.copy()
    reac[5] += np.array([2.0, 1.0, 0.5])          # reactant: H_b further away
    reac[0] = reac[1] + (ts[0] - ts[1]) * 0.8      # and the C-H' bond shorter
    m = UnimolMechanism([[1, 6]], [[2, 1]], ts, reac)
    s = oracle.System("ch4h", nb, C.masses("ch4h"), C.beta_calc_rate(300.0), C.dt_au(0.1))
    s.set_mechanism(m)
    return s, m, ts


def shift_system(oracle, coord, nb=1):
    m = AtomShiftMechanism(3, coord, -0.7, 1.9, 0.3, 2.6)
    s = oracle.System("h3", nb, C.masses("h3"), C.beta_calc_rate(300.0), C.dt_au(0.1))
    s.set_mechanism(m)
    return s, m


def _fd_check(s, x, xi_ideal):
    for mode in (1, 2):
        xi, dxi, d2 = s.calc_xi(x, xi_ideal, mode, hessian=True)
        for a in range(x.shape[0]):
            for d in range(3):
                xp, xm = x.copy(), x.copy()
                xp[a, d] += 1e-5
                xm[a, d] -= 1e-5
                fp, gp = s.calc_xi(xp, xi_ideal, mode)
                fm, gm = s.calc_xi(xm, xi_ideal, mode)
                assert abs((fp - fm) / 2e-5 - dxi[a, d]) < 1e-8
                assert np.abs((gp - gm) / 2e-5 - d2[a, d]).max() < 1e-7


def test_unimolecular_xi(oracle):
    rng = np.random.default_rng(1)
    s, m, ts = unimol_system(oracle, rng)
    # at the TS structure s1 = 0: umbrella form 1, recrossing form (1 - xi_ideal) * s0
    xi, _ = s.calc_xi(ts, 0.9, 1)
    assert abs(xi - 1.0) < 1e-13
    s0 = (m.break_ref[0] - m.break_reac[0]) - (m.form_ref[0] - m.form_reac[0])
    xi2, _ = s.calc_xi(ts, 0.9, 2)
    assert abs(xi2 - 0.1 * s0) < 1e-13
    _fd_check(s, ts + rng.normal(0, 0.1, ts.shape), 0.7)


@pytest.mark.parametrize("coord", [1, 2, 3, 4, 5, 6])
def test_atom_shift_xi(oracle, coord):
    s, m = shift_system(oracle, coord)
    x = C.h3_ts() + np.random.default_rng(coord).normal(0, 0.2, (3, 3))
    c1 = {1: 0, 2: 1, 3: 2, 4: 0, 5: 0, 6: 1}[coord]
    c2 = {4: 1, 5: 2, 6: 2}.get(coord)
    if c2 is None:
        s0, s1 = x[2, c1] - m.shift_lo, x[2, c1] - m.shift_hi
    else:
        s0 = ((x[2, c1] - m.shift_lo) + (x[2, c2] - m.shift2_lo)) / 2
        s1 = ((x[2, c1] - m.shift_hi) + (x[2, c2] - m.shift2_hi)) / 2
    xi, dxi = s.calc_xi(x, 0.3, 1)
    assert abs(xi - s0 / (s0 - s1)) < 1e-14
    assert np.count_nonzero(dxi) == (1 if c2 is None else 2) and np.count_nonzero(dxi[:2]) == 0
    xi2, dxi2 = s.calc_xi(x, 0.3, 2)
    assert abs(xi2 - (0.3 * s1 + 0.7 * s0)) < 1e-14 and abs(dxi2.sum() - 1.0) < 1e-15
    _fd_check(s, x, 0.3)


def test_constrained_dynamics_hold_the_unimolecular_surface(oracle):
    rng = np.random.default_rng(2)
    s, m, ts = unimol_system(oracle, rng, nb=4)
    s.q[:] = ts[None] + rng.normal(0, 0.01, (4,) + ts.shape)
    s.set_rng(C.SEED, 3)
    s.set_thermostat(1, 10, 300.0)
    s.set_kforce(15.0)
    s.mdinit(1.0, 2)
    for i in range(1, 40):
        ep, xr, st = s.verlet(i, 1.0, constrain=1)
        assert st == 0
    assert abs(s.calc_xi(s.q.mean(axis=0), 1.0, 2)[0]) < 1e-8
