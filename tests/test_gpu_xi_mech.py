"""GPU parity of the unimolecular and ATOM_SHIFT reaction coordinates (calc_xi.f90:523-938, SURVEY.md
8(f) N3) against the oracle: xi, its gradient, the umbrella 'hams' force, and whole biased / constrained
/ child trajectories, on the fused in-register path and on the split path."""
import numpy as np
import pytest

from caracal_b200.api import AtomShiftMechanism
from tests import common as C
from tests.test_oracle_xi_mech import shift_system, unimol_system

pytestmark = pytest.mark.gpu


def make(gpu, oracle, kind, nb, path):
    rng = np.random.default_rng(7)
    if kind == "unimol":
        o, m, ts = unimol_system(oracle, rng, nb)
        name = "ch4h"
    else:
        o, m = shift_system(oracle, int(kind[-1]), nb)
        name, ts = "h3", C.h3_ts()
    g = gpu.RPMD(name, nb, C.masses(name), C.beta_calc_rate(300.0), C.dt_au(0.1))
    g.set_mechanism(m)
    g.set_path(path)
    g.set_seed(C.SEED)
    return g, o, m, ts, name


@pytest.mark.parametrize("kind", ["unimol", "shift1", "shift3", "shift5"])
def test_calc_xi_and_hams_match_oracle(gpu, oracle, kind):
    g, o, m, ts, name = make(gpu, oracle, kind, 1, gpu.PATH_AUTO)
    rng = np.random.default_rng(2)
    x = ts[None] + rng.normal(0, 0.1, (40,) + ts.shape)
    for mode in (1, 2):
        xi, dxi = g.calc_xi(x, 0.8, mode)
        for i in range(0, 40, 7):
            xo, dxo = o.calc_xi(x[i], 0.8, mode)
            assert abs(xi[i] - xo) < 1e-12 and np.abs(dxi[i] - dxo).max() < 1e-12
    # hams force of umbrella.f90:144-174 (uses the Hessian of xi): oracle's umbrella on a zero gradient
    xi, dxi, hams = g.calc_xi(x, 0.8, 1, hams=True)
    o.set_kforce(0.0)
    for i in range(0, 40, 7):
        grad = np.zeros((1,) + ts.shape)
        o.umbrella(x[i], 0.8, grad, 0)
        assert np.abs(hams[i] - grad[0]).max() < 1e-10 * max(1.0, np.abs(grad).max())


@pytest.mark.parametrize("kind,constrain,path", [("unimol", 0, "fused"), ("unimol", 1, "fused"), ("unimol", 2, "fused"),
                                                 ("unimol", 0, "split"), ("unimol", 1, "split"), ("shift3", 0, "fused"),
                                                 ("shift3", 1, "split"), ("shift5", 0, "split")])
def test_trajectories_match_oracle(gpu, oracle, kind, constrain, path):
    nb, nsteps = 8, 60
    g, o, m, ts, name = make(gpu, oracle, kind, nb, gpu.PATH_FUSED if path == "fused" else gpu.PATH_SPLIT)
    rng = np.random.default_rng(5)
    q0 = ts[None, None] + rng.normal(0, 0.01, (1, nb) + ts.shape)
    # a window / dividing surface close to where the structure is
    xi0 = float(o.calc_xi(q0[0].mean(axis=0), 1.0, 1)[0]) if constrain == 0 else 1.0
    if constrain == 1 and kind != "unimol":
        xi0 = 0.5
    thermo = (1, 9) if constrain != 2 else (0, 0)
    g.set_thermostat(thermo[0], thermo[1], 300.0)
    o.set_thermostat(thermo[0], thermo[1], 300.0)
    o.set_kforce(15.0)
    tid = np.array([21], dtype=np.uint32)
    q = q0.copy()
    bias_mode = 2 if constrain in (0, 1) else 1
    p, d, dxi, ev = g.mdinit(q, bias_mode, xi_ideal=xi0, k_force=15.0, traj_id=tid)
    ep, xr, st = g.verlet(q, p, d, nsteps=nsteps, constrain=constrain, xi_ideal=xi0, k_force=15.0, dxi=dxi, traj_id=tid,
                          event=ev)
    o.q[:] = q0[0]
    o.set_rng(C.SEED, 21)
    o.mdinit(xi0, bias_mode)
    for i in range(1, nsteps + 1):
        epo, xro, sto = o.verlet(i, xi0, constrain)
        assert sto == 0
    assert st[0] == 0
    assert np.abs(q[0] - o.q).max() < C.TOL_QP
    assert (np.abs(p[0] - o.p) / np.abs(o.p).max()).max() < C.TOL_QP
    assert abs(xr[0] - xro) < 1e-9 and abs(ep[0] - epo) < 1e-9 * max(1.0, abs(epo))


def test_mechanism_argument_checks(gpu):
    g = gpu.RPMD("h3", 4, C.masses("h3"), C.beta_calc_rate(300.0), C.dt_au(0.1))
    with pytest.raises(gpu.CaracalGpuError):
        g.set_mechanism(AtomShiftMechanism(4, 1, 0.0, 1.0))         # atom out of range
    with pytest.raises(gpu.CaracalGpuError):
        g.set_mechanism(AtomShiftMechanism(1, 7, 0.0, 1.0))         # coordinate code out of range
