"""GPU parity of the flexible SPC water box (pes WATER_SPC, egrad_water.f90 via gradient.f90:212-213): energies and
gradients per image against the oracle within 1e-10 relative through crcl_egrad (gas-phase cluster, periodic box with
plain cut-off Coulomb and with Zahn's damped form, the 1000-molecule box of BASELINE config 5's shape), an RPMD
trajectory on the HBM-resident path with it as PES, and size-independent properties at full size."""
import numpy as np
import pytest

from caracal_b200 import water as WT
from tests import common as C

pytestmark = pytest.mark.gpu


def handle(gpu, W, nbeads=1, dt_fs=0.5, kelvin=300.0):
    mass = np.tile([C.atomic_mass_au("O"), C.atomic_mass_au("H"), C.atomic_mass_au("H")], W["n"] // 3)
    g = gpu.RPMD(gpu.PES_WATER, nbeads, mass, C.beta_calc_rate(kelvin), C.dt_au(dt_fs))
    g.set_water(W)
    return g, mass


@pytest.mark.parametrize("nwater,box,zahn,nimg", [(1, None, False, 5), (2, None, False, 40), (27, None, False, 16),
                                                  (64, 12.6, True, 16), (64, 12.6, False, 16), (216, 18.7, True, 4),
                                                  (1000, 31.07, True, 2), (43, 11.0, True, 33)])
def test_egrad_matches_oracle(gpu, oracle, nwater, box, zahn, nimg):
    rng = np.random.default_rng(nwater + nimg)
    W = WT.water_box(nwater, periodic_angstrom=None if box is None else [box, box, box], zahn=zahn)
    g, _ = handle(gpu, W)
    x0 = WT.water_lattice(nwater, box or 3.2 * np.ceil(nwater ** (1 / 3)), rng)
    x = x0[None] + rng.normal(0, 0.04, (nimg,) + x0.shape)
    Vo, go = oracle.Water(W).egrad(x)
    Vd, gd, _ = g.egrad(x)
    assert C.rel_err_E(Vd, Vo).max() < C.TOL_EG
    assert C.rel_err_G(gd.reshape(go.shape), go).max() < C.TOL_EG


def test_edge_cases(gpu, oracle):
    W = WT.water_box(8)
    g, _ = handle(gpu, W)
    V, gr, _ = g.egrad(np.zeros((0, 24, 3)))                       # no image
    assert V.shape == (0,)
    # atoms far outside the periodic box are imaged by the reference's while loop (box_image.f90)
    Wp = WT.water_box(8, periodic_angstrom=[9.0, 9.0, 9.0])
    gp, _ = handle(gpu, Wp)
    rng = np.random.default_rng(2)
    x = WT.water_lattice(8, 9.0, rng)
    x[3:6] += 3 * Wp["box"]
    x[9:12] -= np.array([2, 0, 5]) * Wp["box"]
    Vo, go = oracle.Water(Wp).egrad(x)
    Vd, gd, _ = gp.egrad(x)
    assert C.rel_err_E(Vd, Vo).max() < C.TOL_EG and C.rel_err_G(gd.reshape(go.shape), go).max() < C.TOL_EG
    with pytest.raises(gpu.CaracalGpuError):                       # molecules must be ordered O,H,H
        bad = dict(W, is_O=np.roll(W["is_O"], 1))
        handle(gpu, bad)


def test_rpmd_steps_match_oracle(gpu, oracle):
    """RPMD of a periodic 64-molecule box: 8 beads, 30 steps, Andersen, on the HBM-resident path"""
    nw, L, nb, nsteps = 64, 12.6, 8, 30
    W = WT.water_box(nw, periodic_angstrom=[L, L, L])
    g, m = handle(gpu, W, nbeads=nb)
    g.set_seed(C.SEED)
    g.set_thermostat(1, 7, 300.0)
    rng = np.random.default_rng(4)
    q0 = WT.water_lattice(nw, L, rng)[None, None] + rng.normal(0, 0.01, (2, nb, 3 * nw, 3))
    tid = np.array([5, 9], dtype=np.uint32)
    q = q0.copy()
    p, d, dxi, ev = g.mdinit(q, 0, traj_id=tid)
    ep, xr, st = g.verlet(q, p, d, nsteps=nsteps, constrain=-1, traj_id=tid, event=ev)
    Q = oracle.Water(W)
    for t in range(2):
        o = oracle.System(0, nb, m, C.beta_calc_rate(300.0), C.dt_au(0.5))
        o.set_custom_grad(lambda xyz: tuple(a[0] for a in Q.egrad(xyz)))
        o.q[:] = q0[t]
        o.set_rng(C.SEED, int(tid[t]))
        o.set_thermostat(1, 7, 300.0)
        o.mdinit(0.0, 0)
        for i in range(1, nsteps + 1):
            epo, _, sto = o.verlet(i, 0.0, -1)
            assert sto == 0
        assert st[t] == 0
        assert np.abs(q[t] - o.q).max() < C.TOL_QP
        assert (np.abs(p[t] - o.p) / np.abs(o.p).max()).max() < C.TOL_QP
        assert abs(ep[t] - epo) < 1e-9 * max(1.0, abs(epo))


def test_invariances_at_box_size(gpu):
    """1000 molecules x 8 images (one 8-bead step of config 5's shape): zero net force, invariance under a rigid shift
    and under lattice translations of whole molecules, and additivity over images."""
    rng = np.random.default_rng(6)
    nw, L = 1000, 31.07
    W = WT.water_box(nw, periodic_angstrom=[L, L, L])
    g, _ = handle(gpu, W)
    x0 = WT.water_lattice(nw, L, rng)
    x = x0[None] + rng.normal(0, 0.03, (8,) + x0.shape)
    V, gr, _ = g.egrad(x)
    gr = gr.reshape(x.shape)
    assert np.isfinite(V).all() and np.abs(gr.sum(axis=1)).max() < 1e-9
    y = x + 1.234
    mol = rng.integers(0, nw, 50)
    for m in mol:
        y[:, 3 * m:3 * m + 3] += W["box"] * rng.integers(-2, 3, 3)
    V2, g2, _ = g.egrad(y)
    assert np.abs(V2 - V).max() < 1e-9 * np.abs(V).max() and np.abs(g2.reshape(x.shape) - gr).max() < 1e-9
    V3, g3, _ = g.egrad(x[3:5])
    assert np.array_equal(V3, V[3:5]) or np.abs(V3 - V[3:5]).max() < 1e-12 * np.abs(V).max()
