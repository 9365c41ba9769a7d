"""GPU parity on the reference's SHIPPED QMDFF inputs: the periodic ethanol box of
examples/dynamic/ethanol_box (1125 atoms, 27 A, Zahn, H-bond search) and the two-state DG-EVB system of
examples/evbopt/DG-EVB (min1.qmdff / min2.qmdff with the key file's energy shifts, coord_def.inp; the
distributed-Gaussian coefficients are synthetic, because evb_pars.dat is an output of evbopt.x that the example
does not ship).  Tables from tests/qmdff_file.py; device against the oracle within 1e-10 relative through
crcl_egrad, and a short RPMD trajectory of the box on the split path."""
import numpy as np
import pytest

from tests import common as C
from tests import qmdff_file as QF
from tests.test_gpu_qmdff import torsion_conditioning

pytestmark = pytest.mark.gpu
SYM = {1: "H", 6: "C", 7: "N", 8: "O"}


def masses(T):
    return np.array([C.atomic_mass_au(SYM[int(z)]) for z in T["at"]])


def test_ethanol_box_matches_oracle(gpu, oracle):
    T = QF.tables("box", periodic_angstrom=[27.0, 27.0, 27.0])
    g = gpu.RPMD(gpu.PES_QMDFF, 1, masses(T), C.beta_calc_rate(200.0), C.dt_au(0.5))
    g.set_qmdff(T)
    Q = oracle.Qmdff(T)
    rng = np.random.default_rng(3)
    x0 = QF.box_start_bohr()
    x = x0[None] + rng.normal(0, 0.04, (3,) + x0.shape)
    x[0] = x0                                            # the shipped start structure itself
    Vo, go = Q.egrad(x)
    Vd, gd, _ = g.egrad(x)
    assert abs(Vo[0] / 125 + 5.5e-3) < 1e-3              # about -3.5 kcal/mol per ethanol at the start structure
    assert C.rel_err_E(Vd, Vo).max() < C.TOL_EG
    tol = np.maximum(C.TOL_EG, 2e-17 / torsion_conditioning(T, x) ** 2)
    err = C.rel_err_G(gd.reshape(go.shape), go)
    assert (err < tol).all(), (err / tol).max()
    assert (tol < 1e-9).all()


def test_ethanol_box_rpmd_steps_match_oracle(gpu, oracle):
    """sampling.key's shape in miniature: RPMD of the periodic box (4 beads, 5 steps, Andersen) on the split path"""
    T = QF.tables("box", periodic_angstrom=[27.0, 27.0, 27.0])
    nb, nsteps = 4, 5
    m = masses(T)
    beta, dt = C.beta_calc_rate(200.0), C.dt_au(0.5)
    g = gpu.RPMD(gpu.PES_QMDFF, nb, m, beta, dt)
    g.set_qmdff(T)
    g.set_seed(C.SEED)
    g.set_thermostat(1, 3, 200.0)
    rng = np.random.default_rng(5)
    q0 = QF.box_start_bohr()[None, None] + rng.normal(0, 0.01, (1, nb, T["n"], 3))
    q = q0.copy()
    p, d, dxi, ev = g.mdinit(q, 0)
    ep, xr, st = g.verlet(q, p, d, nsteps=nsteps, constrain=-1, event=ev)
    Q = oracle.Qmdff(T)
    o = oracle.System(0, nb, m, beta, dt)
    o.set_custom_grad(lambda xyz: tuple(a[0] for a in Q.egrad(xyz)))
    o.set_box(True, T["box"])                             # periodic tables: the wrap of verlet.f90:591-641 is on
    o.q[:] = q0[0]
    o.set_rng(C.SEED, 0)
    o.set_thermostat(1, 3, 200.0)
    o.mdinit(0.0, 0)
    for i in range(1, nsteps + 1):
        epo, _, sto = o.verlet(i, 0.0, -1)
        assert sto == 0
    assert st[0] == 0
    assert np.abs(q[0] - o.q).max() < C.TOL_QP
    assert (np.abs(p[0] - o.p) / np.abs(o.p).max()).max() < C.TOL_QP
    assert abs(ep[0] - epo) < 1e-9 * max(1.0, abs(epo))


@pytest.mark.parametrize("mode", [1, 3])
def test_dgevb_example_pair_matches_oracle(gpu, oracle, mode):
    T1, T2, coord_def, frames = QF.dgevb_example()
    nat6, npoints = len(coord_def), 7                     # evbopt.key: points 7; coord_def.inp: four bond lengths
    rng = np.random.default_rng(7)
    E = dict(mode=mode, coord_def=coord_def, g_thres=1e-10)
    tmp = oracle.Dgevb(T1, T2, dict(E, point_int=np.zeros((1, nat6)), alph=np.ones(1), b_vec=np.zeros(400)))
    pts = np.array([tmp.internals(frames[i]) for i in range(npoints)])       # Gaussians centred on path structures
    mat = {1: npoints, 3: npoints * (1 + nat6 + nat6 * (nat6 + 1) // 2)}[mode]
    E.update(point_int=pts, alph=rng.uniform(0.5, 2.5, npoints), b_vec=rng.normal(0, 2e-4, mat))
    g = gpu.RPMD(gpu.PES_DGEVB, 1, masses(T1), C.beta_calc_rate(300.0), C.dt_au(0.5))
    g.set_qmdff(T1)
    g.set_qmdff(T2, second=True)
    g.set_dgevb(E)
    D = oracle.Dgevb(T1, T2, E)
    x = np.concatenate([frames, frames + rng.normal(0, 0.03, frames.shape)])
    Vo, go = D.egrad(x)
    Vd, gd, _ = g.egrad(x)
    # the two diabatic minima differ by the key file's shifts plus the force-field energies: a few 1e-2 Eh
    assert np.isfinite(Vo).all() and -132.5 < Vo.min() < -132.0
    assert (np.abs(Vd - Vo) < C.TOL_EG * np.abs(Vo)).all()
    assert C.rel_err_G(gd.reshape(go.shape), go).max() < C.TOL_EG
