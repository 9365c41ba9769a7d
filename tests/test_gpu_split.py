"""GPU parity of the split (HBM-resident) integrator path: any atom / bead count, device PES or the
host-callback PES seam (custom_grad.f90:35), against the oracle; plus size-independent properties
of the propagation kernel at the periodic-box shape (3000 atoms x 8 beads)."""
import numpy as np
import pytest

from tests import common as C

pytestmark = pytest.mark.gpu


def run_split(gpu, oracle, name, nb, constrain, thermo, astep, ntraj, nsteps=100, xi_ideal=0.98):
    rng = np.random.default_rng(nb * 1000 + constrain + 7)
    g, _ = C.make_pair(name, nb)
    g.set_path(gpu.PATH_SPLIT)
    g.set_seed(C.SEED)
    g.set_thermostat(thermo, astep, 300.0)
    q0 = np.array([C.ring_polymer(name, nb, rng, 0.03) for _ in range(ntraj)])
    tid = np.arange(300, 300 + ntraj, dtype=np.uint32)
    q = q0.copy()
    p, d, dxi, ev = g.mdinit(q, 0, traj_id=tid)
    ep, xr, st = g.verlet(q, p, d, nsteps=nsteps, constrain=constrain, xi_ideal=xi_ideal, traj_id=tid, event=ev)
    wq = wp = 0.0
    for t in range(ntraj):
        _, o = C.make_pair(name, nb)
        o.q[:] = q0[t]
        o.set_rng(C.SEED, int(tid[t]))
        o.set_thermostat(thermo, astep, 300.0)
        o.mdinit(0.0, 0)
        for i in range(1, nsteps + 1):
            epo, xro, sto = o.verlet(i, xi_ideal, constrain)
            assert sto == 0
        assert st[t] in (0, 16)
        wq = max(wq, np.abs(q[t] - o.q).max())
        wp = max(wp, (np.abs(p[t] - o.p) / np.abs(o.p).max()).max())
        assert abs(ep[t] - epo) < 1e-9 * max(1.0, abs(epo))
        if constrain == 2:
            assert abs(xr[t] - xro) < 1e-9
    return wq, wp


@pytest.mark.parametrize("name,nb,constrain,thermo,astep,ntraj", [
    ("h3", 16, -1, 1, 7, 3),      # config 1 through the split path (Andersen + transrot)
    ("ch4h", 16, 2, 0, 0, 3),     # child trajectory
    ("h3", 1, -1, 1, 10, 5),      # one bead
    ("h3", 6, 2, 0, 0, 2),        # non-power-of-two bead counts (register kernel)
    ("h3", 12, -1, 1, 9, 2),
    ("h3", 24, 2, 0, 0, 2),       # shared-memory kernel
    ("oh3", 64, 2, 0, 0, 1),
])
def test_split_path_matches_oracle(gpu, oracle, name, nb, constrain, thermo, astep, ntraj):
    wq, wp = run_split(gpu, oracle, name, nb, constrain, thermo, astep, ntraj)
    assert wq < C.TOL_QP and wp < C.TOL_QP, (wq, wp)


def test_fused_and_split_agree(gpu):
    name, nb, ntraj = "ch4h", 16, 8
    rng = np.random.default_rng(2)
    q0 = np.array([C.ring_polymer(name, nb, rng, 0.03) for _ in range(ntraj)])
    res = []
    for path in (gpu.PATH_FUSED, gpu.PATH_SPLIT):
        g, _ = C.make_pair(name, nb)
        g.set_path(path)
        g.set_seed(5)
        q = q0.copy()
        p, d, dxi, ev = g.mdinit(q, 0)
        g.verlet(q, p, d, nsteps=100, constrain=2, xi_ideal=0.98)
        res.append((q, p))
    assert np.abs(res[0][0] - res[1][0]).max() < C.TOL_QP
    assert np.abs(res[0][1] - res[1][1]).max() < C.TOL_QP * np.abs(res[0][1]).max()


def _toy_pes(natoms):
    """a user's custom_grad: anharmonic chain + pair terms (smooth, cheap, 10 atoms)"""
    k2, k4 = 0.35, 0.08

    def fn(x):
        d = x[1:] - x[:-1]
        r = np.linalg.norm(d, axis=1)
        e = np.sum(0.5 * k2 * (r - 1.8) ** 2 + k4 * (r - 1.8) ** 4) + 0.01 * np.sum(x ** 2)
        dr = (k2 * (r - 1.8) + 4 * k4 * (r - 1.8) ** 3)[:, None] * d / r[:, None]
        g = 0.02 * x
        g[1:] += dr
        g[:-1] -= dr
        return e, g
    return fn


def test_host_callback_pes_matches_oracle(gpu, oracle):
    """custom_grad seam: a user's host routine slots in unchanged; 10 atoms x 8 beads."""
    natoms, nb, ntraj, nsteps = 10, 8, 2, 60
    rng = np.random.default_rng(11)
    mass = np.array([C.atomic_mass_au(s) for s in ["C", "H", "O", "H", "C", "H", "H", "O", "H", "C"]])
    beta, dt = C.beta_calc_rate(300.0), C.dt_au(0.2)
    fn = _toy_pes(natoms)
    x0 = np.cumsum(rng.normal(0, 1.0, (natoms, 3)) + np.array([1.5, 0, 0]), axis=0)
    q0 = x0[None, None] + rng.normal(0, 0.02, (ntraj, nb, natoms, 3))
    g = gpu.RPMD(gpu.PES_HOSTCB, nb, mass, beta, dt)
    g.set_host_gradient(fn)
    g.set_seed(C.SEED)
    g.set_thermostat(1, 13, 300.0)
    tid = np.array([7, 8], dtype=np.uint32)
    q = q0.copy()
    p, d, dxi, ev = g.mdinit(q, 0, traj_id=tid)
    ep, xr, st = g.verlet(q, p, d, nsteps=nsteps, constrain=-1, traj_id=tid, event=ev)
    for t in range(ntraj):
        o = oracle.System(0, nb, mass, beta, dt)
        o.set_custom_grad(fn)
        o.q[:] = q0[t]
        o.set_rng(C.SEED, int(tid[t]))
        o.set_thermostat(1, 13, 300.0)
        o.mdinit(0.0, 0)
        for i in range(1, nsteps + 1):
            epo, _, sto = o.verlet(i, 0.0, -1)
            assert sto == 0
        assert np.abs(q[t] - o.q).max() < C.TOL_QP
        assert (np.abs(p[t] - o.p) / np.abs(o.p).max()).max() < C.TOL_QP
        assert abs(ep[t] - epo) < 1e-9 * max(1.0, abs(epo))


def test_free_ring_polymer_at_box_shape(gpu):
    """3000 atoms x 8 beads (config 5 shape) with a zero PES: the EXACT transform is the exact
    propagator of the free ring polymer (its energy is conserved for any dt, the centroid moves
    uniformly); REFERENCE mode symmetrises the beads (F2)."""
    natoms, nb, nsteps = 3000, 8, 20
    rng = np.random.default_rng(3)
    mass = np.tile([C.atomic_mass_au("O"), C.atomic_mass_au("H"), C.atomic_mass_au("H")], natoms // 3)
    beta, dt = C.beta_calc_rate(300.0), C.dt_au(0.5)
    g = gpu.RPMD(gpu.PES_HOSTCB, nb, mass, beta, dt)
    g.set_host_gradient(lambda x: (0.0, np.zeros_like(x)))
    g.set_thermostat(0, 0, 300.0)
    ts = rng.normal(0, 5.0, (natoms, 3))
    g.set_mechanism(C.Mechanism([[1, 2]], [[3, 4]], [[1, 3], [2, 4]], 16.0, ts))   # only so that constrain=2 runs
    q0 = ts[None, None] + rng.normal(0, 0.05, (1, nb, natoms, 3))
    m = mass[None, None, :, None]
    beta_n = beta / nb

    def frp_energy(q, p):
        return (p ** 2 / (2 * m)).sum() + 0.5 * (m * (q - np.roll(q, 1, axis=1)) ** 2).sum() / beta_n ** 2
    g.set_seed(9)
    g.set_transform(gpu.TRANSFORM_EXACT)
    q = q0.copy()
    p, d, dxi, ev = g.mdinit(q, 0)
    e0, c0, pc = frp_energy(q, p), q.mean(axis=1), p.mean(axis=1)
    ep, xr, st = g.verlet(q, p, d, nsteps=nsteps, constrain=2, xi_ideal=0.5)
    assert st[0] == 0
    assert abs(frp_energy(q, p) - e0) < 1e-10 * e0
    assert np.abs(q.mean(axis=1) - (c0 + nsteps * dt * pc / mass[None, :, None])).max() < 1e-9
    assert np.abs(p.mean(axis=1) - pc).max() < 1e-12 * np.abs(pc).max()
    g.set_transform(gpu.TRANSFORM_REFERENCE)
    q = q0.copy()
    p, d, dxi, ev = g.mdinit(q, 0)
    g.verlet(q, p, d, nsteps=1, constrain=2, xi_ideal=0.5)
    for a in range(1, nb):
        assert np.abs(q[0, a] - q[0, nb - a]).max() < 1e-11
    ms, best, gbs = g.bench_propagate(64, reps=3)
    assert ms > 0 and gbs > 100.0
