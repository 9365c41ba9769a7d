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


# ---- umbrella bias, SHAKE / RATTLE, Nose-Hoover chain, work units on the split path ------------------
def run_split_biased(gpu, oracle, name, nb, constrain, thermo, astep, bias_mode, ntraj=2, nsteps=100, xi0=0.97,
                     kf=0.05 * 300.0, nose_q=100.0):
    rng = np.random.default_rng(nb * 100 + constrain + 3 * thermo)
    g, _ = C.make_pair(name, nb)
    g.set_path(gpu.PATH_SPLIT)
    g.set_seed(C.SEED)
    g.set_thermostat(thermo, astep, 300.0, nose_q)
    q0 = np.array([C.ring_polymer(name, nb, rng, 0.02) for _ in range(ntraj)])
    tid = np.arange(40, 40 + ntraj, dtype=np.uint32)
    xi = np.full(ntraj, xi0) + 0.01 * np.arange(ntraj)
    k = np.full(ntraj, kf)
    q = q0.copy()
    p, d, dxi, ev = g.mdinit(q, bias_mode, xi_ideal=xi, k_force=k, traj_id=tid)
    ep, xr, st = g.verlet(q, p, d, nsteps=nsteps, constrain=constrain, xi_ideal=xi, k_force=k, dxi=dxi, traj_id=tid,
                          event=ev)
    for t in range(ntraj):
        _, o = C.make_pair(name, nb)
        o.q[:] = q0[t]
        o.set_rng(C.SEED, int(tid[t]))
        o.set_thermostat(thermo, astep, 300.0, nose_q)
        o.set_kforce(kf)
        o.mdinit(float(xi[t]), bias_mode)
        for i in range(1, nsteps + 1):
            epo, xro, sto = o.verlet(i, float(xi[t]), constrain)
            assert sto == 0
        assert st[t] == 0
        assert np.abs(q[t] - o.q).max() < C.TOL_QP
        assert (np.abs(p[t] - o.p) / np.abs(o.p).max()).max() < C.TOL_QP
        assert abs(ep[t] - epo) < 1e-9 * max(1.0, abs(epo))
        assert abs(xr[t] - xro) < 1e-9
        assert np.abs(dxi[t] - o.dxi).max() < 1e-9
        if constrain == 1:
            # xi_real is evaluated on the centroid of step 6, before SHAKE moved the beads (verlet.f90:649,
            # 1047); on the constrained structure itself xi vanishes
            assert abs(g.calc_xi(q[t].mean(axis=0), float(xi[t]), 2)[0][0]) < 1e-7


@pytest.mark.parametrize("name,nb,constrain,thermo,astep,bias_mode", [
    ("h3", 8, 0, 1, 9, 2),        # umbrella window dynamics (bias + hams force + transrot)
    ("ch4h", 16, 0, 1, 11, 2),
    ("h3", 8, 3, 1, 9, 2),        # same without the rotation removal
    ("h3", 8, 1, 1, 9, 2),        # constrained parent: SHAKE / RATTLE
    ("ch4h", 4, 1, 1, 7, 2),
    ("oh3", 6, 1, 0, 0, 1),       # non-power-of-two beads, mdinit bias_mode 1
    ("h3", 8, -1, 2, 0, 0),       # Nose-Hoover chain
    ("h3", 12, 0, 2, 0, 2),       # NHC + umbrella
])
def test_split_biased_modes_match_oracle(gpu, oracle, name, nb, constrain, thermo, astep, bias_mode):
    run_split_biased(gpu, oracle, name, nb, constrain, thermo, astep, bias_mode)


def test_split_work_units_equal_fused(gpu):
    """crcl_recross_children and crcl_umbrella_windows through the split kernels give what the fused
    trajectory kernels give (same RNG streams, same operation sequence)."""
    name, nb = "ch4h", 8
    rng = np.random.default_rng(4)
    qp = np.array([C.ring_polymer(name, nb, rng, 0.01) for _ in range(3)])
    q0 = np.array([C.ring_polymer(name, nb, rng, 0.02) for _ in range(3)])
    xi0, kf = np.array([0.9, 0.97, 1.01]), np.full(3, 15.0)
    res = []
    for path in (gpu.PATH_FUSED, gpu.PATH_SPLIT):
        g, _ = C.make_pair(name, nb)
        g.set_path(path)
        g.set_seed(C.SEED)
        num, den, st = g.recross_children(qp, 6, 60, 0.985, pair0=5)
        g.set_thermostat(1, 7, 300.0)
        avg, var, st2 = g.umbrella_windows(q0, xi0, kf, 2, 30, 50, traj_id0=9)
        assert (st == 0).all() and (st2 == 0).all()
        res.append((num, den, avg, var))
    (n1, d1, a1, v1), (n2, d2, a2, v2) = res
    assert abs(d1 - d2) < 1e-10 * abs(d1) and np.abs(n1 - n2).max() < 1e-10 * abs(d1)
    assert np.abs(a1 - a2).max() < 1e-9 and np.abs(v1 - v2).max() < 1e-10


def test_rate_pipeline_on_a_qmdff_surface_runs_on_the_split_path(gpu):
    """configuration 4 shape in miniature: a DG-EVB surface (two QMDFFs + coupling, 9 atoms) through
    mdinit / umbrella windows / constrained parent / recrossing children on the split path."""
    from tests.qmdff_synth import make_dgevb
    T1, T2, E = make_dgevb(seed=5, mode=3, npoints=4)
    nb = 4
    mass = np.array([C.atomic_mass_au({1: "H", 6: "C", 8: "O"}[int(z)]) for z in T1["at"]])
    g = gpu.RPMD(gpu.PES_DGEVB, nb, mass, C.beta_calc_rate(300.0), C.dt_au(0.2))
    g.set_qmdff(T1)
    g.set_qmdff(T2, second=True)
    g.set_dgevb(E)
    g.set_seed(C.SEED)
    ts = T1["xyz"]
    # "reaction": O-H bond (3-9) breaks, H moves to C1 (1-9); fragments: the rest / the hydrogen
    g.set_mechanism(C.Mechanism([[1, 9]], [[3, 9]], [[1, 2, 3, 4, 5, 6, 7, 8], [9]], 12.0, ts))
    g.set_thermostat(1, 5, 300.0)
    q0 = np.repeat(ts[None, None], nb, axis=1) + np.random.default_rng(1).normal(0, 0.01, (1, nb) + ts.shape)
    # the reference structure is the "TS" of this mechanism: s1 = 0 there, xi = 1 in the umbrella form
    avg, var, st = g.umbrella_windows(np.repeat(q0, 2, axis=0), np.array([1.0, 0.99]), np.array([15.0, 15.0]), 2, 20, 40)
    assert (st == 0).all() and np.isfinite(avg).all() and (var >= 0).all() and np.abs(avg - 1.0).max() < 0.1
    q = q0.copy()
    p, d, dxi, ev = g.mdinit(q, 2, xi_ideal=1.0, k_force=15.0)
    ep, xr, st = g.verlet(q, p, d, nsteps=30, constrain=1, xi_ideal=1.0, k_force=15.0, dxi=dxi, event=ev)
    assert st[0] == 0 and abs(g.calc_xi(q[0].mean(axis=0), 1.0, 2)[0][0]) < 1e-7   # generic-size calc_xi
    g.set_thermostat(0, 0, 300.0)
    num, den, stc = g.recross_children(q, 4, 25, 1.0)
    assert (stc == 0).all() and den > 0 and np.isfinite(num).all() and abs(num[0] / den - 1.0) < 0.5


def test_graph_replay_is_bit_identical_to_eager_launches(gpu):
    """Steps 2..n of a split-path call are replayed from a CUDA graph (crcl_set_graph); the replay launches the
    same kernels with the same arguments, so every mode must give bit-identical state with it on and off:
    Andersen on some steps only (two graph variants), SHAKE / RATTLE, NHC, the umbrella accumulators and the
    per-step theta rows of the recrossing work unit."""
    name, nb, ntraj = "ch4h", 8, 3
    rng = np.random.default_rng(12)
    q0 = np.array([C.ring_polymer(name, nb, rng, 0.02) for _ in range(ntraj)])
    qp = np.array([C.ring_polymer(name, nb, rng, 0.01) for _ in range(2)])
    xi, k = np.full(ntraj, 0.97), np.full(ntraj, 15.0)
    out = []
    for on in (1, 0):
        g, _ = C.make_pair(name, nb)
        g.set_path(gpu.PATH_SPLIT)
        g.set_graph(on)
        g.set_seed(C.SEED)
        res = []
        for constrain, thermo, astep, bias in ((-1, 1, 7, 0), (0, 1, 5, 2), (1, 1, 9, 2), (2, 0, 0, 0), (0, 2, 0, 2)):
            g.set_thermostat(thermo, astep, 300.0, 100.0)
            q = q0.copy()
            p, d, dxi, ev = g.mdinit(q, bias, xi_ideal=xi, k_force=k)
            l0 = g.launch_count()
            ep, xr, st = g.verlet(q, p, d, nsteps=40, constrain=constrain, xi_ideal=xi, k_force=k, dxi=dxi, event=ev)
            res += [q, p, d, ep, xr, np.array([g.launch_count() - l0])]
        g.set_thermostat(0, 0, 300.0)
        num, den, st = g.recross_children(qp, 3, 30, 0.985, pair0=2)
        g.set_thermostat(1, 7, 300.0)
        avg, var, st2 = g.umbrella_windows(q0, np.array([0.9, 0.97, 1.01]), k, 2, 20, 30, traj_id0=4)
        out.append(res + [num, np.array([den]), avg, var])
    for a, b in zip(*out):
        assert np.array_equal(a, b)
