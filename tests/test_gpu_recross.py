"""GPU parity of the work-unit seam: recrossing child pairs (recross.f90:515-628 worker body)
and the umbrella-window worker body (calc_rate.f90:1387-1700), plus the RNG hook."""
import numpy as np
import pytest

from tests import common as C

pytestmark = pytest.mark.gpu


def parents(g, name, nb, nparent, xi_dag, steps=40, seed=5):
    """constrained parent snapshots (recross.f90:240-297), produced by the product itself"""
    rng = np.random.default_rng(seed)
    g.set_thermostat(1, 7, 300.0)
    q = np.array([C.SYSTEMS[name]["ts"]()[None] + rng.normal(0, 0.01, (nb,) + C.SYSTEMS[name]["ts"]().shape)
                  for _ in range(nparent)])
    p, d, dxi, ev = g.mdinit(q, 2, xi_dag, 0.0)
    g.verlet(q, p, d, nsteps=steps, constrain=1, xi_ideal=xi_dag, k_force=0.0, dxi=dxi, event=ev)
    return q


@pytest.mark.parametrize("name,nb,npairs,evol", [("h3", 8, 12, 120), ("ch4h", 16, 6, 80), ("oh3", 64, 2, 40),
                                                 ("h3", 1, 7, 50), ("brh2", 16, 5, 60), ("ch4oh", 8, 4, 60), ("geh4oh", 8, 3, 50), ("ch4cn", 8, 3, 50),
                                                 ("clnh3", 8, 3, 50), ("nh3oh", 8, 3, 50), ("h2co", 4, 2, 30)])
def test_kappa_sums_match_oracle(gpu, oracle, name, nb, npairs, evol):
    g, o = C.make_pair(name, nb)
    g.set_seed(C.SEED)
    xi_dag = 0.98
    qp = parents(g, name, nb, 3, xi_dag)
    num, den, st = g.recross_children(qp, npairs, evol, xi_dag, pair0=4)
    onum, oden, ost = o.recross_children(qp, 4, npairs, evol, xi_dag, C.SEED, nthreads=4)
    assert ost == 0 and (st == 0).all()
    assert abs(den - oden) < 1e-9 * abs(oden)
    # theta flips only if xi_real crosses zero within 1e-9 of a step boundary: compare exactly first,
    scale = np.abs(onum).max() + abs(oden)
    assert np.abs(num - onum).max() < 1e-9 * scale


def test_sharding_invariance_and_determinism(gpu):
    """BASELINE shape: 512 +/- pairs of CH4+H 16 beads; splitting the pair range over calls
    (the multi-GPU decomposition) changes nothing, and reruns are bit-identical."""
    name, nb = "ch4h", 16
    g, _ = C.make_pair(name, nb)
    g.set_seed(C.SEED)
    xi_dag = 0.98
    qp = parents(g, name, nb, 8, xi_dag)
    evol = 100
    num, den, st = g.recross_children(qp, 512, evol, xi_dag)
    num2, den2, _ = g.recross_children(qp, 512, evol, xi_dag)
    assert np.array_equal(num, num2) and den == den2          # deterministic reduction
    parts = [g.recross_children(qp, 128, evol, xi_dag, pair0=128 * r) for r in range(4)]
    snum = sum(p[0] for p in parts)
    sden = sum(p[1] for p in parts)
    assert np.abs(snum - num).max() < 1e-11 * np.abs(num).max()
    assert abs(sden - den) < 1e-11 * den
    # physics sanity of kappa(t): starts near 1 and stays within [0, 1.05]
    kappa = num / den
    assert 0.9 < kappa[0] <= 1.0 + 1e-12
    assert (kappa > -0.05).all() and (kappa < 1.05).all()
    assert (st == 0).all()


def test_empty_pair_range(gpu):
    g, _ = C.make_pair("h3", 8)
    qp = np.array([C.ring_polymer("h3", 8, np.random.default_rng(0))])
    num, den, st = g.recross_children(qp, 0, 10, 0.98)
    assert den == 0.0 and (num == 0).all()


def test_umbrella_window_matches_oracle(gpu, oracle):
    name, nb, ntraj, equi, samp = "h3", 8, 3, 30, 60
    g, _ = C.make_pair(name, nb)
    g.set_seed(C.SEED)
    g.set_thermostat(1, 11, 300.0)
    xi0, kf = 0.9, 0.05 * 300.0
    q0 = C.ring_polymer(name, nb, np.random.default_rng(8), 0.02)
    avg, var, st = g.umbrella_window(q0, xi0, kf, ntraj, equi, samp, traj_id0=50)
    for t in range(ntraj):
        _, o = C.make_pair(name, nb)
        o.q[:] = q0
        o.set_rng(C.SEED, 50 + t)
        o.set_thermostat(1, 11, 300.0)
        o.set_kforce(kf)
        o.mdinit(xi0, 2)
        for i in range(1, equi + 1):
            o.verlet(i, xi0, 0)
        o.gradient_all()          # calc_rate.f90:1619-1623: forces without the bias before sampling
        xs = []
        for i in range(1, samp + 1):
            xs.append(o.verlet(i, xi0, 0)[1])
        xs = np.array(xs)
        assert abs(avg[t] - xs.mean()) < 1e-9
        assert abs(var[t] - (np.mean(xs ** 2) - xs.mean() ** 2)) < 1e-9
    assert (st == 0).all()


def test_umbrella_windows_batch_equals_single_windows(gpu):
    name, nb, ntraj, equi, samp = "h3", 8, 3, 20, 40
    g, _ = C.make_pair(name, nb)
    g.set_seed(C.SEED)
    g.set_thermostat(1, 7, 300.0)
    rng = np.random.default_rng(9)
    q0 = np.array([C.ring_polymer(name, nb, rng, 0.02) for _ in range(4)])
    xi0, kf = np.array([0.2, 0.5, 0.9, 1.02]), np.array([15.0, 15.0, 12.0, 15.0])
    avg, var, st = g.umbrella_windows(q0, xi0, kf, ntraj, equi, samp, traj_id0=7)
    assert (st == 0).all()
    for w in range(4):
        a1, v1, s1 = g.umbrella_window(q0[w], xi0[w], kf[w], ntraj, equi, samp, traj_id0=7 + w * ntraj)
        assert np.array_equal(a1, avg[w]) and np.array_equal(v1, var[w])


def test_rng_stream_matches_oracle_and_is_normal(gpu, oracle):
    from scipy import stats
    g, _ = C.make_pair("h3", 8)
    for traj, event, bead in [(0, 0, 0), (77, 3, 5), (2 ** 31 + 3, 11, 63)]:
        z = g.rng_normals(C.SEED, traj, event, bead, 9)
        assert np.abs(z - oracle.normals(C.SEED, traj, event, bead, 9)).max() < 1e-13
    z = g.rng_normals(12345, 1, 0, 0, 400000)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert stats.kstest(z, "norm").pvalue > 1e-3


def test_calc_xi_entry_point(gpu, oracle):
    name = "ch4h"
    g, o = C.make_pair(name, 1)
    x = C.ts_cloud(name, 200, 0.1, np.random.default_rng(4))
    for mode in (1, 2):
        xi, dxi = g.calc_xi(x, 0.93, mode)
        for i in range(0, 200, 17):
            xo, dxo = o.calc_xi(x[i], 0.93, mode)
            assert abs(xi[i] - xo) < 1e-12 and np.abs(dxi[i] - dxo).max() < 1e-12
