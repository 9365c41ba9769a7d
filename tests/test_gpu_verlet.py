"""GPU parity of the integrator seam (crcl_mdinit / crcl_verlet) against the oracle: positions
and momenta after 100 steps from identical initial state within 1e-8 (BASELINE.json north_star),
for every constrain mode and thermostat of the graded paths, several bead counts (1 bead,
sub-warp groups, full warp, multi-warp) and ragged batch sizes."""
import json
import os

import numpy as np
import pytest

from tests import common as C

pytestmark = pytest.mark.gpu

CASES = [
    # name, nbeads, constrain, thermostat, andersen_step, bias_mode, xi_ideal, k_force, ntraj
    ("h3", 16, -1, 1, 70, 0, 0.0, 0.0, 5),        # config 1: dynamic.x NVT Andersen
    ("h3", 16, -1, 1, 7, 0, 0.0, 0.0, 3),         # several Andersen redraws inside 100 steps
    ("h3", 8, 0, 1, 80, 2, 0.9, 15.0, 9),         # umbrella window (rate.key: 8 beads)
    ("h3", 8, 1, 1, 31, 2, 0.98, 15.0, 4),        # constrained parent (SHAKE/RATTLE)
    ("h3", 8, 2, 0, 0, 2, 0.98, 0.0, 6),          # child trajectory
    ("h3", 1, 0, 1, 50, 2, 0.95, 15.0, 33),       # nbeads=1 (start-structure generation phase)
    ("h3", 4, -1, 2, 0, 0, 0.0, 0.0, 3),          # Nose-Hoover chain
    ("h3", 32, 2, 0, 0, 2, 0.98, 0.0, 3),         # full-warp group
    ("h3", 64, 2, 0, 0, 2, 0.98, 0.0, 2),         # CTA-wide group
    ("oh3", 64, 2, 0, 0, 2, 0.97, 0.0, 2),        # config 3 shape
    ("oh3", 4, 0, 2, 0, 2, 0.8, 15.0, 5),
    ("ch4h", 16, 2, 0, 0, 2, 0.97, 0.0, 5),       # config 2: child trajectories
    ("ch4h", 16, 1, 1, 31, 2, 0.97, 15.0, 3),     # config 2: parent
    ("ch4h", 16, 0, 1, 80, 2, 0.5, 15.0, 3),      # config 2: umbrella window
    ("ch4h", 2, -1, 0, 0, 0, 0.0, 0.0, 17),
    ("ch4h", 32, 2, 0, 0, 2, 0.97, 0.0, 2),       # four warps per trajectory: tensor-core transform with 4 row tiles
    ("ch4h", 8, 0, 1, 40, 2, 0.9, 15.0, 3),       # one warp per trajectory
    ("brh2", 16, 0, 1, 40, 2, 0.9, 15.0, 3),      # DIM-3C Br + H2, umbrella window
    ("brh2", 8, 1, 1, 31, 2, 0.98, 15.0, 2),      # constrained parent
    ("brh2", 32, 2, 0, 0, 2, 0.98, 0.0, 3),       # child trajectories
    ("o3", 16, 0, 1, 40, 2, 0.9, 15.0, 3),        # O3 1 1A" PIP surface, umbrella window
    ("o3", 8, 2, 0, 0, 2, 0.98, 0.0, 4),          # child trajectories
    ("ch4oh", 16, 0, 1, 40, 2, 0.9, 15.0, 3),     # CH4 + OH (7 atoms), umbrella window
    ("ch4oh", 8, 1, 1, 31, 2, 0.98, 15.0, 2),     # constrained parent
    ("ch4oh", 16, 2, 0, 0, 2, 0.98, 0.0, 3),      # child trajectories
    ("geh4oh", 8, 0, 1, 40, 2, 0.9, 15.0, 3),     # GeH4 + OH, umbrella window
    ("geh4oh", 16, 2, 0, 0, 2, 0.98, 0.0, 2),     # child trajectories
    ("ch4cn", 8, 0, 1, 40, 2, 0.9, 15.0, 3),      # CH4 + CN, umbrella window
    ("ch4cn", 16, 2, 0, 0, 2, 0.98, 0.0, 2),      # child trajectories
    ("clnh3", 8, 0, 1, 40, 2, 0.9, 15.0, 3),      # NH3 + Cl (5 atoms), umbrella window
    ("clnh3", 16, 1, 1, 31, 2, 0.98, 15.0, 2),    # constrained parent
    ("clnh3", 32, 2, 0, 0, 2, 0.98, 0.0, 2),      # child trajectories
    ("nh3oh", 8, 0, 1, 40, 2, 0.9, 15.0, 3),      # NH3 + OH (numeric gradient, four lanes per bead)
    ("nh3oh", 16, 2, 0, 0, 2, 0.98, 0.0, 2),      # child trajectories
    ("nh3oh", 1, 0, 1, 50, 2, 0.95, 15.0, 5),     # one bead
    ("ch4h", 1, 0, 1, 50, 2, 0.95, 15.0, 5),      # one bead, eight quads per trajectory (PesSpreadQ)
    ("ch4h", 1, 1, 1, 31, 2, 0.98, 15.0, 3),      # the same under SHAKE / RATTLE
    ("ch4oh", 1, 0, 1, 50, 2, 0.9, 15.0, 3),      # 7 atoms: six owned slots per surface lane
    ("oh3", 1, 0, 1, 50, 2, 0.9, 15.0, 5),        # one bead, 12 components on 16 lanes (PesSpread)
    ("h3", 2, 0, 1, 50, 2, 0.9, 15.0, 3),         # two beads x eight lanes
    ("oh3", 4, 1, 1, 31, 2, 0.95, 15.0, 3),       # four beads x four lanes, SHAKE / RATTLE
    ("brh2", 4, 0, 1, 40, 2, 0.9, 15.0, 3),       # four beads x four lanes
    ("h2co", 8, 0, 1, 40, 2, 0.9, 15.0, 2),       # H2CO fit (central-difference gradient, four lanes per bead)
    ("h2co", 4, 2, 0, 0, 2, 0.98, 0.0, 2),        # child trajectories
]


def run_case(gpu, oracle, name, nb, constrain, thermo, astep, bias_mode, xi_ideal, kf, ntraj, nsteps=100,
             chunks=(100,), spread_max=None):
    rng = np.random.default_rng(abs(hash((name, nb, constrain, thermo))) % 2 ** 31)
    g, _ = C.make_pair(name, nb)
    if spread_max is not None:
        g.set_spread_max_beads(spread_max)
    g.set_seed(C.SEED)
    g.set_thermostat(thermo, astep, 300.0, 100.0)
    q0 = np.array([C.ring_polymer(name, nb, rng, 0.03) for _ in range(ntraj)])
    tid = np.arange(100, 100 + ntraj, dtype=np.uint32)
    # product
    q = q0.copy()
    p, d, dxi, ev = g.mdinit(q, bias_mode, xi_ideal, kf, traj_id=tid)
    st = np.zeros(ntraj, dtype=np.int32)
    done = 0
    for ch in chunks:
        ep, xr, st = g.verlet(q, p, d, nsteps=ch, istep0=done, constrain=constrain, xi_ideal=xi_ideal, k_force=kf,
                              dxi=dxi, status=st, traj_id=tid, event=ev)
        done += ch
    assert done == nsteps
    # oracle, one trajectory at a time
    worst_q = worst_p = 0.0
    for t in range(ntraj):
        _, o = C.make_pair(name, nb)
        o.q[:] = q0[t]
        o.set_rng(C.SEED, int(tid[t]))
        o.set_thermostat(thermo, astep, 300.0, 100.0)
        o.set_kforce(kf)
        o.mdinit(xi_ideal, bias_mode)
        for i in range(1, nsteps + 1):
            epo, xro, sto = o.verlet(i, xi_ideal, constrain)
            assert sto == 0
        assert st[t] in (0, 16)
        worst_q = max(worst_q, np.abs(q[t] - o.q).max())
        worst_p = max(worst_p, (np.abs(p[t] - o.p) / np.abs(o.p).max()).max())
        assert abs(ep[t] - epo) < 1e-9 * max(1.0, abs(epo))
        if constrain >= 0:
            assert abs(xr[t] - xro) < 1e-9
    return worst_q, worst_p


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-nb%d-c%d-th%d" % (c[0], c[1], c[2], c[3]))
def test_100_steps_match_oracle(gpu, oracle, case):
    wq, wp = run_case(gpu, oracle, *case)
    assert wq < C.TOL_QP, "positions differ by %g" % wq
    assert wp < C.TOL_QP, "momenta differ by %g (relative to max |p|)" % wp


@pytest.mark.parametrize("name,nb,constrain", [("h3", 1, 0), ("ch4h", 1, 0), ("brh2", 1, 0), ("h3", 8, 0), ("h3", 8, 1),
                                               ("oh3", 4, 0), ("o3", 2, 0)])
def test_packed_form_of_few_bead_trajectories_matches_oracle(gpu, oracle, name, nb, constrain):
    """Small batches of trajectories with fewer threads than components run spread over 16 / 32 lanes by default
    (crcl_set_spread_max_beads); the packed form that large batches keep is held to the same bar."""
    wq, wp = run_case(gpu, oracle, name, nb, constrain, 1, 31 if constrain == 1 else 50, 2, 0.95, 15.0, 4, spread_max=0)
    assert wq < C.TOL_QP and wp < C.TOL_QP, (wq, wp)


def test_chunked_calls_equal_single_call(gpu, oracle):
    """the drivers call verlet one step at a time: 100 x 1 step == 1 x 100 steps (incl. Andersen phase)"""
    wq, wp = run_case(gpu, oracle, "h3", 8, 0, 1, 7, 2, 0.9, 15.0, 3, chunks=(1,) * 20 + (30, 50))
    assert wq < C.TOL_QP and wp < C.TOL_QP


def test_golden_trajectories(gpu):
    with open(os.path.join(os.path.dirname(__file__), "golden", "traj_golden.json")) as f:
        T = json.load(f)
    for rec in T:
        g, _ = C.make_pair(rec["name"], rec["nbeads"])
        g.set_seed(C.SEED)
        g.set_thermostat(rec["thermostat"], rec["andersen_step"], 300.0, rec.get("nose_q", 0.0))
        q = np.array(rec["q0"])[None].copy()
        tid = np.array([rec["traj"]], dtype=np.uint32)
        p, d, dxi, ev = g.mdinit(q, rec["bias_mode"], rec["xi_ideal"], rec["k_force"], traj_id=tid)
        g.verlet(q, p, d, nsteps=rec["nsteps"], constrain=rec["constrain"], xi_ideal=rec["xi_ideal"],
                 k_force=rec["k_force"], dxi=dxi, traj_id=tid, event=ev)
        assert np.abs(q[0] - np.array(rec["q"])).max() < C.TOL_QP
        pr = np.array(rec["p"])
        assert (np.abs(p[0] - pr) / np.abs(pr).max()).max() < C.TOL_QP


def test_bead_symmetry_and_exact_transform(gpu):
    """F2 at scale: in REFERENCE mode beads a and N-a coincide after step 1; in EXACT mode they do
    not, and the ring-polymer Hamiltonian is conserved with the O(dt^2) error of velocity Verlet
    (halving dt over the same time span quarters the energy error)."""
    name, nb, ntraj = "h3", 16, 2048
    rng = np.random.default_rng(3)
    q0 = C.h3_ts()[None, None] + rng.normal(0, 0.02, (ntraj, nb, 3, 3))
    m = C.masses(name)[None, None, :, None]
    beta = C.beta_calc_rate(300.0)
    beta_n = beta / nb

    def ham(q, p):
        V = gpu.egrad(name, q.reshape(-1, 3, 3))[0].reshape(ntraj, nb).sum(axis=1)
        spring = 0.5 * (m * (q - np.roll(q, 1, axis=1)) ** 2).sum(axis=(1, 2, 3)) / beta_n ** 2
        return (p ** 2 / (2 * m)).sum(axis=(1, 2, 3)) + spring + V
    g, _ = C.make_pair(name, nb)
    g.set_seed(1)
    q = q0.copy()
    p0, d0, dxi, ev = g.mdinit(q, 2, 0.98, 0.0)
    p, d = p0.copy(), d0.copy()
    g.verlet(q, p, d, nsteps=50, constrain=2, xi_ideal=0.98, dxi=dxi)
    assert max(np.abs(q[:, a] - q[:, nb - a]).max() for a in range(1, nb)) < 1e-12
    errs = []
    for fac in (1, 2):
        g.set_transform(gpu.TRANSFORM_EXACT)
        g.set_beta_dt(beta, C.dt_au(0.1) / fac)
        q, p, d = q0.copy(), p0.copy(), d0.copy()
        e0 = ham(q, p)
        g.verlet(q, p, d, nsteps=100 * fac, constrain=2, xi_ideal=0.98, dxi=dxi.copy())
        assert max(np.abs(q[:, a] - q[:, nb - a]).max() for a in range(1, nb)) > 1e-3
        errs.append(np.abs(ham(q, p) - e0))
    assert errs[0].max() < 2e-3
    ratio = np.median(errs[0] / np.maximum(errs[1], 1e-14))
    assert 2.5 < ratio < 6.0, ratio


def test_nan_status_is_reported_not_fatal(gpu):
    g, _ = C.make_pair("h3", 4)
    q = np.array([C.ring_polymer("h3", 4, np.random.default_rng(1)) for _ in range(3)])
    p, d, dxi, ev = g.mdinit(q, 2, 0.98, 0.0)
    p[1, 0, 0, 0] = np.nan
    ep, xr, st = g.verlet(q, p, d, nsteps=5, constrain=2, xi_ideal=0.98, dxi=dxi)
    assert st[1] & 2 and st[0] == 0 and st[2] == 0
    assert np.isfinite(q[0]).all() and np.isfinite(q[2]).all()


def test_unsupported_bead_count_is_an_error(gpu):
    """12 beads do not fit the fused kernels: an explicit error when that path is forced, the split
    path when the choice is left to the library"""
    g = gpu.RPMD("h3", 12, C.masses("h3"), C.beta_calc_rate(300.0), C.dt_au(0.1))
    q = np.array([C.ring_polymer("h3", 12, np.random.default_rng(0))])
    g.set_path(gpu.PATH_FUSED)
    with pytest.raises(gpu.CaracalGpuError, match="ENOSUP"):
        g.mdinit(q)
    g.set_path(gpu.PATH_AUTO)
    p, d, dxi, ev = g.mdinit(q)
    assert np.isfinite(p).all() and np.abs(p).max() > 0
    # the split path covers the constrained modes too, but needs the MECHA tables like the fused one
    with pytest.raises(gpu.CaracalGpuError, match="ESTATE"):
        g.verlet(q, p, d, nsteps=1, constrain=1, xi_ideal=0.9, k_force=1.0)
    g.set_mechanism(C.mechanism("h3"))
    p, d, dxi, ev = g.mdinit(q, 2, xi_ideal=0.98, k_force=15.0)
    ep, xr, st = g.verlet(q, p, d, nsteps=3, constrain=1, xi_ideal=0.98, k_force=15.0, dxi=dxi, event=ev)
    assert st[0] == 0 and np.isfinite(q).all()
