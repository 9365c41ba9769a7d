"""GPU parity of the round-2 additions, each through the C-ABI against the oracle:
  * periodic wrap of verlet.f90:591-641 (fused into the propagation kernel of the HBM-resident path): the shipped ethanol
    box for 100 steps with atoms started on a face, and a zero-force host-callback system against plain Python loops;
  * rpmd_check.f90:88-116 as status bits, fused and split path;
  * Andersen stream continuity over single-step calls, host-callback PES in crcl_umbrella_windows;
  * the NCCL communicator behind the C-ABI on one rank (the multi-rank form runs under torchrun: tests/multi_gpu_comm.py).
"""
import ctypes as _ct

import numpy as np
import pytest

from tests import common as C
from tests import qmdff_file as QF
from tests.test_gpu_qmdff_examples import masses
from tests.test_independent_checks import free_rp_np, wrap_py

pytestmark = pytest.mark.gpu


def test_ethanol_box_100_steps_with_atoms_crossing_faces(gpu, oracle):
    """config 5's shape in miniature, with the wrap: 2 beads, 100 steps; the shipped start structure has atoms exactly on
    the x = 0, y = 0, z = 0 faces, so coordinates are shifted from step 1 on; q within 1e-8 of the oracle"""
    T = QF.tables("box", periodic_angstrom=[27.0, 27.0, 27.0])
    nb, nsteps = 2, 100
    m = masses(T)
    beta, dt = C.beta_calc_rate(200.0), C.dt_au(0.5)
    g = gpu.RPMD(gpu.PES_QMDFF, nb, m, beta, dt)
    g.set_qmdff(T)                                        # periodic tables switch the wrap on (pbc_mod is one global)
    g.set_seed(C.SEED)
    g.set_thermostat(1, 30, 200.0)
    rng = np.random.default_rng(5)
    x0 = QF.box_start_bohr()
    assert x0.min() == 0.0 and (x0 < 0.05).sum() >= 3     # atoms within 0.05 bohr of a face
    q0 = x0[None, None] + rng.normal(0, 0.01, (1, nb, T["n"], 3))
    q = q0.copy()
    p, d, dxi, ev = g.mdinit(q, 0)
    ep, xr, st = g.verlet(q, p, d, nsteps=nsteps, constrain=-1, event=ev)
    Q = oracle.Qmdff(T)
    o = oracle.System(0, nb, m, beta, dt)
    o.set_custom_grad(lambda xyz: tuple(a[0] for a in Q.egrad(xyz)))
    o.set_box(True, T["box"])
    o.q[:] = q0[0]
    o.set_rng(C.SEED, 0)
    o.set_thermostat(1, 30, 200.0)
    o.mdinit(0.0, 0)
    for i in range(1, nsteps + 1):
        epo, _, sto = o.verlet(i, 0.0, -1)
        assert sto == 0
    assert st[0] == 0
    moved = np.abs(o.q - q0[0]).max(axis=0) > 0.5 * T["box"][0]
    assert moved.sum() >= 3, "no atom wrapped"            # several coordinates went through a face
    # a ring polymer that straddles a face is shifted back and forth bead by bead (verlet.f90:593-640 visits the beads in
    # order and every shift moves all of them), so beads end within the polymer's width of the box, not strictly inside
    assert (o.q >= -0.5).all() and (o.q <= np.asarray(T["box"]) + 0.5).all()
    assert np.abs(q[0] - o.q).max() < C.TOL_QP
    assert (np.abs(p[0] - o.p) / np.abs(o.p).max()).max() < C.TOL_QP
    assert abs(ep[0] - epo) < 1e-9 * max(1.0, abs(epo))


@pytest.mark.parametrize("nb", [1, 6, 16, 24])
def test_wrap_on_a_zero_potential_against_python_loops(gpu, nb):
    """register kernel (<= 16 beads) and shared-memory kernel (24): q after one step == numpy free ring polymer +
    the loops of verlet.f90:591-641, incl. ring polymers straddling a face and coordinates several boxes away"""
    natoms, ntraj = 5, 64
    rng = np.random.default_rng(nb)
    mass = np.array([C.atomic_mass_au(s) for s in ["C", "H", "O", "H", "H"]])
    beta, dt = C.beta_calc_rate(300.0), C.dt_au(0.1)
    box = np.array([9.0, 7.5, 11.0])
    g = gpu.RPMD(gpu.PES_HOSTCB, nb, mass, beta, dt)
    g.set_host_gradient(lambda x: (0.0, np.zeros_like(x)))
    g.set_box(True, box)
    q0 = rng.uniform(-0.3, 0.3, (ntraj, nb, natoms, 3)) + rng.choice([0.0, 1.0], (ntraj, 1, natoms, 3)) * box + \
        rng.choice([0.0, 0.0, 0.0, 2.0, -3.0], (ntraj, 1, natoms, 3)) * box
    p0 = rng.normal(0, 2.0, q0.shape)
    q, p, d = q0.copy(), p0.copy(), np.zeros_like(q0)
    ep, xr, st = g.verlet(q, p, d, nsteps=1, constrain=-1)
    assert (st == 0).all()
    nshift = 0
    for t in range(ntraj):
        if nb == 1:
            qf, pf = q0[t] + p0[t] * dt / mass[None, :, None], p0[t]
        else:
            qf, pf = free_rp_np(q0[t], p0[t], mass, beta, dt)
        qw, fatal = wrap_py(qf, box)
        assert not fatal
        nshift += int((np.abs(qw - qf) > 1.0).any(axis=0).sum())
        assert np.abs(q[t] - qw).max() < 1e-11
    assert nshift > ntraj
    # the give-up path: status bit, no abort, other trajectories untouched
    q, p = q0.copy(), np.zeros_like(q0)
    q[3, :, 0, 0] -= 200 * box[0]
    ep, xr, st = g.verlet(q, p, d, nsteps=1, constrain=-1)
    assert st[3] & gpu.lib.TRAJ_PBC_FAIL and (np.delete(st, 3) == 0).all()
    # switched off again: nothing is shifted
    g.set_box(False)
    q = q0.copy()
    g.verlet(q, np.zeros_like(q0), d, nsteps=1, constrain=-1)
    assert np.abs(q - q0).max() < 1e-9 or nb > 1


@pytest.mark.parametrize("path", ["fused", "split"])
def test_rpmd_check_status_bits_match_oracle(gpu, oracle, path):
    name, nb, ntraj, nsteps = "h3", 8, 6, 40
    rng = np.random.default_rng(17)
    q0 = np.array([C.ring_polymer(name, nb, rng, 0.02) for _ in range(ntraj)])
    tid = np.arange(50, 50 + ntraj, dtype=np.uint32)
    xi0 = np.array([0.98, 0.98, 0.98, 0.5, 0.98, 0.7])     # trajectories 3 and 5 start far from their window
    e_ts = float(gpu.egrad(name, C.h3_ts()[None])[0][0])
    # window 0.98 with a loose and a tight energy tolerance; the tight one trips on thermal energy
    for e_tol, xi_tol in ((0.4, 0.1), (0.0008, 0.1), (0.4, 0.004)):
        g, _ = C.make_pair(name, nb)
        g.set_path(gpu.PATH_FUSED if path == "fused" else gpu.PATH_SPLIT)
        g.set_seed(C.SEED)
        g.set_thermostat(1, 11, 300.0)
        g.set_rpmd_check(True, e_ts, e_tol, xi_tol)
        q = q0.copy()
        kf = np.full(ntraj, 15.0)
        p, d, dxi, ev = g.mdinit(q, 2, xi0, kf, traj_id=tid)
        ep, xr, st = g.verlet(q, p, d, nsteps=nsteps, constrain=0, xi_ideal=xi0, k_force=kf, dxi=dxi, traj_id=tid, event=ev)
        sto = np.zeros(ntraj, dtype=np.int32)
        for t in range(ntraj):
            _, o = C.make_pair(name, nb)
            o.q[:] = q0[t]
            o.set_rng(C.SEED, int(tid[t]))
            o.set_thermostat(1, 11, 300.0)
            o.set_kforce(15.0)
            o.set_rpmd_check(True, e_ts, e_tol, xi_tol)
            o.mdinit(float(xi0[t]), 2)
            for i in range(1, nsteps + 1):
                code = o.verlet(i, float(xi0[t]), 0)[2]
                sto[t] |= code
                if code and path == "fused":
                    break                                   # the fused kernels freeze a failed trajectory
        fatal = gpu.lib.TRAJ_FATAL
        if path == "fused":
            assert ((st & fatal) == (sto & fatal)).all(), (st, sto)
        else:
            # the HBM-resident path flags and carries on: every bit the oracle saw up to its first failure is set
            assert ((st & fatal) != 0).tolist() == ((sto & fatal) != 0).tolist(), (st, sto)
        if (e_tol, xi_tol) == (0.4, 0.1):
            assert (st[[0, 1, 2, 4]] & fatal == 0).all() and st[3] & gpu.lib.TRAJ_XI_RANGE and st[5] & gpu.lib.TRAJ_XI_RANGE
        elif e_tol < 0.001:
            assert (st & gpu.lib.TRAJ_ENERGY).any()
        else:
            assert (st & gpu.lib.TRAJ_XI_RANGE).all()


def test_andersen_draws_continue_over_single_step_calls(gpu):
    """a driver that calls crcl_verlet one step at a time must hand the event counter back in, or every
    Andersen resample repeats the first draw.  With event in/out two successive resamples differ; without it they repeat."""
    name, nb = "h3", 4
    g, _ = C.make_pair(name, nb)
    g.set_seed(C.SEED)
    g.set_thermostat(1, 1, 300.0)                           # resample on every step
    q = np.array([C.ring_polymer(name, nb, np.random.default_rng(1), 0.02)])
    tid = np.array([9], dtype=np.uint32)
    p, d, dxi, ev = g.mdinit(q, 0, traj_id=tid)
    assert ev[0] == 1
    g.verlet(q, p, d, nsteps=1, istep0=0, constrain=2, xi_ideal=0.98, traj_id=tid, event=ev)   # children: no draw
    assert ev[0] == 1
    g.set_mechanism(C.mechanism(name))
    pa, pb = [], []
    for i in range(3):
        g.verlet(q, p, d, nsteps=1, istep0=i, constrain=3, xi_ideal=0.98, k_force=0.0, traj_id=tid, event=ev)
        pa.append(p.copy())
    assert ev[0] == 4
    assert np.abs(pa[0] - pa[1]).max() > 1e-3 and np.abs(pa[1] - pa[2]).max() > 1e-3
    for i in range(2):                                       # the failure mode: no counter -> the same stream element
        g.verlet(q, p, d, nsteps=1, istep0=i, constrain=3, xi_ideal=0.98, k_force=0.0, traj_id=tid)
        pb.append(p.copy())
    assert np.abs(pb[0] - pb[1]).max() == 0.0


def test_umbrella_windows_with_host_callback_pes(gpu, oracle):
    """the split branch of crcl_umbrella_windows recomputes the plain forces through the same
    dispatch as the steps (host callback included).  H + H2 with the oracle's BKMP2 as the user's custom_grad."""
    from tests.oracle_handle import OracleRPMD
    name, nb = "h3", 4
    m, beta, dt = C.masses(name), C.beta_calc_rate(300.0), C.dt_au(0.1)
    g = gpu.RPMD(gpu.PES_HOSTCB, nb, m, beta, dt)
    g.set_host_gradient(lambda x: tuple(a[0] for a in oracle.egrad(name, x[None])[:2]))
    g.set_mechanism(C.mechanism(name))
    g.set_seed(C.SEED)
    g.set_thermostat(1, 7, 300.0)
    xi = np.array([0.95, 1.0])
    q0 = np.array([C.ring_polymer(name, nb, np.random.default_rng(k), 0.0) for k in range(2)])
    avg, var, st = g.umbrella_windows(q0, xi, np.full(2, 15.0), 2, 10, 20, traj_id0=77)
    o = OracleRPMD(name, nb, m, beta, dt)
    o.set_mechanism(C.mechanism(name))
    o.set_seed(C.SEED)
    o.set_thermostat(1, 7, 300.0)
    ao, vo, so = o.umbrella_windows(q0, xi, np.full(2, 15.0), 2, 10, 20, traj_id0=77)
    assert (st == 0).all() and (so == 0).all()
    assert np.abs(avg - ao).max() < 1e-9 and np.abs(var - vo).max() < 1e-9


def test_single_rank_communicator(gpu):
    """crcl_comm_init with one rank: the collective form of the work units returns what the plain form returns"""
    name, nb = "ch4h", 16
    g, _ = C.make_pair(name, nb)
    g.set_seed(C.SEED)
    qp = np.array([C.ring_polymer(name, nb, np.random.default_rng(k), 0.01) for k in range(2)])
    num0, den0, st0 = g.recross_children(qp, 8, 30, 0.98, pair0=5)
    nr, rk, ver = g.comm_info()
    assert (nr, rk) == (1, 0) and ver >= 21800, ver
    g.comm_init(1, 0, gpu.RPMD.comm_unique_id())
    assert g.comm_info()[:2] == (1, 0)
    num1, den1, st1 = g.recross_children(qp, 8, 30, 0.98, pair0=5)
    assert den1 == den0 and (num1 == num0).all() and (st1 == st0).all()
    g.set_thermostat(1, 7, 300.0)
    q0 = np.array([C.ring_polymer(name, nb, np.random.default_rng(9), 0.0)])
    a1, v1, s1 = g.umbrella_windows(q0, [0.9], [15.0], 3, 5, 10, traj_id0=3)
    g.comm_destroy()
    a0, v0, s0 = g.umbrella_windows(q0, [0.9], [15.0], 3, 5, 10, traj_id0=3)
    assert (a1 == a0).all() and (v1 == v0).all() and (s1 == s0).all()


def test_multi_rank_communicator(gpu):
    """the collective form over all GPUs of the box (torchrun, one rank per GPU); skipped on a one-GPU box"""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU: the multi-rank form is covered by tests/multi_gpu_comm.py under gpurun --gpus N")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                          "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(root, "tests", "multi_gpu_comm.py")],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert res.stdout.count(" ok: kappa") == n


@pytest.mark.parametrize("nmol,nimg", [(125, 3), (385, 2)])
def test_cell_sweep_equals_n2_sweep_and_oracle(gpu, oracle, nmol, nimg):
    """the cell-binned sweep of the inter-molecular part (qm_inter_cell_kernel, M = 3 and 2) against
    the O(N^2) sweep it replaces (CRCL_QM_CELLS=0) and, for the smaller box, the oracle: same pairs, same bits per pair,
    only the accumulation order differs -> 1e-12 relative between the sweeps, 1e-10 against the oracle.  Atoms are pushed
    out of the box and onto faces so that the wrapped coordinates and the shifts of the wrap segments are exercised."""
    import os
    from tests.qmdff_synth import make_system
    from tests.test_gpu_qmdff import handle, torsion_conditioning
    T = make_system(nmol=nmol, seed=21, periodic=True, zahn=True, hb=False)
    rng = np.random.default_rng(4)
    x = T["xyz"][None] + rng.normal(0, 0.06, (nimg,) + T["xyz"].shape)
    box = np.asarray(T["box"])
    x[0] += 0.37 * box                                          # whole image displaced: most atoms leave the box
    x[-1] -= np.floor(x[-1].min(axis=0) / box) * box + x[-1].min(axis=0) % box   # the lowest atom exactly on the faces
    res = {}
    for key, env in (("m3", {"CRCL_QM_CELLS": "1", "CRCL_QM_CELL_M": "3"}), ("m2", {"CRCL_QM_CELLS": "1", "CRCL_QM_CELL_M": "2"}),
                     ("n2", {"CRCL_QM_CELLS": "0"})):
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            g, _ = handle(gpu, T)
            res[key] = g.egrad(x)[:2]
            g.close()
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    V0, g0 = res["n2"]
    for key in ("m3", "m2"):
        V, gr = res[key]
        assert np.abs(V - V0).max() < 1e-12 * np.abs(V0).max(), key
        assert C.rel_err_G(gr.reshape(nimg, -1, 3), g0.reshape(nimg, -1, 3)).max() < 1e-12, key
    if nmol <= 125:
        Vo, go = oracle.Qmdff(T).egrad(x)
        assert C.rel_err_E(res["m3"][0], Vo).max() < C.TOL_EG
        tol = np.maximum(C.TOL_EG, 2e-17 / torsion_conditioning(T, x) ** 2)
        assert (C.rel_err_G(res["m3"][1].reshape(go.shape), go) < tol).all()


@pytest.mark.parametrize("nb", [16, 64])
def test_tensor_core_transform_equals_the_fma_transform(gpu, nb):
    """crcl_bench_transform: the free ring-polymer step as mma.sync.m8n8k4.f64 tiles against the FMA loop of the
    trajectory kernels on the same seeded trajectories, three steps: same numbers to rounding (different summation
    order only), and the step is not the identity.  The fused form of the same tiles (Traj::free_rp_dmma) is covered by
    every lane-split case of tests/test_gpu_verlet.py / test_gpu_recross.py against the oracle."""
    g, _ = C.make_pair("ch4h", 16)
    r = g.bench_transform(nb, 64, 2)
    assert r["max_rel_diff"] < 1e-14
    assert r["max_dq"] > 1e-3
    assert r["ms_dfma"] > 0 and r["ms_dmma"] > 0


def test_egrad_unaligned_device_arrays(gpu, oracle):
    """crcl_egrad_dev on device arrays that are only 8-byte aligned (offset by one double into their allocations)"""
    import torch
    name = "clnh3"
    q = C.ts_cloud(name, 333, 0.1, np.random.default_rng(4))
    Vo, go, _ = oracle.egrad(name, q)
    g, _ = C.make_pair(name, 1)
    n = q.size
    buf_q = torch.zeros(n + 1, dtype=torch.float64, device="cuda")
    buf_g = torch.zeros(n + 1, dtype=torch.float64, device="cuda")
    V = torch.zeros(len(q), dtype=torch.float64, device="cuda")
    buf_q[1:] = torch.from_numpy(q.ravel()).cuda()
    torch.cuda.synchronize()
    rc = g._lib.crcl_egrad_dev(g._h, gpu.PES_CLNH3, _ct.c_void_p(buf_q.data_ptr() + 8), 5, len(q),
                               _ct.c_void_p(V.data_ptr()), _ct.c_void_p(buf_g.data_ptr() + 8), None)
    assert rc == 0
    g.synchronize()
    assert C.rel_err_E(V.cpu().numpy(), Vo).max() < C.TOL_EG
    assert C.rel_err_G(buf_g[1:].cpu().numpy().reshape(go.shape), go).max() < C.TOL_EG
