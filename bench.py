#!/usr/bin/env python
"""bench.py -- RPMD bead-steps/s of the recrossing hot path on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[1]): calc_rate CH4+H on the CBE surface
(egrad_ch4h), 16 beads, T = 300 K, dt = 0.1 fs, 512 +/- child pairs = 1024 child trajectories
x child_evol = 1000 free steps per GPU, started from 8 constrained parent snapshots.  One bench
"step" = one pass of the work unit over that batch = 1024*16*1000 bead-steps.

Multi-GPU (one rank per GPU under torchrun): the job's NCCL communicator sits behind the C-ABI
(crcl_comm_init); crcl_recross_children(_dev) is then a collective call over the GLOBAL pair range, every
rank runs its contiguous block and the kappa(t) sums (child_evol+1 doubles) are all-reduced inside the
library.  Two legs are timed in every run:
  weak   (the line's `value`, "scaling": "weak"): 512 pairs PER GPU, the shape that keeps a B200 busy;
  strong (`strong_scaling` in the same line): BASELINE.json configs[1] as written, 1024 children IN TOTAL
         sharded over the N GPUs (128 children per GPU at N = 8).
--scaling strong makes the strong leg the line's `value` instead.

  python bench.py --gpus N --steps K --warmup W           product (one rank per GPU under torchrun)
  python bench.py --impl reference ...                    the CPU restatement of the reference path
                                                          on all host cores (the
                                                          Fortran reference cannot be built here)
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NBEADS, NPAIRS, CHILD_EVOL, NPARENT = 16, 512, 1000, 8
KELVIN, DT_FS, XI_DAG = 300.0, 0.1, 0.99
PARENT_EQUI, PARENT_INTERVAL = 300, 100
SEED = 20261017
METRIC = "RPMD bead-steps/s (traj x beads x steps)"
UNIT = "bead-steps/s"
WORKLOAD = ("calc_rate CH4+H (egrad_ch4h, CBE) 16 beads: %d recrossing child trajectories (%d +/- pairs) x %d steps "
            "per GPU from %d constrained parent snapshots, 300 K, dt 0.1 fs" % (2 * NPAIRS, NPAIRS, CHILD_EVOL, NPARENT))


def system():
    from caracal_b200 import systems as S
    from caracal_b200.api import beta_calc_rate, dt_au
    return S.masses("ch4h"), beta_calc_rate(KELVIN), dt_au(DT_FS), S.mechanism("ch4h"), S.ch5_ts()


def flops_per_bead_step():
    """Algorithmic flops of one bead-step of a child trajectory:
    one egrad_ch4h evaluation (oracle's dynamic census) + the bead transform in matrix form
    (24*N*natoms) + two half kicks (2*2*3*natoms)."""
    with open(os.path.join(ROOT, "oracle", "flop_census.json")) as f:
        pes = json.load(f)["ch4h"]["flops"]
    return pes + 24 * NBEADS * 6 + 2 * 2 * 3 * 6, pes


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def summary(self):
        rows = [r for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(rows[0][1]), "reasons": reasons, "samples": len(rows)}


def cpu_sample_sizes(cores, long=False):
    """bounded sample of the same workload for the CPU leg: whole +/- pairs per thread.
    long: ~10 s (cpu_baseline of the native arm); short: ~2 s per step (--impl reference)."""
    return (16 * cores, 1000) if long else (8 * cores, 500)


def run_cpu_reference(q_parents, steps, warmup, cores, long=False):
    """The oracle's recrossing work unit on all host cores: whole +/- pairs per worker, as the
    reference's MPI master/worker does (recross.f90:334-417,512-628)."""
    from oracle import oracle as O
    m, beta, dt, mech, _ = system()
    o = O.System("ch4h", NBEADS, m, beta, dt)
    o.set_mechanism(mech)
    npairs, evol = cpu_sample_sizes(cores, long)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        num, den, st = o.recross_children(q_parents, it * npairs, npairs, evol, XI_DAG, SEED, nthreads=cores)
        t1 = time.perf_counter()
        if it >= warmup:
            times.append(t1 - t0)
    bead_steps = 2 * npairs * NBEADS * evol
    sec = float(np.mean(times))
    return bead_steps / sec, sec, "%d +/- pairs x %d steps x %d beads per step on %d threads" % (npairs, evol, NBEADS, cores)


def make_parents_gpu(g):
    """8 constrained parent snapshots (recross.f90:240-297,344-440): equilibrate with SHAKE/RATTLE +
    Andersen, then one snapshot every PARENT_INTERVAL steps."""
    rng = np.random.default_rng(SEED)
    _, _, _, _, ts = system()
    g.set_thermostat(1, int(np.sqrt(PARENT_EQUI)), KELVIN)
    q = (ts[None, None] + rng.normal(0, 0.005, (1, NBEADS) + ts.shape)).copy()
    tid = np.array([4000000000], dtype=np.uint32)
    p, d, dxi, ev = g.mdinit(q, 2, XI_DAG, 0.0, traj_id=tid)
    g.verlet(q, p, d, nsteps=PARENT_EQUI, constrain=1, xi_ideal=XI_DAG, k_force=0.0, dxi=dxi, traj_id=tid, event=ev)
    snaps = []
    g.set_thermostat(1, int(np.sqrt(PARENT_INTERVAL)), KELVIN)
    done = PARENT_EQUI
    for _ in range(NPARENT):
        g.verlet(q, p, d, nsteps=PARENT_INTERVAL, istep0=done, constrain=1, xi_ideal=XI_DAG, k_force=0.0, dxi=dxi,
                 traj_id=tid, event=ev)
        done += PARENT_INTERVAL
        snaps.append(q[0].copy())
    return np.array(snaps)


def make_parents_cpu():
    """same protocol with the oracle (for --impl reference on a box where only the CPU leg runs)"""
    from oracle import oracle as O
    m, beta, dt, mech, ts = system()
    rng = np.random.default_rng(SEED)
    o = O.System("ch4h", NBEADS, m, beta, dt)
    o.set_mechanism(mech)
    o.q[:] = ts[None] + rng.normal(0, 0.005, (NBEADS,) + ts.shape)
    o.set_rng(SEED, 4000000000)
    o.set_thermostat(1, int(np.sqrt(PARENT_EQUI)), KELVIN)
    o.mdinit(XI_DAG, 2)
    for i in range(1, PARENT_EQUI + 1):
        o.verlet(i, XI_DAG, 1)
    o.set_thermostat(1, int(np.sqrt(PARENT_INTERVAL)), KELVIN)
    snaps, done = [], PARENT_EQUI
    for _ in range(NPARENT):
        for i in range(done + 1, done + PARENT_INTERVAL + 1):
            o.verlet(i, XI_DAG, 1)
        done += PARENT_INTERVAL
        snaps.append(o.q.copy())
    return np.array(snaps)


# ---- BASELINE.json configs[0], [2], [3], [4]: the same JSON line for the other configurations ---------------------------
def census(name):
    with open(os.path.join(ROOT, "oracle", "flop_census.json")) as f:
        return json.load(f)[name]["flops"]


def census_qmdff(key):
    with open(os.path.join(ROOT, "oracle", "flop_census_qmdff.json")) as f:
        return json.load(f)[key]


def other_config(args, emit, rank, world, local_rank, cores):
    """c1 H + H2 NVT (replicas), c3 OH + H2 64-bead recrossing children over the temperature sweep, c4 DG-EVB 20 atoms x
    32 beads, c5 periodic QMDFF box ~3000 atoms x 8 beads.  One bench step = one call of the work unit named in
    `workload`; value from device-resident state (crcl_verlet_dev / crcl_recross_children_dev), e2e through the
    host-pointer call.  Multi-GPU: independent replicas of the same work per rank (weak), no exchange (c1, c4, c5 do not
    shard: SURVEY.md 8e "replicas only"; c3 shards like c2, measured here per GPU)."""
    import numpy as np
    from caracal_b200 import systems as S
    from caracal_b200.api import atomic_mass_au, beta_calc_rate, beta_dynamic, dt_au
    cfg = args.config
    ref = args.impl == "reference"
    if ref and rank != 0:
        return 0
    O = None
    if ref or not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
    if not ref:
        import torch
        import torch.distributed as dist
        import caracal_b200
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU leg)")
        torch.cuda.set_device(local_rank)
        dev = torch.device("cuda", local_rank)
        if world > 1:
            dist.init_process_group("nccl", device_id=dev)
        caracal_b200.build_if_needed()
    rng = np.random.default_rng(SEED + rank)
    spec = {}
    # ------------------------------------------------------------------------------------------ set-up per configuration
    if cfg == "c1":
        name, nb, kelvin, ntraj, nsteps = "h3", 16, 300.0, 65536, 200
        mass, beta, dt = S.masses(name), beta_dynamic(kelvin), dt_au(0.1)
        fl = census("h3") + 24 * nb * 3 + 2 * 2 * 3 * 3
        spec = dict(kind="verlet", pes=name, nbeads=nb, natoms=3, ntraj=ntraj, nsteps=nsteps, constrain=-1, thermostat=(1, 70, kelvin),
                    flops_per_bead_step=fl, kernel="verlet_kernel<PesH3,16>",
                    workload="RPMD NVT trajectory on analytic H+H2 PES (egrad_h3), 16 beads, T=300 K, Andersen every 70 steps, dt 0.1 fs: "
                             "%d independent replicas x %d steps per bench step (a single trajectory does not shard; 10k steps = %d "
                             "bench steps)" % (ntraj, nsteps, 10000 // nsteps))
        q0 = np.array([S.ring_polymer(name, nb, rng, 0.02) for _ in range(64)])
        q0 = np.ascontiguousarray(np.resize(q0, (ntraj,) + q0.shape[1:]))
        cpu_steps = 200000
    elif cfg == "c3":
        name, nb, npairs, evol = "oh3", 64, 512, 500
        temps = [200.0, 300.0, 400.0, 600.0, 800.0, 1000.0]
        mass, dt = S.masses(name), dt_au(0.1)
        fl = census("oh3") + 24 * nb * 4 + 2 * 2 * 3 * 4
        spec = dict(kind="recross", pes=name, nbeads=nb, natoms=4, npairs=npairs, evol=evol, temps=temps, flops_per_bead_step=fl,
                    kernel="recross_kernel<PesOH3,64>",
                    workload="calc_rate OH+H2 (egrad_oh3) 64 beads, temperature sweep 200-1000 K: per bench step %d recrossing child "
                             "trajectories x %d steps at each of the %d temperatures" % (2 * npairs, evol, len(temps)))
    elif cfg == "c4":
        from caracal_b200.qmdff_synth import HEXANE, make_dgevb
        T1, T2, E = make_dgevb(seed=5, mode=3, npoints=7, template=HEXANE)
        nb, kelvin, ntraj, nsteps = 32, 300.0, 256, 50
        sym = {1: "H", 6: "C", 8: "O"}
        mass = np.array([atomic_mass_au(sym[int(z)]) for z in T1["at"]])
        beta, dt = beta_calc_rate(kelvin), dt_au(0.2)
        cz = census_qmdff("dgevb_hexane_mode3_7points")
        fl = cz["flops_per_image"] + 24 * nb * T1["n"] + 2 * 2 * 3 * T1["n"]
        spec = dict(kind="verlet", pes="dgevb", nbeads=nb, natoms=int(T1["n"]), ntraj=ntraj, nsteps=nsteps, constrain=-1,
                    thermostat=(1, 10, kelvin), flops_per_bead_step=fl, kernel="qm_bonded_kernel + dgevb_mix_kernel (split path)",
                    tables=(T1, T2, E),
                    workload="dG-EVB-QMDFF RPMD on a synthetic 20-atom two-state system (n-hexane-like QMDFF pair, 7 Gaussians, mode 3, "
                             "nat6 = 12), 32 beads, NVT Andersen: %d trajectories x %d steps per bench step" % (ntraj, nsteps))
        q0 = T1["xyz"][None, None] + rng.normal(0, 0.01, (ntraj, nb) + T1["xyz"].shape)
        cpu_steps = 1500
    else:
        from caracal_b200.qmdff_synth import make_system
        T = make_system(nmol=385, seed=12, periodic=True, zahn=True, hb=True)
        nb, kelvin, ntraj, nsteps = 8, 300.0, 1, 100
        sym = {1: "H", 6: "C", 8: "O", 17: "CL"}
        mass = np.array([atomic_mass_au(sym[int(z)]) for z in T["at"]])
        beta, dt = beta_dynamic(kelvin), dt_au(0.5)
        cz = census_qmdff("box_385_molecules_hb1")
        fl = cz["flops_per_image"] + 24 * nb * T["n"] + 2 * 2 * 3 * T["n"]
        spec = dict(kind="verlet", pes="qmdff", nbeads=nb, natoms=int(T["n"]), ntraj=ntraj, nsteps=nsteps, constrain=-1,
                    thermostat=(1, 70, kelvin), flops_per_bead_step=fl, kernel="qm_inter_cell_kernel (split path)", tables=(T,),
                    workload="periodic QMDFF box NVT (synthetic: 385 molecules = %d atoms, Zahn Coulomb, 10 A cut-offs, H bonds), 8 beads "
                             "RPMD: %d trajectory x %d steps per bench step" % (T["n"], ntraj, nsteps))
        q0 = T["xyz"][None, None] + rng.normal(0, 0.01, (ntraj, nb) + T["xyz"].shape)
        cpu_steps = 5
    nb, natoms = spec["nbeads"], spec["natoms"]
    config = {"workload": spec["workload"], "pes": spec["pes"], "natoms": natoms, "nbeads": nb,
              "transform": "reference (rfft/irfft as written)", "parallelism": "independent replicas per GPU, %d GPU(s)" % world,
              "dt_fs": {"c1": 0.1, "c3": 0.1, "c4": 0.2, "c5": 0.5}[cfg],
              "l2": "256 MiB device memset between timed steps (inside the timed region)",
              "flops_per_bead_step": spec["flops_per_bead_step"]}

    # ------------------------------------------------------------------------------------------ CPU restatement of the unit
    def cpu_leg():
        """-> (bead-steps/s, seconds, sample text, cores used)"""
        if spec["kind"] == "recross":
            o = O.System(spec["pes"], nb, mass, beta_calc_rate(300.0), dt)
            o.set_mechanism(S.mechanism(spec["pes"]))
            qp = np.array([S.ring_polymer(spec["pes"], nb, np.random.default_rng(k), 0.01) for k in range(8)])
            npr, ev = 8 * cores, 500
            t0 = time.perf_counter()
            o.recross_children(qp, 0, npr, ev, XI_DAG, SEED, nthreads=cores)
            sec = time.perf_counter() - t0
            return 2 * npr * nb * ev / sec, sec, "%d +/- pairs x %d steps x %d beads on %d threads (300 K)" % (npr, ev, nb, cores), cores
        o = O.System(spec["pes"] if cfg == "c1" else 0, nb, mass, beta, dt)
        if cfg == "c4":
            D = O.Dgevb(*spec["tables"])
            o.set_custom_grad(lambda xyz: tuple(a[0] for a in D.egrad(xyz)))
        elif cfg == "c5":
            Q = O.Qmdff(spec["tables"][0])
            o.set_custom_grad(lambda xyz: tuple(a[0] for a in Q.egrad(xyz)))
            o.set_box(True, spec["tables"][0]["box"])
        o.set_thermostat(*spec["thermostat"])
        o.set_rng(SEED, 0)
        o.q[:] = q0[0]
        o.mdinit(0.0, 0)
        t0 = time.perf_counter()
        for i in range(1, cpu_steps + 1):
            o.verlet(i, 0.0, -1)
        sec = time.perf_counter() - t0
        # one trajectory on one core: the reference's MPI gives dynamic.x no speed-up
        return nb * cpu_steps / sec, sec, "1 trajectory x %d steps x %d beads on 1 core" % (cpu_steps, nb), 1
    if ref:
        times = []
        for it in range(max(args.warmup, 1) + args.steps):
            v, sec, sample, used = cpu_leg()
            if it >= max(args.warmup, 1):
                times.append((v, sec))
            if sum(t[1] for t in times) > 120.0:
                break
        v = float(np.mean([t[0] for t in times]))
        emit({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
              "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * float(np.mean([t[1] for t in times])), "higher_is_better": True,
              "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
              "cpu_baseline": {"value": v, "unit": UNIT, "cores": used, "kind": "port", "sample": sample},
              "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
              "note": "C restatement (oracle/) of the reference's Fortran path (SURVEY.md F1)"})
        return 0

    # ------------------------------------------------------------------------------------------ product
    # everything (torch's flush, the library's launches, the timing events) on one non-default stream: the HBM-resident
    # path replays its steps from a CUDA graph, and a capture cannot start on the legacy default stream
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    T = lambda a, dt_=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt_, device=dev)
    if spec["kind"] == "recross":
        hs = []
        for kel in spec["temps"]:
            g = caracal_b200.RPMD(spec["pes"], nb, mass, beta_calc_rate(kel), dt, device=local_rank)
            g.set_mechanism(S.mechanism(spec["pes"]))
            g.set_seed(SEED)
            g.set_stream(stream.cuda_stream)
            hs.append(g)
        qp = np.array([S.ring_polymer(spec["pes"], nb, np.random.default_rng(k), 0.01) for k in range(8)])
        d_qp = T(qp)
        d_sums = torch.zeros(len(hs), spec["evol"] + 1, dtype=torch.float64, device=dev)
        bead_steps = len(hs) * 2 * spec["npairs"] * nb * spec["evol"]
        h2d, d2h = qp.nbytes * len(hs), len(hs) * (8 * (spec["evol"] + 1) + 4 * spec["npairs"])

        def step_dev(it):
            flush.zero_()
            for k, g in enumerate(hs):
                g.recross_children_dev(d_qp.data_ptr(), 8, spec["npairs"], spec["evol"], XI_DAG, d_sums[k].data_ptr(),
                                       d_sums[k].data_ptr() + 8 * spec["evol"], pair0=(rank * 1000 + it) * spec["npairs"])

        def step_host(it):
            for g in hs:
                g.recross_children(qp, spec["npairs"], spec["evol"], XI_DAG, pair0=(rank * 1000 + 500 + it) * spec["npairs"])
        g0 = hs[0]
    else:
        pid = {"h3": "h3", "dgevb": caracal_b200.PES_DGEVB, "qmdff": caracal_b200.PES_QMDFF}[spec["pes"]]
        g0 = caracal_b200.RPMD(pid, nb, mass, beta, dt, device=local_rank)
        if cfg == "c4":
            g0.set_qmdff(spec["tables"][0])
            g0.set_qmdff(spec["tables"][1], second=True)
            g0.set_dgevb(spec["tables"][2])
        elif cfg == "c5":
            g0.set_qmdff(spec["tables"][0])
        g0.set_seed(SEED)
        g0.set_thermostat(*spec["thermostat"])
        ntraj, nsteps = spec["ntraj"], spec["nsteps"]
        q = q0.copy()
        p, d, dxi, ev = g0.mdinit(q, 0)
        g0.set_stream(stream.cuda_stream)
        dq, dp, dd = T(q), T(p), T(d)
        dep, dxr = torch.zeros(ntraj, dtype=torch.float64, device=dev), torch.zeros(ntraj, dtype=torch.float64, device=dev)
        ddxi = torch.zeros(ntraj * natoms * 3, dtype=torch.float64, device=dev)
        dst = torch.zeros(ntraj, dtype=torch.int32, device=dev)
        dtid = torch.arange(ntraj, dtype=torch.int32, device=dev)
        dev_ = torch.as_tensor(ev.astype(np.int32), device=dev)
        bead_steps = ntraj * nb * nsteps
        h2d, d2h = 3 * q.nbytes, 3 * q.nbytes + 16 * ntraj
        hq, hp, hd = (torch.as_tensor(a).pin_memory().numpy() for a in (q, p, d))
        hev = ev.copy()
        done = [0]

        def step_dev(it):
            flush.zero_()
            g0.verlet_dev(ntraj, nsteps, dq.data_ptr(), dp.data_ptr(), dd.data_ptr(), dep.data_ptr(), dxr.data_ptr(), ddxi.data_ptr(),
                          dst.data_ptr(), dev_.data_ptr(), istep0=done[0], constrain=-1, d_traj_id=dtid.data_ptr())
            done[0] += nsteps

        def step_host(it):
            g0.verlet(hq, hp, hd, nsteps=nsteps, istep0=done[0], constrain=-1, event=hev)
            done[0] += nsteps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    handles = hs if spec["kind"] == "recross" else [g0]
    for it in range(args.warmup):
        step_dev(it)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = sum(g.launch_count() for g in handles)
    for g in handles:
        g.kernel_timings()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(args.steps):
        step_dev(args.warmup + it)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = sum(g.launch_count() for g in handles) - l0
    kms = np.concatenate([g.kernel_timings() for g in handles]) if spec["kind"] == "recross" else g0.kernel_timings()
    sampler.stop_flag.set()
    sampler.join()
    value = world * bead_steps * args.steps / (ms * 1e-3)
    nfailed = 0
    if spec["kind"] != "recross":
        st_host = dst.cpu().numpy()
        nfailed = int(((st_host & caracal_b200.lib.TRAJ_FATAL) != 0).sum())
        if nfailed > 0.05 * len(st_host):
            raise SystemExit("bench.py: %d of %d trajectories failed (status %s)" % (nfailed, len(st_host), np.unique(st_host)))
    # end to end: host buffers through the C-ABI
    barrier()
    for it in range(1):
        step_host(it)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, args.steps // 4)
    for it in range(e2e_steps):
        step_host(1 + it)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2e_val = world * bead_steps * e2e_steps / (e2e_ms * 1e-3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    peak = g0.measure_fp64_tflops(16384)
    step_ms = ms / args.steps
    if spec["kind"] == "recross" or not len(kms):
        kernel_ms = float(np.mean(kms)) if len(kms) else step_ms
        per_launch = bead_steps / max(len(handles), 1) if spec["kind"] == "recross" else bead_steps
    else:
        # fused path: one trajectory-kernel launch per step; split path: the step is many launches, use the whole step
        fused = cfg == "c1"
        kernel_ms = float(np.mean(kms)) if fused else step_ms
        per_launch = bead_steps
    achieved = per_launch * spec["flops_per_bead_step"] / (kernel_ms * 1e-3) / 1e12
    roofline = {"bound": "fp64", "kernel": spec["kernel"], "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak if peak > 0 else None, "traffic": None,
                "peak_source": "builder-measured: DFMA microbenchmark run in this process (MEASURED_PEAKS.json has no FP64 entry)",
                "algorithmic_flops_per_launch": per_launch * spec["flops_per_bead_step"],
                "note": "algorithmic flops = the reference's own operation count from the counting build of the restatement "
                        "(oracle/flop_census*.json; pairs outside the cut-offs not counted) + 24 N natoms for the transform"}
    cpu_baseline = None
    if not args.no_cpu_baseline:
        v, sec, sample, used = cpu_leg()
        cpu_baseline = {"value": v, "unit": UNIT, "cores": used, "kind": "port", "sample": sample}
    emit({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
          "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
          "config": config, "clocks": sampler.summary(),
          "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
          "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
          "kernel_ms": {"mean": kernel_ms, "n": int(len(kms)), "share_of_step": kernel_ms / step_ms},
          "failed_trajectories": nfailed})
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="which leg is the line's `value`: 512 pairs per GPU (weak) or 512 pairs in total (strong)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4", "c5"],
                    help="BASELINE.json configs[0..4]; c2 (default) is the configuration the metric is quoted on")
    args = ap.parse_args()
    # the contract is ONE JSON line on stdout: libraries that write to the C-level stdout (NCCL prints its version
    # banner there) are sent to stderr for the duration of the run, the line itself goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(obj) + "\n").encode())
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    if args.config != "c2":
        return other_config(args, emit, rank, world, local_rank, cores)
    fl_bs, fl_pes = flops_per_bead_step()
    config = {"workload": WORKLOAD, "pes": "ch4h", "natoms": 6, "nbeads": NBEADS, "child_pairs_per_gpu": NPAIRS,
              "child_steps": CHILD_EVOL, "parents": NPARENT, "kelvin": KELVIN, "dt_fs": DT_FS, "xi_ideal": XI_DAG,
              "transform": "reference (rfft/irfft as written)", "parallelism": "trajectory shards, %d GPU(s)" % world,
              "l2": "256 MiB device memset between timed steps (inside the timed region)",
              "flops_per_bead_step": fl_bs, "flops_per_egrad_ch4h": fl_pes}

    if args.impl == "reference":
        if rank != 0:
            return 0
        from oracle import oracle as O
        O.build()
        qp = make_parents_cpu()
        val, sec, sample = run_cpu_reference(qp, args.steps, max(args.warmup, 1), cores)
        emit({
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": max(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "C restatement (oracle/) of the reference's Fortran path; gfortran/MPI/FFTW are absent, the "
                    "reference itself cannot be built (SURVEY.md F1)"})
        return 0

    import torch
    import torch.distributed as dist
    import caracal_b200
    from caracal_b200.shard import comm_init_from_torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU leg)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    caracal_b200.build_if_needed()
    m, beta, dt, mech, ts = system()
    g = caracal_b200.RPMD("ch4h", NBEADS, m, beta, dt, device=local_rank)
    g.set_mechanism(mech)
    g.set_seed(SEED)
    qp = make_parents_gpu(g)               # identical on every rank (same seed and stream)
    stream = torch.cuda.current_stream()
    g.set_stream(stream.cuda_stream)
    if world > 1:
        # the job's communicator goes behind the C-ABI: from here on the work unit is a collective call over the global
        # pair range with the all-reduce of the kappa(t) sums inside the library (torch only ships the 128-byte id)
        comm_init_from_torch(g, device=dev)
    nccl_version = g.comm_info()[2]

    d_qp = torch.as_tensor(qp, device=dev).contiguous()
    d_sums = torch.zeros(CHILD_EVOL + 1, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def device_leg(npairs_global, steps, warmup, base):
        """inputs resident in HBM; K timed steps bracketed by barrier + synchronize, CUDA events, max over ranks"""
        def step(it):
            flush.zero_()
            g.recross_children_dev(d_qp.data_ptr(), NPARENT, npairs_global, CHILD_EVOL, XI_DAG, d_sums.data_ptr(),
                                   d_sums.data_ptr() + 8 * CHILD_EVOL, pair0=base + it * npairs_global)
        for it in range(warmup):
            step(it)
        barrier()
        l0 = g.launch_count()
        g.kernel_timings()                 # drop the warm-up launches from the event ring
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(steps):
            step(warmup + it)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        kms = g.kernel_timings()
        return dict(ms_per_step=ms / steps, value=2 * npairs_global * NBEADS * CHILD_EVOL * steps / (ms * 1e-3),
                    launches=g.launch_count() - l0, kernel_ms=kms, kappa_end=float(d_sums[CHILD_EVOL - 1] / d_sums[CHILD_EVOL]))

    def e2e_leg(npairs_global, steps, base):
        """the same work unit through the host-pointer C-ABI call: pinned host buffers, H2D / D2H and the all-reduce
        inside the timed region"""
        h_qp = torch.as_tensor(qp).pin_memory()
        qp_np = h_qp.numpy()
        barrier()
        for it in range(2):
            g.recross_children(qp_np, npairs_global, CHILD_EVOL, XI_DAG, pair0=base + it * npairs_global)
        barrier()
        t0 = time.perf_counter()
        for it in range(steps):
            num, den, st = g.recross_children(qp_np, npairs_global, CHILD_EVOL, XI_DAG, pair0=base + (2 + it) * npairs_global)
        barrier()
        ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        return 2 * npairs_global * NBEADS * CHILD_EVOL * steps / (ms * 1e-3)

    sampler = ClockSampler(local_rank)
    sampler.start()
    legs = {}
    main_leg = args.scaling
    np_glob = {"weak": world * NPAIRS, "strong": NPAIRS}
    legs[main_leg] = device_leg(np_glob[main_leg], args.steps, args.warmup, 0)
    sampler.stop_flag.set()
    sampler.join()
    other = "strong" if main_leg == "weak" else "weak"
    if world > 1:
        legs[other] = device_leg(np_glob[other], max(5, args.steps // 2), 3, 1 << 24)
    else:
        legs[other] = legs[main_leg]       # one GPU: the two shapes coincide
    e2e_steps = max(3, args.steps // 2)
    e2e_val = e2e_leg(np_glob[main_leg], e2e_steps, 1 << 26)
    e2e_other = e2e_leg(np_glob[other], e2e_steps, 1 << 27) if world > 1 else e2e_val

    if rank != 0:
        if world > 1:
            g.comm_destroy()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ------------------------------------------------------------
    L = legs[main_leg]
    kms = L["kernel_ms"]
    local_pairs = np_glob[main_leg] // world     # rank 0's block (blocks differ by at most one pair)
    bead_steps_launch = 2 * local_pairs * NBEADS * CHILD_EVOL
    kernel_ms = float(np.mean(kms)) if len(kms) else L["ms_per_step"]
    peak = g.measure_fp64_tflops(16384)
    achieved = bead_steps_launch * fl_bs / (kernel_ms * 1e-3) / 1e12
    # DRAM bytes of one launch of this kernel at this shape from an `ncu --set full` capture
    # (profiles/traffic_recross.json, written by profiles/ncu_traffic.py); null when not captured
    traffic = None
    hardware = None   # the same capture's pipe utilisation: the hardware's view beside the census-based fraction
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_recross.json")) as f:
            tr = json.load(f)
        if tr.get("child_steps") == CHILD_EVOL and tr.get("child_pairs") == local_pairs:
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            if "fp64_pipe_active_pct" in tr:
                hardware = {"fp64_pipe_active_frac": tr["fp64_pipe_active_pct"] / 100.0,
                            # the free ring-polymer step runs as mma.sync.m8n8k4.f64 on the tensor sub-pipe, which
                            # sm__pipe_fp64_cycles_active does not count
                            "dmma_pipe_active_frac": tr.get("dmma_pipe_active_pct", 0.0) / 100.0,
                            "issue_active_frac": tr.get("issue_active_pct", 0.0) / 100.0,
                            "warp_instructions_per_launch": tr.get("warp_instructions"),
                            "source": "ncu --set full, profiles/" + tr.get("source", "traffic_recross.json")}
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"bound": "fp64", "kernel": "recross_kernel<PesCH4H,16>", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak if peak > 0 else None, "traffic": traffic,
                "peak_source": "builder-measured: DFMA microbenchmark run in this process (MEASURED_PEAKS.json has no FP64 entry)",
                "algorithmic_flops_per_launch": bead_steps_launch * fl_bs, "hardware": hardware,
                "note": "algorithmic flops = reference's own operation count (oracle census: 411 libm calls, 1813 "
                        "divisions per image as written); the kernel executes about a quarter of them (re-derived PES), "
                        "so frac overstates the pipe utilisation -- that is hardware.fp64_pipe_active_frac (ncu)"}
    cpu_baseline = None
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        v, sec, sample = run_cpu_reference(qp, 1, 0, cores, long=True)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    def leg_summary(name, e2e):
        X = legs[name]
        return {"value": X["value"], "ms_per_step": X["ms_per_step"], "e2e": e2e, "child_pairs_total": np_glob[name],
                "children_per_gpu": 2 * np_glob[name] // world, "kappa_end": X["kappa_end"]}
    out = {"metric": METRIC, "value": L["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": L["ms_per_step"], "higher_is_better": True, "scaling": main_leg, "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": config, "clocks": sampler.summary(),
           "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(qp.nbytes),
                   "d2h_bytes_per_step": int(8 * (CHILD_EVOL + 1) + 4 * np_glob[main_leg])},
           "gpu_launches": int(L["launches"]), "roofline": roofline, "cpu_baseline": cpu_baseline,
           "kernel_ms": {"mean": kernel_ms, "n": int(len(kms)), "share_of_step": kernel_ms / L["ms_per_step"]},
           "weak_scaling": leg_summary("weak", e2e_val if main_leg == "weak" else e2e_other),
           "strong_scaling": leg_summary("strong", e2e_val if main_leg == "strong" else e2e_other),
           "collective": {"where": "inside libcaracal_gpu.so (crcl_comm_init: ncclAllReduce of child_evol+1 doubles per step)",
                          "nccl_version": nccl_version, "ranks": world}}
    emit(out)
    if world > 1:
        g.comm_destroy()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
