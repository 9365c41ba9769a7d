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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="which leg is the line's `value`: 512 pairs per GPU (weak) or 512 pairs in total (strong)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_recross.json")) as f:
            tr = json.load(f)
        if tr.get("child_steps") == CHILD_EVOL and tr.get("child_pairs") == local_pairs:
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"bound": "fp64", "kernel": "recross_kernel<PesCH4H,16>", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak if peak > 0 else None, "traffic": traffic,
                "peak_source": "builder-measured: DFMA microbenchmark run in this process (MEASURED_PEAKS.json has no FP64 entry)",
                "algorithmic_flops_per_launch": bead_steps_launch * fl_bs,
                "note": "algorithmic flops = reference's own operation count (oracle census); the kernel executes "
                        "fewer (re-derived PES), see DESIGN.md"}
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
