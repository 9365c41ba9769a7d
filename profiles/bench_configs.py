"""Throughput of the hot path at the shapes of BASELINE.json configs[0], [2], [3], [4] (configs[1] is
bench.py), each next to the CPU oracle timed on a bounded sample of the same work on ONE host core
(dynamic.x gets nothing from MPI for these surfaces; the work-unit configs scale with the
worker count, see bench.py's cpu_baseline for that).  Wall-clock around the C-ABI call with host
buffers, i.e. end to end.  Run on the GPU box:  python profiles/bench_configs.py [out.json]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import caracal_b200  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests import common as C  # noqa: E402
from tests.qmdff_synth import make_dgevb, make_system  # noqa: E402

O.build()
rows = []


def timed(fn, reps=3):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return min(t)


def report(**kw):
    kw["gpu_over_one_core"] = kw["gpu_bead_steps_per_s"] / kw["cpu_bead_steps_per_s_one_core"]
    rows.append(kw)
    print(json.dumps(kw), flush=True)


# ---- C1: RPMD NVT on H + H2 (egrad_h3), 16 beads, Andersen every 70 steps (dynamic.x) ---------------
name, nb, nsteps = "h3", 16, 1000
g, o = C.make_pair(name, nb)
g.set_seed(C.SEED)
g.set_thermostat(1, 70, 300.0)
rng = np.random.default_rng(1)
o.q[:] = C.ring_polymer(name, nb, rng, 0.02)
o.set_rng(C.SEED, 1)
o.set_thermostat(1, 70, 300.0)
o.mdinit(0.0, 0)
t0 = time.perf_counter()
for i in range(1, 2001):
    o.verlet(i, 0.0, -1)
cpu = 2000 * nb / (time.perf_counter() - t0)
for ntraj in (1, 1024, 65536):
    q = np.array([C.ring_polymer(name, nb, rng, 0.02) for _ in range(min(ntraj, 64))])
    q = np.ascontiguousarray(np.resize(q, (ntraj,) + q.shape[1:]))
    p, d, dxi, ev = g.mdinit(q, 0)
    sec = timed(lambda: g.verlet(q, p, d, nsteps=nsteps, constrain=-1, event=ev), reps=2)
    report(config="C1 dynamic.x H+H2 (egrad_h3) 16 beads NVT Andersen", ntraj=ntraj, steps=nsteps,
           gpu_bead_steps_per_s=ntraj * nb * nsteps / sec, cpu_bead_steps_per_s_one_core=cpu)
g.close()

# ---- C2 umbrella phase: CH4+H 16 beads, 111 windows x 10 trajectories x (10k + 20k) steps ----
name, nb = "ch4h", 16
g, o = C.make_pair(name, nb)
g.set_seed(C.SEED)
g.set_thermostat(1, 70, 300.0)
xi0 = np.linspace(-0.05, 1.05, 111)
kf = np.full(111, 0.05 * 300.0)
q0 = np.array([C.ring_polymer(name, nb, rng, 0.01) for _ in range(111)])
equi, samp, ntw = 10000, 20000, 10
t0 = time.perf_counter()
avg, var, st = g.umbrella_windows(q0, xi0, kf, ntw, equi, samp)
sec = time.perf_counter() - t0
o.set_thermostat(1, 70, 300.0)
t0 = time.perf_counter()
o.q[:] = q0[55]
o.set_rng(C.SEED, 0)
o.set_kforce(float(kf[55]))
o.mdinit(float(xi0[55]), 2)
for i in range(1, 301):
    o.verlet(i, float(xi0[55]), 0)
cpu = 300 * nb / (time.perf_counter() - t0)
report(config="C2 calc_rate CH4+H (egrad_ch4h) 16 beads umbrella phase: 111 windows x 10 trajectories x (10k+20k) steps",
       ntraj=111 * ntw, steps=equi + samp, gpu_bead_steps_per_s=111 * ntw * nb * (equi + samp) / sec, gpu_seconds=sec,
       frac_status0=float((st == 0).mean()),
       cpu_bead_steps_per_s_one_core=cpu)
g.close()

# ---- C3: calc_rate OH + H2 (egrad_oh3), 64 beads, recrossing children, temperature sweep ------------
name, nb, npairs, evol = "oh3", 64, 512, 500
for kelvin in (200.0, 300.0, 1000.0):
    g, o = C.make_pair(name, nb, kelvin=kelvin)
    g.set_seed(C.SEED)
    qp = np.array([C.ring_polymer(name, nb, rng, 0.01) for _ in range(8)])
    sec = timed(lambda: g.recross_children(qp, npairs, evol, 0.98), reps=2)
    t0 = time.perf_counter()
    o.recross_children(qp, 0, 2, 100, 0.98, C.SEED, nthreads=1)
    cpu = 2 * 2 * nb * 100 / (time.perf_counter() - t0)
    report(config="C3 calc_rate OH+H2 (egrad_oh3) 64 beads, %d children x %d steps, %d K" % (2 * npairs, evol, kelvin),
           ntraj=2 * npairs, steps=evol, gpu_bead_steps_per_s=2 * npairs * nb * evol / sec,
           cpu_bead_steps_per_s_one_core=cpu)
    g.close()

# ---- C4: DG-EVB-QMDFF RPMD, 32 beads, on the split path -------------------------------------------------
# (the only DG-EVB fixture the reference ships is the 6-atom examples/evbopt/DG-EVB, whose evb_pars.dat needs
# evbopt.x; synthetic two-state systems: 9 atoms (ethanol-like) and SURVEY 8(d)'s 20 atoms / nat6 = 12 / 7 points)
from tests.qmdff_synth import HEXANE  # noqa: E402
for label, tpl in (("9 atoms", None), ("20 atoms, nat6 12", HEXANE)):
    T1, T2, E = make_dgevb(seed=5, mode=3, npoints=7, template=tpl)
    nb, ntraj, nsteps = 32, 256, 100
    mass = np.array([C.atomic_mass_au({1: "H", 6: "C", 8: "O"}[int(z)]) for z in T1["at"]])
    beta, dt = C.beta_calc_rate(300.0), C.dt_au(0.2)
    g = caracal_b200.RPMD(caracal_b200.PES_DGEVB, nb, mass, beta, dt)
    g.set_qmdff(T1)
    g.set_qmdff(T2, second=True)
    g.set_dgevb(E)
    g.set_seed(C.SEED)
    g.set_thermostat(1, 10, 300.0)
    q = np.ascontiguousarray(T1["xyz"][None, None] + rng.normal(0, 0.01, (ntraj, nb) + T1["xyz"].shape))
    p, d, dxi, ev = g.mdinit(q, 0)
    secs = {}
    for graph in (0, 1):
        g.set_graph(graph)
        secs[graph] = timed(lambda: g.verlet(q, p, d, nsteps=nsteps, constrain=-1, event=ev), reps=2)
    D = O.Dgevb(T1, T2, E)
    os_ = O.System(0, nb, mass, beta, dt)
    os_.set_custom_grad(lambda x: tuple(a[0] for a in D.egrad(x)))
    os_.q[:] = q[0]
    os_.set_rng(C.SEED, 0)
    os_.set_thermostat(1, 10, 300.0)
    os_.mdinit(0.0, 0)
    t0 = time.perf_counter()
    for i in range(1, 6):
        os_.verlet(i, 0.0, -1)
    cpu = 5 * nb / (time.perf_counter() - t0)
    for graph in (0, 1):
        report(config="C4 DG-EVB-QMDFF RPMD (2 x QMDFF + mode-3 coupling, %s synthetic) 32 beads, split path, " % label
                      + ("CUDA-graph replay" if graph else "one launch per kernel"),
               ntraj=ntraj, steps=nsteps, gpu_bead_steps_per_s=ntraj * nb * nsteps / secs[graph],
               gpu_us_per_step=1e6 * secs[graph] / nsteps, cpu_bead_steps_per_s_one_core=cpu)
    g.close()

# ---- C5: periodic QMDFF box NVT (~3000 atoms, Zahn, H bonds), classical and 8 beads ------------------
T = make_system(nmol=385, seed=12, periodic=True, zahn=True, hb=True)
mass = np.array([C.atomic_mass_au({1: "H", 6: "C", 8: "O", 17: "CL"}[int(z)]) for z in T["at"]])
Q = O.Qmdff(T)
t0 = time.perf_counter()
Q.egrad(T["xyz"][None])
cpu_img = time.perf_counter() - t0
for nb, nsteps in ((1, 50), (8, 50)):
    g = caracal_b200.RPMD(caracal_b200.PES_QMDFF, nb, mass, C.beta_calc_rate(300.0), C.dt_au(0.5))
    g.set_qmdff(T)
    g.set_seed(C.SEED)
    g.set_thermostat(1, 70, 300.0)
    q = np.ascontiguousarray(T["xyz"][None, None] + rng.normal(0, 0.01, (1, nb) + T["xyz"].shape))
    p, d, dxi, ev = g.mdinit(q, 0)
    sec = timed(lambda: g.verlet(q, p, d, nsteps=nsteps, constrain=-1, event=ev), reps=2)
    report(config="C5 periodic QMDFF box NVT, %d atoms, Zahn + H bonds, %d bead(s), split path" % (T["n"], nb), ntraj=1,
           steps=nsteps, gpu_bead_steps_per_s=nb * nsteps / sec, gpu_ms_per_step=1e3 * sec / nsteps,
           cpu_bead_steps_per_s_one_core=1.0 / cpu_img)
    g.close()

# ---- C5 (native water model): flexible SPC box, pes WATER_SPC, 1000 molecules, 31.07 A, Zahn ---------------
from caracal_b200 import water as WT  # noqa: E402
Wt = WT.water_box(1000, periodic_angstrom=[31.07] * 3)
xw = WT.water_lattice(1000, 31.07, rng)
mass = np.tile([C.atomic_mass_au("O"), C.atomic_mass_au("H"), C.atomic_mass_au("H")], 1000)
Qw = O.Water(Wt)
t0 = time.perf_counter()
Qw.egrad(xw[None])
cpu_img = time.perf_counter() - t0
for nb, nsteps in ((1, 200), (8, 200)):
    g = caracal_b200.RPMD(caracal_b200.PES_WATER, nb, mass, C.beta_calc_rate(300.0), C.dt_au(0.5))
    g.set_water(Wt)
    g.set_seed(C.SEED)
    g.set_thermostat(1, 70, 300.0)
    q = np.ascontiguousarray(xw[None, None] + rng.normal(0, 0.01, (1, nb) + xw.shape))
    p, d, dxi, ev = g.mdinit(q, 0)
    sec = timed(lambda: g.verlet(q, p, d, nsteps=nsteps, constrain=-1, event=ev), reps=2)
    report(config="C5 flexible SPC water box (pes WATER_SPC) NVT, 3000 atoms, Zahn, %d bead(s), split path" % nb, ntraj=1,
           steps=nsteps, gpu_bead_steps_per_s=nb * nsteps / sec, gpu_ms_per_step=1e3 * sec / nsteps,
           cpu_bead_steps_per_s_one_core=1.0 / cpu_img)
    g.close()

# ---- N4: the 7-atom CBE-family surfaces (one-lane PesCBE1<K>), recrossing children at config 2's shape ----------
for name in ("ch4oh", "geh4oh", "clnh3", "nh3oh", "h2co"):
    nb, npairs, evol = 16, 512, (100 if name == "h2co" else 500)
    g, o = C.make_pair(name, nb)
    g.set_seed(C.SEED)
    qp = np.array([C.ring_polymer(name, nb, rng, 0.01) for _ in range(8)])
    sec = timed(lambda: g.recross_children(qp, npairs, evol, 0.98), reps=2)
    t0 = time.perf_counter()
    o.recross_children(qp, 0, 2, 50, 0.98, C.SEED, nthreads=1)
    cpu = 2 * 2 * nb * 50 / (time.perf_counter() - t0)
    report(config="N4 calc_rate %s (egrad_%s) 16 beads, %d children x %d steps" % (name, name, 2 * npairs, evol),
           ntraj=2 * npairs, steps=evol, gpu_bead_steps_per_s=2 * npairs * nb * evol / sec,
           cpu_bead_steps_per_s_one_core=cpu)
    g.close()

if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
