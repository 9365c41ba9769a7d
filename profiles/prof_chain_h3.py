"""Small driver for ncu: one link of the start-structure chain of calc_rate (calc_rate.f90:651-1148) -- ONE
one-bead H + H2 trajectory, mdinit(bias) + biased steps at a window (rate.py generate_start_structures).
  ncu --set full --clock-control none --import-source on -k regex:verlet_kernel -c 1 \
      -o gpurun_out/prof_chain_h3 python profiles/prof_chain_h3.py [steps] [constrain] [pes] [spread_max]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import caracal_b200  # noqa: E402
from tests import common as C  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
constrain = int(sys.argv[2]) if len(sys.argv) > 2 else 0
name = sys.argv[3] if len(sys.argv) > 3 else "h3"
kelvin = 300.0
m, mech = C.masses(name), C.mechanism(name, dist_inf=16.0 / C.BOHR)
beta, dt = C.beta_calc_rate(kelvin), C.dt_au(0.1)
g1 = caracal_b200.RPMD(name, 1, m, beta, dt)
g1.set_mechanism(mech)
g1.set_seed(C.SEED)
g1.set_thermostat(1, 70, kelvin)
if len(sys.argv) > 4:
    g1.set_spread_max_beads(int(sys.argv[4]))   # 0: the one-thread-per-trajectory form
ts = C.h3_ts() if name == "h3" else C.ring_polymer(name, 1, np.random.default_rng(1), 0.0)
q = np.array(ts, dtype=np.float64).reshape(1, 1, len(m), 3).copy()
tid = np.array([7], dtype=np.uint32)
xi0, kf = np.array([0.95]), np.array([0.05 * kelvin])
for it in range(3):
    t0 = time.perf_counter()
    p, d, dxi, ev = g1.mdinit(q, 2, xi_ideal=xi0, k_force=kf, traj_id=tid)
    ep, xr, st = g1.verlet(q, p, d, nsteps=steps, constrain=constrain, xi_ideal=xi0, k_force=kf, dxi=dxi,
                           traj_id=tid, event=ev)
    sec = time.perf_counter() - t0
    print("%s spread_max %s pass %d: %.3f us per step (wall, mdinit + %d steps)  xi %.4f  status %d" % (name, sys.argv[4] if len(sys.argv) > 4 else "default", it, 1e6 * sec / steps, steps, xr[0], st[0]))
