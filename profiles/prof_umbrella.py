"""Small driver for ncu: the umbrella phase of config 2 (CH4+H, 16 beads, 111 windows x 10 trajectories,
constrain = 0: bias + hams force + Andersen + transrot) with short equilibration / sampling legs.
  ncu --set full --clock-control none --import-source on -k regex:verlet_kernel -c 1 \
      -o gpurun_out/prof_umbrella python profiles/prof_umbrella.py [equi] [sample] [traj_per_window]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import caracal_b200  # noqa: E402,F401
from tests import common as C  # noqa: E402

equi = int(sys.argv[1]) if len(sys.argv) > 1 else 50
samp = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ntw = int(sys.argv[3]) if len(sys.argv) > 3 else 10
name, nb = "ch4h", 16
g, _ = C.make_pair(name, nb)
g.set_seed(C.SEED)
g.set_thermostat(1, 70, 300.0)
rng = np.random.default_rng(1)
xi0 = np.linspace(-0.05, 1.05, 111)
kf = np.full(111, 0.05 * 300.0)
q0 = np.array([C.ring_polymer(name, nb, rng, 0.01) for _ in range(111)])
for it in range(2):
    t0 = time.perf_counter()
    avg, var, st = g.umbrella_windows(q0, xi0, kf, ntw, equi, samp)
    sec = time.perf_counter() - t0
print("bead-steps/s %.4g  status0 %.3f" % (111 * ntw * nb * (equi + samp) / sec, (st == 0).mean()))
