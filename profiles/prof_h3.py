"""ncu driver: config-1 shape, H + H2 (egrad_h3) 16 beads NVT Andersen, a batch of replicas.
  ncu --set full --clock-control none --import-source on -k regex:verlet_kernel -c 1 -o gpurun_out/prof_h3 \
      python profiles/prof_h3.py [ntraj] [nsteps] [pes] [nbeads]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import caracal_b200  # noqa: E402
from tests import common as C  # noqa: E402

ntraj = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
name = sys.argv[3] if len(sys.argv) > 3 else "h3"
nb = int(sys.argv[4]) if len(sys.argv) > 4 else 16
g, _ = C.make_pair(name, nb)
g.set_seed(C.SEED)
g.set_thermostat(1, 70, 300.0)
rng = np.random.default_rng(0)
q = np.array([C.ring_polymer(name, nb, rng, 0.02) for _ in range(64)])
q = np.ascontiguousarray(np.resize(q, (ntraj,) + q.shape[1:]))
p, d, dxi, ev = g.mdinit(q, 0)
import time  # noqa: E402
g.verlet(q, p, d, nsteps=nsteps, constrain=-1, event=ev)
t0 = time.perf_counter()
g.verlet(q, p, d, nsteps=nsteps, constrain=-1, event=ev)
sec = time.perf_counter() - t0
print("%s nb %d ntraj %d: %.2f ms per call, %.3e bead-steps/s (wall, incl. copies)" % (name, nb, ntraj, sec * 1e3, ntraj * nb * nsteps / sec))
