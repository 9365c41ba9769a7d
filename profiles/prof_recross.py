"""Small driver for ncu: CH4+H 16-bead recrossing children, BASELINE batch (512 pairs) with a
short child_evol so that an `ncu --set full` replay stays cheap.  Usage (on the GPU box):
  ncu --set full --clock-control none --import-source on -k regex:recross_kernel -c 1 \
      -o gpurun_out/prof_recross python profiles/prof_recross.py [steps] [pairs]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import caracal_b200  # noqa: E402
from tests import common as C  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 512
name = sys.argv[3] if len(sys.argv) > 3 else "ch4h"
nb = int(sys.argv[4]) if len(sys.argv) > 4 else 16
g, _ = C.make_pair(name, nb)
g.set_seed(C.SEED)
rng = np.random.default_rng(0)
qp = np.array([C.ring_polymer(name, nb, rng, 0.01) for _ in range(8)])
for it in range(2):
    num, den, st = g.recross_children(qp, pairs, steps, 0.99, pair0=it * pairs)
print("kappa end", num[-1] / den, "kernel ms", g.kernel_timings())
