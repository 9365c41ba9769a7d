"""ncu driver for the PES seam: one egrad_kernel<PES> launch on 2^18 images around the TS.
  ncu --set full --clock-control none --import-source on -k regex:egrad_kernel -c 1 -o gpurun_out/prof_egrad_<pes> \
      python profiles/prof_egrad.py <pes> [log2 images]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import caracal_b200  # noqa: E402,F401
from tests import common as C  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ch4h"
nimg = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 18)
g, _ = C.make_pair(name, 1)
rng = np.random.default_rng(0)
q = C.ts_cloud(name, 4096, 0.15, rng)
q = np.ascontiguousarray(np.resize(q, (nimg,) + q.shape[1:]))
for _ in range(2):
    V, grad, info = g.egrad(q)
print(name, "images", nimg, "V[0]", V[0], "kernel ms", g.kernel_timings())
