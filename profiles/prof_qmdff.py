"""Small driver for ncu: QMDFF egrad of a periodic Zahn box at the config-5 shape (~3000 atoms,
8 images = 8 beads of one RPMD step) with the H-bond terms on.  Usage (on the GPU box):
  ncu --set full --clock-control none --import-source on -k regex:qm_inter -c 1 \
      -o gpurun_out/prof_qm_inter python profiles/prof_qmdff.py [nmol] [nimg] [hb]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import caracal_b200  # noqa: E402
from tests.qmdff_synth import make_system  # noqa: E402
from tests.test_gpu_qmdff import handle  # noqa: E402

nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 385
nimg = int(sys.argv[2]) if len(sys.argv) > 2 else 8
hb = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
T = make_system(nmol=nmol, seed=12, periodic=True, zahn=True, hb=hb)
g, _ = handle(caracal_b200, T)
x = T["xyz"][None] + np.random.default_rng(3).normal(0, 0.05, (nimg,) + T["xyz"].shape)
for _ in range(2):
    V, grad, _ = g.egrad(x)
print("natoms", T["n"], "V[0]", V[0], "kernel ms", g.kernel_timings())
