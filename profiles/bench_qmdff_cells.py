"""A/B of the inter-molecular sweep at the config-5 shape (385 molecules = 3080 atoms, periodic Zahn box, 10 A cut-offs):
O(N^2) sweep (CRCL_QM_CELLS=0) against the cell sweep with M = 2 and M = 3; whole-egrad time per call from the library's
CUDA events, 8 images (one 8-bead RPMD step) and 64.  python profiles/bench_qmdff_cells.py [out.json]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import caracal_b200  # noqa: E402
from tests.qmdff_synth import make_system  # noqa: E402
from tests.test_gpu_qmdff import handle  # noqa: E402

rows = []
for nmol, nimg, hb in [(385, 8, False), (385, 8, True), (385, 64, True), (1000, 8, True)]:
    T = make_system(nmol=nmol, seed=12, periodic=True, zahn=True, hb=hb)
    x = T["xyz"][None] + np.random.default_rng(3).normal(0, 0.05, (nimg,) + T["xyz"].shape)
    for label, env in (("n2", {"CRCL_QM_CELLS": "0"}), ("cells_m2", {"CRCL_QM_CELLS": "1", "CRCL_QM_CELL_M": "2"}),
                       ("cells_m3", {"CRCL_QM_CELLS": "1", "CRCL_QM_CELL_M": "3"})):
        os.environ.update(env)
        g, _ = handle(caracal_b200, T)
        for _ in range(3):
            V, grad, _ = g.egrad(x)
        g.kernel_timings()
        for _ in range(10):
            g.egrad(x)
        ms = g.kernel_timings()
        row = dict(natoms=int(T["n"]), nimg=nimg, hb=hb, sweep=label, egrad_ms_mean=float(np.mean(ms)), egrad_ms_min=float(np.min(ms)),
                   V0=float(V[0]))
        rows.append(row)
        print(json.dumps(row), flush=True)
        g.close()
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
