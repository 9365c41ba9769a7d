"""Timing of the QMDFF kernels at the periodic-box shape (config 5: ~3000 atoms, 8 beads = 8 images
per step) and the CPU oracle beside it.  Run on the GPU box: python profiles/bench_qmdff.py"""
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
from tests.qmdff_synth import make_system  # noqa: E402
from tests.test_gpu_qmdff import handle  # noqa: E402

rows = []
for nmol, nimg in [(385, 8), (385, 64), (125, 64)]:
    T = make_system(nmol=nmol, seed=12, periodic=True, zahn=True)
    g, _ = handle(caracal_b200, T)
    rng = np.random.default_rng(3)
    x = T["xyz"][None] + rng.normal(0, 0.05, (nimg,) + T["xyz"].shape)
    for _ in range(3):
        V, grad, _ = g.egrad(x)
    g.kernel_timings()
    for _ in range(5):
        g.egrad(x)
    ms = float(np.mean(g.kernel_timings()))
    n = T["n"]
    inter_pairs = n * (n - 1) // 2
    t0 = time.perf_counter()
    Q = O.Qmdff(T)
    Q.egrad(x[:1])
    cpu = time.perf_counter() - t0
    rows.append(dict(natoms=n, nimg=nimg, gpu_ms=ms, images_per_s=nimg / (ms * 1e-3), pair_tests_per_s=2 * inter_pairs * nimg / (ms * 1e-3),
                     cpu_oracle_s_per_image=cpu))
    print("natoms %d images %d: GPU %.3f ms (%.0f images/s, %.2e ordered pair tests/s); CPU oracle %.3f s/image"
          % (n, nimg, ms, nimg / (ms * 1e-3), 2 * inter_pairs * nimg / (ms * 1e-3), cpu))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "bench_qmdff.json"), "w"), indent=1)
