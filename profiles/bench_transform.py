"""The bead transform (free ring-polymer step) from shared memory: FMA form (the loop of the fused trajectory kernels) against
an FP64 tensor-core form (mma.sync.m8n8k4.f64), csrc/transform_bench.cuh.  north_star: "FP64 DMMA tensor cores are used for
the bead transform only if ncu shows they beat the FFMA path".  Run on the GPU box:
    python profiles/bench_transform.py [out.json]
    ncu --set full --clock-control none -k regex:transform_bench -c 4 -o ... python profiles/bench_transform.py - ncu"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import caracal_b200  # noqa: E402
from caracal_b200.api import beta_calc_rate, dt_au  # noqa: E402
from caracal_b200 import systems as S  # noqa: E402

under_ncu = len(sys.argv) > 2 and sys.argv[2] == "ncu"
rows = []
g = caracal_b200.RPMD("ch4h", 16, S.masses("ch4h"), beta_calc_rate(300.0), dt_au(0.1))
peak, dmma = (0.0, 0.0) if under_ncu else (g.measure_fp64_tflops(16384), g.measure_dmma_tflops(16384))
for nb, ntraj, reps in ((16, 148 * 4 * 4 * 8, 200), (64, 148 * 4 * 8, 50)):
    if under_ncu:
        ntraj, reps = ntraj // 8, 5      # ncu: the 3-step comparison launches are the ones captured (-c 4 takes both pairs)
    r = g.bench_transform(nb, ntraj, reps)
    row = dict(nbeads=nb, ntraj=ntraj, steps_per_launch=reps, **r)
    row["tflops_dfma"] = r["flops"] / (r["ms_dfma"] * 1e-3) / 1e12
    row["tflops_dmma"] = r["flops"] / (r["ms_dmma"] * 1e-3) / 1e12
    row["dmma_over_dfma"] = r["ms_dfma"] / r["ms_dmma"]
    row["dfma_peak_tflops"], row["dmma_peak_tflops"] = peak, dmma
    rows.append(row)
    print(json.dumps(row), flush=True)
if len(sys.argv) > 1 and sys.argv[1] != "-":
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
