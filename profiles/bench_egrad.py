"""PES-only kernels (egrad_kernel<PES>, one thread per image): evaluations/s and achieved FP64 FLOP/s by
the reference's own operation count (oracle/flop_census.json), against the DFMA peak measured in the
same process; plus the FP64 tensor-core (DMMA m8n8k4) peak next to it -- the evidence for keeping the
bead transform on the DFMA pipe.  Run on the GPU box:  python profiles/bench_egrad.py [out.json]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import caracal_b200  # noqa: E402
from tests import common as C  # noqa: E402

census = json.load(open(os.path.join(ROOT, "oracle", "flop_census.json")))
rows = []
rng = np.random.default_rng(0)
peak = dmma = None
for name in ("h3", "oh3", "ch4h", "brh2", "o3", "ch4oh", "geh4oh", "ch4cn", "clnh3", "nh3oh", "h2co"):
    g, _ = C.make_pair(name, 1)
    if peak is None:
        peak, dmma = g.measure_fp64_tflops(16384), g.measure_dmma_tflops(16384)
    nimg = 1 << 20
    q = C.ts_cloud(name, 4096, 0.15, rng)
    q = np.ascontiguousarray(np.resize(q, (nimg,) + q.shape[1:]))
    g.kernel_timings()
    for _ in range(4):
        g.egrad(q)
    ms = float(np.min(g.kernel_timings()[1:]))
    fl = census[name]["flops"]
    natoms = q.shape[1]
    row = dict(pes=name, images=nimg, kernel_ms=ms, evaluations_per_s=nimg / (ms * 1e-3), flops_per_evaluation=fl,
               achieved_tflops=nimg * fl / (ms * 1e-3) / 1e12, dfma_peak_tflops=peak, frac=nimg * fl / (ms * 1e-3) / 1e12 / peak,
               algorithmic_bytes_per_evaluation=48 * natoms + 8, hbm_gbs=nimg * (48 * natoms + 8) / (ms * 1e-3) / 1e9)
    rows.append(row)
    print(json.dumps(row), flush=True)
    g.close()
print(json.dumps(dict(dfma_peak_tflops=peak, dmma_m8n8k4_peak_tflops=dmma)))
if len(sys.argv) > 1:
    json.dump(dict(kernels=rows, dfma_peak_tflops=peak, dmma_m8n8k4_peak_tflops=dmma), open(sys.argv[1], "w"), indent=1)
