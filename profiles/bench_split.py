"""HBM roofline of the split path's propagation kernel (half kick + free ring polymer + centroid)
at the periodic-box shape (3000 atoms x 8 beads) and two other shapes.  Run on the GPU box:
  python profiles/bench_split.py
Algorithmic bytes: 120 B per (trajectory, bead, atom) (read q,p,g; write q,p).  Peak: the measured
copy bandwidth in MEASURED_PEAKS.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import caracal_b200  # noqa: E402
from tests import common as C  # noqa: E402

peak = 6538.9
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
rows = []
for natoms, nb, ntraj in [(3000, 8, 2048), (3000, 1, 8192), (1125, 8, 4096), (20, 32, 65536), (3000, 16, 1024)]:
    mass = np.full(natoms, C.atomic_mass_au("O"))
    g = caracal_b200.RPMD(caracal_b200.PES_HOSTCB, nb, mass, C.beta_calc_rate(300.0), C.dt_au(0.5))
    ms, best, gbs = g.bench_propagate(ntraj, reps=10)
    nbytes = 120.0 * ntraj * nb * natoms
    rows.append(dict(natoms=natoms, nbeads=nb, ntraj=ntraj, bytes=nbytes, ms_mean=ms, ms_best=best, gbs=gbs,
                     frac_of_measured_peak=gbs / peak))
    print("natoms %5d nbeads %3d ntraj %6d  %.3f GB  %.3f ms  %.0f GB/s  %.2f of measured %.0f GB/s"
          % (natoms, nb, ntraj, nbytes / 1e9, ms, gbs, gbs / peak, peak))
    g.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "bench_split.json"), "w"), indent=1)
