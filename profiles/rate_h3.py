"""The reference's shipped rate example examples/calc_rate/h+h2/rate.key at FULL size on one GPU
(111 windows x 10 trajectories x (10k + 20k) steps, 8 beads, 300 K; 10 000 children x 500 steps):
kappa(t), PMF barrier and k(T) next to the values the reference publishes as figures (SURVEY.md
section 6).  Usage (GPU box): python profiles/rate_h3.py [out.json [nbeads [exact|asis [norot|rot [seed]]]]]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import caracal_b200  # noqa: E402
from caracal_b200 import rate as R  # noqa: E402
from tests import common as C  # noqa: E402

name, kelvin = "h3", 300.0
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 8
exact = len(sys.argv) > 3 and sys.argv[3] == "exact"   # true normal-mode transform instead of rfft/irfft as written
seed = int(sys.argv[5]) if len(sys.argv) > 5 else C.SEED   # another seed = another draw of parents and children
norot = len(sys.argv) > 4 and sys.argv[4] == "norot"   # constrain = 3: no removal of net rotation in phases 1-2
m, mech = C.masses(name), C.mechanism(name, dist_inf=16.0 / C.BOHR)   # DIST_INF 16 (Angstrom), calc_rate_read.f90:693
beta, dt = C.beta_calc_rate(kelvin), C.dt_au(0.1)
g, g1 = caracal_b200.RPMD(name, nb, m, beta, dt), caracal_b200.RPMD(name, 1, m, beta, dt)
for h in (g, g1):
    h.set_mechanism(mech)
    h.set_seed(seed)
    if exact:
        h.set_transform(caracal_b200.TRANSFORM_EXACT)
t0 = time.perf_counter()
out = R.calc_rate(g, g1, C.h3_ts(), m, mech, kelvin, beta, npaths=2, umbr_constrain=3 if norot else 0,
                  log=lambda *a: print(*a, flush=True))
wall = time.perf_counter() - t0
res = dict(example="examples/calc_rate/h+h2/rate.key (full size)", nbeads=nb, seed=seed,
           transform="exact" if exact else "reference (rfft/irfft as written)",
           rotation_removed_in_umbrella_phase=not norot, wall_s=wall, phase_s=out["timings"],
           pmf_at={"%.1f" % x: float((out["pmf"][int(np.argmin(np.abs(out["bin_coord"][:-1] - x)))]
                                       - out["pmf"][out["minlocate"]]) * R.HARTREE_KJ) for x in (0.2, 0.4, 0.6, 0.8, 0.9)}, xi_barrier=float(out["xi_barrier"]),
           delta_w_kj=float(out["delta_w_kj"]), kappa=float(out["kappa"]), kappa_t_every_50=out["kappa_t"][49::50].tolist(),
           k_t_cm3_per_molecule_s=float(out["k_t_molec"]), n_rerun=int(out["n_rerun"]),
           published=dict(kappa_50fs=0.767, delta_w_kj=48.5, source="manual/figures/kappa_h3.png, pmf_h3.png"))
print(json.dumps(res))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
