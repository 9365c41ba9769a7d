"""ncu driver: one RPMD verlet call at the config-4 shape (DG-EVB, 32 beads, split path).
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c4.csv \
      python profiles/prof_c4.py [ntraj] [nsteps] [20]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import caracal_b200  # noqa: E402
from tests import common as C  # noqa: E402
from tests.qmdff_synth import HEXANE, make_dgevb  # noqa: E402

ntraj = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
big = len(sys.argv) > 3 and sys.argv[3] == "20"   # 20 atoms, nat6 = 12
T1, T2, E = make_dgevb(seed=5, mode=3, npoints=7, template=HEXANE if big else None)
nb = 32
mass = np.array([C.atomic_mass_au({1: "H", 6: "C", 8: "O"}[int(z)]) for z in T1["at"]])
g = caracal_b200.RPMD(caracal_b200.PES_DGEVB, nb, mass, C.beta_calc_rate(300.0), C.dt_au(0.2))
g.set_qmdff(T1)
g.set_qmdff(T2, second=True)
g.set_dgevb(E)
g.set_seed(C.SEED)
g.set_thermostat(1, 10, 300.0)
rng = np.random.default_rng(0)
q = np.ascontiguousarray(T1["xyz"][None, None] + rng.normal(0, 0.01, (ntraj, nb) + T1["xyz"].shape))
p, d, dxi, ev = g.mdinit(q, 0)
g.verlet(q, p, d, nsteps=nsteps, constrain=-1, event=ev)
print("ok", g.launch_count())
