"""Latency / throughput of one biased step on the fused path (CH4+H, 16 beads): constrain 0 (umbrella + removal of net
rotation, as calc_rate.f90 runs its umbrella phase), 3 (the same without transrot), 1 (SHAKE / RATTLE parent), 2 (child),
at a full batch (1110 trajectories = 111 windows x 10) and at the batch one of eight GPUs sees (139).
    python profiles/bench_umbrella_step.py [out.json]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import caracal_b200  # noqa: E402
from caracal_b200 import systems as S  # noqa: E402
from caracal_b200.api import beta_calc_rate, dt_au  # noqa: E402

name, nb = sys.argv[2] if len(sys.argv) > 2 else "ch4h", int(sys.argv[3]) if len(sys.argv) > 3 else 16
g = caracal_b200.RPMD(name, nb, S.masses(name), beta_calc_rate(300.0), dt_au(0.1))
g.set_mechanism(S.mechanism(name))
g.set_seed(1)
rows = []
for ntraj in (1110, 139):
    q0 = np.repeat(S.SYSTEMS[name]["ts"]()[None, None], ntraj, axis=0).repeat(nb, axis=1)
    xi = np.linspace(0.0, 1.05, ntraj)
    kf = np.full(ntraj, 15.0)
    for constrain, thermo in ((0, 1), (3, 1), (1, 1), (2, 0), (-1, 1)):
        g.set_thermostat(thermo, 80, 300.0)
        q = q0.copy()
        p, d, dxi, ev = g.mdinit(q, 2, xi, kf)
        nsteps = 3000
        g.verlet(q, p, d, nsteps=100, constrain=constrain, xi_ideal=xi, k_force=kf, dxi=dxi, event=ev)
        g.kernel_timings()
        g.verlet(q, p, d, nsteps=nsteps, constrain=constrain, xi_ideal=xi, k_force=kf, dxi=dxi, event=ev)
        ms = float(g.kernel_timings()[-1])
        rows.append(dict(pes=name, nbeads=nb, ntraj=ntraj, constrain=constrain, us_per_step=1e3 * ms / nsteps,
                         bead_steps_per_s=ntraj * nb * nsteps / (ms * 1e-3)))
        print(json.dumps(rows[-1]), flush=True)
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
