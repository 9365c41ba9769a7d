# round 2, capture E (1 GPU): whole GPU suite with the CUDA-graph replay live, graph on/off A/B at configs 4 and 5
set -x
O=gpurun_out/r2e
mkdir -p $O
python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
python - > $O/graph_ab.log 2>&1 <<'PY'
import json, time, sys, os
import numpy as np
sys.path.insert(0, os.getcwd())
import caracal_b200
from caracal_b200.api import atomic_mass_au, beta_calc_rate, beta_dynamic, dt_au
from caracal_b200.qmdff_synth import HEXANE, make_dgevb, make_system
rows = []
def run(label, g, q0, nsteps):
    g.set_seed(1); g.set_thermostat(1, 10, 300.0)
    for on in (0, 1, 0, 1):
        g.set_graph(on)
        q = q0.copy()
        p, d, dxi, ev = g.mdinit(q, 0)
        g.verlet(q, p, d, nsteps=10, constrain=-1, event=ev)
        t0 = time.perf_counter()
        g.verlet(q, p, d, nsteps=nsteps, constrain=-1, event=ev)
        sec = time.perf_counter() - t0
        rows.append(dict(config=label, graph=on, us_per_step=1e6 * sec / nsteps))
        print(json.dumps(rows[-1]), flush=True)
rng = np.random.default_rng(0)
T1, T2, E = make_dgevb(seed=5, mode=3, npoints=7, template=HEXANE)
mass = np.array([atomic_mass_au({1: "H", 6: "C", 8: "O"}[int(z)]) for z in T1["at"]])
g = caracal_b200.RPMD(caracal_b200.PES_DGEVB, 32, mass, beta_calc_rate(300.0), dt_au(0.2))
g.set_qmdff(T1); g.set_qmdff(T2, second=True); g.set_dgevb(E)
run("c4: 256 traj x 32 beads x 20 atoms", g, T1["xyz"][None, None] + rng.normal(0, 0.01, (256, 32) + T1["xyz"].shape), 200)
g.close()
T = make_system(nmol=385, seed=12, periodic=True, zahn=True, hb=True)
mass = np.array([atomic_mass_au({1: "H", 6: "C", 8: "O", 17: "CL"}[int(z)]) for z in T["at"]])
for nb in (8, 1):
    g = caracal_b200.RPMD(caracal_b200.PES_QMDFF, nb, mass, beta_dynamic(300.0), dt_au(0.5))
    g.set_qmdff(T)
    run("c5: 1 traj x %d beads x 3030 atoms" % nb, g, T["xyz"][None, None] + rng.normal(0, 0.01, (1, nb) + T["xyz"].shape), 300)
    g.close()
json.dump(rows, open("gpurun_out/r2e/graph_ab.json", "w"), indent=1)
PY
for c in c4 c5; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > $O/bench_$c.json 2> $O/bench_$c.err
done
ls -la $O
