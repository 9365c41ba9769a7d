"""Generates tests/golden/*.json from the CPU oracle.

The reference cannot run in this container (no Fortran compiler) and ships no
expected outputs (F5), so these fixtures freeze the ORACLE's outputs on seeded inputs: they pin
the oracle against regressions and give the GPU parity tests fixed vectors that need no oracle
build on the GPU box.  Regenerate with:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests import common as C  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(C.SEED)
    pes = {}
    for name in ("h3", "oh3", "ch4h"):
        q = np.concatenate([C.SYSTEMS[name]["ts"]()[None], C.ts_cloud(name, 15, 0.15, rng)])
        V, g, _ = O.egrad(name, q)
        pes[name] = dict(q=q.tolist(), V=V.tolist(), g=g.tolist())
    # surfaces added later draw from their own stream so that the records above stay as first generated
    for k, name in enumerate(("ch4oh", "geh4oh")):
        r2 = np.random.default_rng(C.SEED + 100 + k)
        q = np.concatenate([C.SYSTEMS[name]["ts"]()[None], C.ts_cloud(name, 15, 0.15, r2)])
        V, g, _ = O.egrad(name, q)
        pes[name] = dict(q=q.tolist(), V=V.tolist(), g=g.tolist())
    for k, name in enumerate(("clnh3", "nh3oh", "h2co")):
        r2 = np.random.default_rng(C.SEED + 200 + k)
        q = np.concatenate([C.SYSTEMS[name]["ts"]()[None], C.ts_cloud(name, 15, 0.15, r2)])
        V, g, _ = O.egrad(name, q)
        pes[name] = dict(q=q.tolist(), V=V.tolist(), g=g.tolist())
    with open(os.path.join(HERE, "pes_golden.json"), "w") as f:
        json.dump(pes, f)
    traj = []
    cases = [
        dict(name="h3", nbeads=16, constrain=-1, thermostat=1, andersen_step=70, bias_mode=0, k_force=0.0, xi_ideal=0.0),
        dict(name="h3", nbeads=8, constrain=0, thermostat=1, andersen_step=80, bias_mode=2, k_force=0.05 * 300, xi_ideal=0.9),
        dict(name="h3", nbeads=8, constrain=1, thermostat=1, andersen_step=31, bias_mode=2, k_force=0.05 * 300, xi_ideal=0.98),
        dict(name="ch4h", nbeads=16, constrain=2, thermostat=0, andersen_step=0, bias_mode=2, k_force=0.0, xi_ideal=0.97),
        dict(name="oh3", nbeads=4, constrain=0, thermostat=2, andersen_step=0, bias_mode=2, k_force=0.05 * 300, xi_ideal=0.8, nose_q=100.0),
    ]
    for t, c in enumerate(cases):
        s = O.System(c["name"], c["nbeads"], C.masses(c["name"]), C.beta_calc_rate(300.0), C.dt_au(0.1))
        s.set_mechanism(C.mechanism(c["name"]))
        q0 = C.ring_polymer(c["name"], c["nbeads"], rng, 0.03)
        s.q[:] = q0
        s.set_rng(C.SEED, t)
        s.set_thermostat(c["thermostat"], c["andersen_step"], 300.0, c.get("nose_q", 0.0))
        s.set_kforce(c["k_force"])
        s.mdinit(c["xi_ideal"], c["bias_mode"])
        for i in range(1, 101):
            s.verlet(i, c["xi_ideal"], c["constrain"])
        rec = dict(c, traj=t, nsteps=100, q0=q0.tolist(), q=s.q.tolist(), p=s.p.tolist())
        traj.append(rec)
    with open(os.path.join(HERE, "traj_golden.json"), "w") as f:
        json.dump(traj, f)
    print("wrote golden fixtures")


if __name__ == "__main__":
    main()
