import numpy as np
from tests import common as C
for nb, ntraj in ((16, 6), (32, 3), (8, 5)):
    g, _ = C.make_pair("ch4h", nb)
    g.set_seed(C.SEED)
    rng = np.random.default_rng(nb)
    q = np.array([C.ring_polymer("ch4h", nb, rng, 0.02) for _ in range(ntraj)])
    p, d, dxi, ev = g.mdinit(q, 2, 0.97, 0.0)
    g.verlet(q, p, d, nsteps=6, constrain=2, xi_ideal=0.97, k_force=0.0, dxi=dxi, event=ev)
    num, den, st = g.recross_children(q, 4, 5, 0.97)
    print(nb, "ok", float(den), int(st.max()))
