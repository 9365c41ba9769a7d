# round 2, capture AC (1 GPU): compute-sanitizer racecheck + synccheck on the tensor-core form of the free ring-polymer step
# (shared-memory state read as B fragments by both warps of a trajectory, written back after the barrier) at 8, 16 and 32 beads
set -x
O=gpurun_out/r2ac
mkdir -p $O
cat > /tmp/san_dmma.py <<'PY'
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
PY
cp /tmp/san_dmma.py $O/san_dmma.py
for tool in racecheck synccheck memcheck; do
  PYTHONPATH=. timeout 900 compute-sanitizer --tool $tool python /tmp/san_dmma.py > $O/sanitizer_$tool.log 2>&1
done
tail -3 $O/sanitizer_*.log
