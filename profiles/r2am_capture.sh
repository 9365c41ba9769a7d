# round 2, capture AM (1 GPU): one-bead trajectories of the lane-split surfaces as eight quads (PesSpreadQ), the batch limit of
# the spread forms behind the C-ABI (crcl_set_spread_max_traj) -- GPU suite, chain link by surface and form
set -x
O=gpurun_out/r2am
mkdir -p $O
python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
for pes in h3 ch4h oh3; do
  for sm in 1024 0; do
    timeout 200 python profiles/prof_chain_h3.py 5000 0 $pes $sm >> $O/chain_time.log 2>&1
  done
done
ls -la $O
