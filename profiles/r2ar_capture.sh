# round 2, capture AR (1 GPU): sanity of the library rebuilt after a comment-only edit of traj_inst.cuh -- smoke and the
# few-bead trajectory cases
set -x
O=gpurun_out/r2ar
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log
timeout 80 python -m pytest tests/test_gpu_verlet.py -q -x -k "nb1 or nb2 or nb4 or packed or nb8" > $O/pytest_few_bead.log 2>&1; echo "pytest exit $?" >> $O/pytest_few_bead.log
