# round 2, capture V (1 GPU): the free ring-polymer step of the headline kernel on the FP64 tensor cores (CRCL_DMMA_TRANSFORM in
# traj_ch4h_recross.cu) -- parity of the recrossing work unit, then the bench line of the DMMA build against the FMA build
set -x
O=gpurun_out/r2v
mkdir -p $O
python -m pytest tests/test_gpu_recross.py tests/test_gpu_verlet.py -q -m gpu -x > $O/pytest_recross.log 2>&1; echo "exit $?" >> $O/pytest_recross.log
for v in "" _nodmma "" _nodmma; do
  CRCL_LIB_PATH=/root/repo/caracal_b200/libcaracal_gpu$v.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench$v.json 2> $O/bench$v.err
  python - <<PY
import json
d = json.loads(open("$O/bench$v.json").read().strip().splitlines()[-1])
print("variant '$v' ms/step %.3f value %.4e" % (d["ms_per_step"], d["value"]))
PY
done
ls -la $O
