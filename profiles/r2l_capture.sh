# round 2, capture L (1 GPU): the rest of the GPU suite without -x (capture K stopped at the first failure), A/B of the
# shared-memory lane gather in the CBE surface
set -x
O=gpurun_out/r2l
mkdir -p $O
python -m pytest tests/test_gpu_rate.py tests/test_gpu_recross.py tests/test_gpu_round2.py tests/test_gpu_split.py tests/test_gpu_verlet.py tests/test_gpu_water.py tests/test_gpu_xi_mech.py -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
for v in "" _smem; do
  CRCL_LIB_PATH=$PWD/caracal_b200/libcaracal_gpu$v.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench$v.json 2> $O/bench$v.err
done
CRCL_LIB_PATH=$PWD/caracal_b200/libcaracal_gpu_smem.so python -m pytest tests/test_gpu_recross.py tests/test_gpu_verlet.py -q -m gpu -k ch4h > $O/pytest_smem.log 2>&1
ls -la $O
