# round 2, capture I (1 GPU): biased-step latency with the barrier-free transrot
set -x
O=gpurun_out/r2i
mkdir -p $O
python -m pytest tests/test_gpu_verlet.py tests/test_gpu_rate.py tests/test_gpu_recross.py tests/test_gpu_xi_mech.py -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
python profiles/bench_umbrella_step.py $O/umbrella_step_ch4h.json > $O/umbrella_step_ch4h.log 2>&1
python profiles/bench_umbrella_step.py $O/umbrella_step_h3.json h3 8 > $O/umbrella_step_h3.log 2>&1
python profiles/umbrella_multi_gpu.py $O/umbrella_n1.json > $O/umbrella_n1.log 2>&1
timeout 900 python bench.py --config c1 --steps 10 --warmup 3 > $O/bench_c1.json 2> $O/bench_c1.err
ls -la $O
