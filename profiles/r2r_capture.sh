# round 2, capture R (1 GPU): DG-EVB mixing kernel with fm math and the Wilson term reduced straight into the output
# (no shared-memory CAS atomics) -- whole GPU suite, config 4 bench + launch list, the H + H2 rate example at full size
set -x
O=gpurun_out/r2r
mkdir -p $O
python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 900 python bench.py --config c4 --steps 10 --warmup 3 > $O/bench_c4.json 2> $O/bench_c4.err
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_c4.csv python profiles/prof_c4.py > $O/prof_c4.log 2>&1
python profiles/bench_configs.py $O/bench_configs.json > $O/bench_configs.log 2>&1
python profiles/rate_h3.py $O/rate_h3_nb8_exact_norot.json 8 exact norot > $O/rate_h3_exact.log 2>&1
python profiles/rate_h3.py $O/rate_h3_nb8_asis.json 8 > $O/rate_h3_asis.log 2>&1
ls -la $O
