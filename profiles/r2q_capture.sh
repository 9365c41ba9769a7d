# round 2, capture Q (1 GPU): 256-thread CTAs in the CH4 + H recross unit; term-major indexing of the bonded / nci QMDFF terms
# and fm math there (A/B through CRCL_QM_TERM_MAJOR=0), launch list of a config-4 step
set -x
O=gpurun_out/r2q
mkdir -p $O
python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
for c in c4 c5; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > $O/bench_$c.json 2> $O/bench_$c.err
  CRCL_QM_TERM_MAJOR=0 timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_${c}_imgmajor.json 2> $O/bench_${c}_imgmajor.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_c4.csv python profiles/prof_c4.py > $O/prof_c4.log 2>&1
CRCL_QM_TERM_MAJOR=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_c4_imgmajor.csv python profiles/prof_c4.py > $O/prof_c4_imgmajor.log 2>&1
ls -la $O
