# round 2, capture AF (1 GPU): closing validation at HEAD -- whole GPU suite, smoke, default bench line
set -x
O=gpurun_out/r2af
mkdir -p $O
python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
ls -la $O
