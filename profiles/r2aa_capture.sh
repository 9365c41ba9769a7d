# round 2, capture AA (1 GPU): the H2CO surface -- whole GPU suite, PES-only bench incl. h2co, N4 children rows
set -x
O=gpurun_out/r2aa
mkdir -p $O
python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
python profiles/bench_egrad.py $O/bench_egrad.json > $O/bench_egrad.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:egrad_kernel -c 1 -f -o $O/prof_egrad_h2co python profiles/prof_egrad.py h2co 16 > $O/prof_egrad_h2co.log 2>&1
python profiles/ncu_summary.py $O/prof_egrad_h2co.ncu-rep 25 > $O/egrad_h2co_summary.txt 2>&1
rm -f $O/prof_egrad_h2co.ncu-rep
python profiles/bench_configs.py $O/bench_configs.json > $O/bench_configs.log 2>&1
ls -la $O
