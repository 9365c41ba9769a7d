# round 2, capture T (1 GPU) at HEAD: whole GPU suite with the own SPME FFT (cuFFT gone) and the NH3 + Cl / NH3 + OH surfaces,
# memcheck of the FFT path, PES-only bench of every surface, default bench line, ncu of the two new PES kernels and the FFT sweep
set -x
O=gpurun_out/r2t
mkdir -p $O
python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ewald.py -q -m gpu > $O/sanitizer_ewald.log 2>&1
python profiles/bench_egrad.py $O/bench_egrad.json > $O/bench_egrad.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
for p in clnh3 nh3oh; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:egrad_kernel -c 1 -f -o $O/prof_egrad_$p python profiles/prof_egrad.py $p > $O/prof_egrad_$p.log 2>&1
  python profiles/ncu_summary.py $O/prof_egrad_$p.ncu-rep 30 > $O/egrad_${p}_summary.txt 2>&1
  rm -f $O/prof_egrad_$p.ncu-rep
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_ewald.csv python -m pytest tests/test_gpu_ewald.py -q -m gpu > $O/launches_ewald.log 2>&1
python profiles/bench_configs.py $O/bench_configs.json > $O/bench_configs.log 2>&1
ls -la $O
