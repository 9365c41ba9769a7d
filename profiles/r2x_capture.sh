# round 2, capture X (1 GPU): PES seam kernel with the CTA tile staged through shared memory -- whole GPU suite, PES-only bench,
# ncu of the lightest and of one mid-weight surface; headline capture once more for the per-pipe counters (DMMA vs DFMA)
set -x
O=gpurun_out/r2x
mkdir -p $O
python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
python profiles/bench_egrad.py $O/bench_egrad.json > $O/bench_egrad.log 2>&1
for p in oh3 clnh3; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:egrad_kernel -c 1 -f -o $O/prof_egrad_$p python profiles/prof_egrad.py $p 20 > $O/prof_egrad_$p.log 2>&1
  python profiles/ncu_summary.py $O/prof_egrad_$p.ncu-rep 20 > $O/egrad_${p}_summary.txt 2>&1
  rm -f $O/prof_egrad_$p.ncu-rep
done
timeout 400 ncu --set full --clock-control none --import-source on -c 1 -k regex:recross_kernel --launch-skip 1 -o $O/recross -f python profiles/prof_recross.py 1000 512 > $O/recross.log 2>&1
python profiles/ncu_traffic.py $O/recross.ncu-rep 1000 512 > $O/traffic.log 2>&1; cp profiles/traffic_recross.json $O/
rm -f $O/recross.ncu-rep
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
ls -la $O
