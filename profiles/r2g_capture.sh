# round 2, capture G (1 GPU): biased-step latency after the transrot rework (batched reductions, register-resident invert3,
# epot sum only on the last step); whole GPU suite; configs 1, 4, 5
set -x
O=gpurun_out/r2g
mkdir -p $O
python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
python profiles/bench_umbrella_step.py $O/umbrella_step_ch4h.json > $O/umbrella_step_ch4h.log 2>&1
python profiles/bench_umbrella_step.py $O/umbrella_step_h3.json h3 8 > $O/umbrella_step_h3.log 2>&1
python profiles/umbrella_multi_gpu.py $O/umbrella_n1.json > $O/umbrella_n1.log 2>&1
for c in c1 c4 c5; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > $O/bench_$c.json 2> $O/bench_$c.err
done
NCU="ncu --set full --clock-control none --import-source on -c 1"
timeout 300 $NCU -k regex:verlet_kernel --launch-skip 1 -o $O/verlet_ch4h_umbrella -f python profiles/prof_umbrella.py > $O/verlet_ch4h_umbrella.log 2>&1
python profiles/ncu_summary.py $O/verlet_ch4h_umbrella.ncu-rep 40 > $O/verlet_ch4h_umbrella.txt 2>&1; rm -f $O/verlet_ch4h_umbrella.ncu-rep
ls -la $O
