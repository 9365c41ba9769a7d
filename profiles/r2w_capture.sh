# round 2, capture W (1 GPU): free ring-polymer step on the FP64 tensor cores for every lane-split surface (CRCL_DMMA_TRANSFORM
# default on) -- whole GPU suite, smoke, bench line of the DMMA build against the FMA build of the headline unit, ncu of the
# headline kernel, the other configurations that run lane-split surfaces (umbrella phase, N4 children)
set -x
O=gpurun_out/r2w
mkdir -p $O
python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
for v in "" _nodmma; do
  CRCL_LIB_PATH=/root/repo/caracal_b200/libcaracal_gpu$v.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n1$v.json 2> $O/bench_n1$v.err
done
NCU="ncu --set full --clock-control none --import-source on -c 1"
cap() {  # name, kernel regex, skip, command...
  n=$1; k=$2; sk=$3; shift 3
  timeout 400 $NCU -k regex:$k --launch-skip $sk -o $O/$n -f "$@" > $O/$n.log 2>&1
  python profiles/ncu_summary.py $O/$n.ncu-rep 40 > $O/$n.txt 2>&1
}
cap recross_ch4h_nb16_1000 recross_kernel 1 python profiles/prof_recross.py 1000 512
python profiles/ncu_traffic.py $O/recross_ch4h_nb16_1000.ncu-rep 1000 512 > $O/traffic.log 2>&1; cp profiles/traffic_recross.json $O/
rm -f $O/recross_ch4h_nb16_1000.ncu-rep
python profiles/bench_umbrella_step.py $O/umbrella_step_ch4h.json > $O/umbrella_step_ch4h.log 2>&1
python profiles/umbrella_multi_gpu.py $O/umbrella_n1.json > $O/umbrella_n1.log 2>&1
python profiles/bench_configs.py $O/bench_configs.json > $O/bench_configs.log 2>&1
ls -la $O
