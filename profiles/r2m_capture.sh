# round 2, capture M (1 GPU): fm::exp with the library's 0 / inf at the ends of the range (fixes the rate pipeline at
# R = 30 a0), two-phase shared-memory lane gather in the CBE surface (A/B against the shuffles), branch-free divisions in
# transrot / invert3 -- whole GPU suite, headline bench + ncu, configs 1 and 3, biased-step latency
set -x
O=gpurun_out/r2m
mkdir -p $O
python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
CRCL_LIB_PATH=$PWD/caracal_b200/libcaracal_gpu_shfl.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_shfl.json 2> $O/bench_shfl.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n1_again.json 2> $O/bench_n1_again.err
for c in c1 c3; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > $O/bench_$c.json 2> $O/bench_$c.err
done
python profiles/bench_egrad.py $O/bench_egrad.json > $O/bench_egrad.log 2>&1
python profiles/bench_umbrella_step.py $O/umbrella_step_ch4h.json > $O/umbrella_step_ch4h.log 2>&1
python profiles/umbrella_multi_gpu.py $O/umbrella_n1.json > $O/umbrella_n1.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -c 1"
cap() {  # name, kernel regex, skip, command...
  n=$1; k=$2; sk=$3; shift 3
  timeout 400 $NCU -k regex:$k --launch-skip $sk -o $O/$n -f "$@" > $O/$n.log 2>&1
  python profiles/ncu_summary.py $O/$n.ncu-rep 50 > $O/$n.txt 2>&1
}
cap recross_ch4h_nb16_1000 recross_kernel 1 python profiles/prof_recross.py 1000 512
python profiles/ncu_traffic.py $O/recross_ch4h_nb16_1000.ncu-rep 1000 512 > $O/traffic.log 2>&1; cp profiles/traffic_recross.json $O/
rm -f $O/recross_ch4h_nb16_1000.ncu-rep
cap verlet_h3_nb16 verlet_kernel 1 python profiles/prof_h3.py 16384 50; rm -f $O/verlet_h3_nb16.ncu-rep
ls -la $O
