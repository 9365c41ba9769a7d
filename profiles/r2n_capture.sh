# round 2, capture N (1 GPU): own late-use values parked in the lane-exchange block, merged half kicks in the child loop,
# fm math in xi_value and the QMDFF pair stage, cofactor inverse of the inertia tensor; A/B: 256-thread CTAs (four
# trajectories in step) for the headline kernel -- whole GPU suite, every bench configuration, ncu of the headline kernel
set -x
O=gpurun_out/r2n
mkdir -p $O
python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
CRCL_LIB_PATH=$PWD/caracal_b200/libcaracal_gpu_ctpb256.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_ctpb256.json 2> $O/bench_ctpb256.err
for c in c1 c3 c4 c5; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > $O/bench_$c.json 2> $O/bench_$c.err
done
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
