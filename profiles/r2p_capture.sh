# round 2, capture P (1 GPU): switching-function derivatives read from the exchange block at their point of use, loop state
# of the child loop in registers -- whole GPU suite, smoke, every bench configuration, ncu of the headline kernel
set -x
O=gpurun_out/r2p
mkdir -p $O
python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -c 1"
cap() {  # name, kernel regex, skip, command...
  n=$1; k=$2; sk=$3; shift 3
  timeout 400 $NCU -k regex:$k --launch-skip $sk -o $O/$n -f "$@" > $O/$n.log 2>&1
  python profiles/ncu_summary.py $O/$n.ncu-rep 40 > $O/$n.txt 2>&1
}
cap recross_ch4h_nb16_1000 recross_kernel 1 python profiles/prof_recross.py 1000 512
python profiles/ncu_traffic.py $O/recross_ch4h_nb16_1000.ncu-rep 1000 512 > $O/traffic.log 2>&1; cp profiles/traffic_recross.json $O/
rm -f $O/recross_ch4h_nb16_1000.ncu-rep
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
for c in c1 c3 c4 c5; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > $O/bench_$c.json 2> $O/bench_$c.err
done
python profiles/bench_egrad.py $O/bench_egrad.json > $O/bench_egrad.log 2>&1
python profiles/bench_umbrella_step.py $O/umbrella_step_ch4h.json > $O/umbrella_step_ch4h.log 2>&1
python profiles/umbrella_multi_gpu.py $O/umbrella_n1.json > $O/umbrella_n1.log 2>&1
python profiles/bench_configs.py $O/bench_configs.json > $O/bench_configs.log 2>&1
cap verlet_ch4h_umbrella verlet_kernel 1 python profiles/prof_umbrella.py; rm -f $O/verlet_ch4h_umbrella.ncu-rep
ls -la $O
