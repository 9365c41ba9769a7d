# round 2, capture K (1 GPU): branch-free FP64 math in every analytic surface (H3 with one reciprocal per distance and
# the triplet branches as selects), child step specialised at compile time, CH4 + CN -- whole GPU suite, headline
# bench + ncu, PES-only bench, configs 1 and 3
set -x
O=gpurun_out/r2k
mkdir -p $O
python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
python profiles/bench_egrad.py $O/bench_egrad.json > $O/bench_egrad.log 2>&1
for c in c1 c3; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > $O/bench_$c.json 2> $O/bench_$c.err
done
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
for p in h3 ch4h; do cap egrad_$p egrad_kernel 0 python profiles/prof_egrad.py $p; rm -f $O/egrad_$p.ncu-rep; done
python profiles/bench_umbrella_step.py $O/umbrella_step_ch4h.json > $O/umbrella_step_ch4h.log 2>&1
ls -la $O
