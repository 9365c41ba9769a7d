# round 2, capture A (1 GPU): GPU test suite at HEAD, smoke, ncu summaries of every PES-seam kernel
# and "before" captures of the kernels this round changes; summaries are made on the box (the .ncu-rep files with
# source are ~20 MB each and gpurun_out/ is capped at 64 MiB)
set -x
O=gpurun_out/r2a
mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -c 1"
cap() {  # name, kernel regex, skip, command...
  n=$1; k=$2; sk=$3; shift 3
  timeout 300 $NCU -k regex:$k --launch-skip $sk -o $O/$n -f "$@" > $O/$n.log 2>&1
  python profiles/ncu_summary.py $O/$n.ncu-rep 30 > $O/$n.txt 2>&1
}
for p in h3 oh3 ch4h brh2 o3 ch4oh; do cap egrad_$p egrad_kernel 0 python profiles/prof_egrad.py $p; rm -f $O/egrad_$p.ncu-rep; done
cap verlet_h3_nb16 verlet_kernel 1 python profiles/prof_h3.py 16384 50; rm -f $O/verlet_h3_nb16.ncu-rep
cap qm_inter qm_inter 0 python profiles/prof_qmdff.py; rm -f $O/qm_inter.ncu-rep
cap qm_hb_search qm_hb_search 0 python profiles/prof_qmdff.py; rm -f $O/qm_hb_search.ncu-rep
cap recross_ch4h_nb16_1000 recross_kernel 1 python profiles/prof_recross.py 1000 512
python profiles/ncu_traffic.py $O/recross_ch4h_nb16_1000.ncu-rep 1000 512 > $O/traffic.log 2>&1; cp profiles/traffic_recross.json $O/
rm -f $O/recross_ch4h_nb16_1000.ncu-rep
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
ls -la $O
