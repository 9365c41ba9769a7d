# round 2, capture D (1 GPU): cell sweep with per-warp queues + c6 classes: parity, A/B timing, ncu; configs 4 and 5
set -x
O=gpurun_out/r2d
mkdir -p $O
python -m pytest tests/test_gpu_round2.py tests/test_gpu_qmdff.py tests/test_gpu_qmdff_examples.py tests/test_gpu_dgevb.py -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
python profiles/bench_qmdff_cells.py $O/bench_qmdff_cells.json > $O/bench_qmdff_cells.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -c 1"
cap() { n=$1; k=$2; sk=$3; shift 3
  timeout 300 $NCU -k regex:$k --launch-skip $sk -o $O/$n -f "$@" > $O/$n.log 2>&1
  python profiles/ncu_summary.py $O/$n.ncu-rep 30 > $O/$n.txt 2>&1; rm -f $O/$n.ncu-rep; }
CRCL_QM_CELL_M=2 cap qm_inter_cell_m2 qm_inter_cell 0 python profiles/prof_qmdff.py
CRCL_QM_CELLS=0 cap qm_inter_n2_c6cls qm_inter_kernel 0 python profiles/prof_qmdff.py
for c in c4 c5; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > $O/bench_$c.json 2> $O/bench_$c.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_c5.csv python bench.py --config c5 --steps 2 --warmup 1 --no-cpu-baseline > $O/launches_c5.log 2>&1
ls -la $O
