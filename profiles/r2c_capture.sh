# round 2, capture C (1 GPU): cell sweep parity + A/B timing + ncu, the other BASELINE configurations through bench.py
# --config, register-cap variants of the headline child kernel
set -x
O=gpurun_out/r2c
mkdir -p $O
python -m pytest tests/test_gpu_round2.py tests/test_gpu_qmdff.py tests/test_gpu_qmdff_examples.py -q -m gpu > $O/pytest_gpu_round2.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_round2.log
python profiles/bench_qmdff_cells.py $O/bench_qmdff_cells.json > $O/bench_qmdff_cells.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -c 1"
cap() { n=$1; k=$2; sk=$3; shift 3
  timeout 300 $NCU -k regex:$k --launch-skip $sk -o $O/$n -f "$@" > $O/$n.log 2>&1
  python profiles/ncu_summary.py $O/$n.ncu-rep 30 > $O/$n.txt 2>&1; rm -f $O/$n.ncu-rep; }
cap qm_inter_cell qm_inter_cell 0 python profiles/prof_qmdff.py
CRCL_QM_CELL_M=2 cap qm_inter_cell_m2 qm_inter_cell 0 python profiles/prof_qmdff.py
for c in c1 c3 c4 c5; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > $O/bench_$c.json 2> $O/bench_$c.err
done
for v in r64_144 r64_168; do
  if [ -f caracal_b200/libcaracal_gpu_$v.so ]; then
    CRCL_LIB_PATH=$PWD/caracal_b200/libcaracal_gpu_$v.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_n1_$v.json 2> $O/bench_n1_$v.err
  fi
done
ls -la $O
