# round 2, capture J (1 GPU): branch-free FP64 math (fm::) in the CBE surfaces, child step with two barriers and the
# centroid from the staged sums, masses out of the registers -- whole GPU suite, headline bench, A/B variants, ncu
set -x
O=gpurun_out/r2j
mkdir -p $O
python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
for v in libm nofuse; do
  CRCL_LIB_PATH=$PWD/caracal_b200/libcaracal_gpu_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_$v.json 2> $O/bench_$v.err
done
NCU="ncu --set full --clock-control none --import-source on -c 1"
timeout 400 $NCU -k regex:recross_kernel --launch-skip 1 -o $O/recross_ch4h_nb16_1000 -f python profiles/prof_recross.py 1000 512 > $O/recross_ch4h_nb16_1000.log 2>&1
python profiles/ncu_summary.py $O/recross_ch4h_nb16_1000.ncu-rep 60 > $O/recross_ch4h_nb16_1000.txt 2>&1
python profiles/ncu_traffic.py $O/recross_ch4h_nb16_1000.ncu-rep 1000 512 > $O/traffic.log 2>&1; cp profiles/traffic_recross.json $O/
rm -f $O/recross_ch4h_nb16_1000.ncu-rep
python profiles/bench_egrad.py $O/bench_egrad.json > $O/bench_egrad.log 2>&1
ls -la $O
