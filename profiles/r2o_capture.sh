# round 2, capture O (1 GPU): A/B of the linear centroid in the child loop (two barriers fewer per step), alone and
# with 256-thread CTAs; the ncu report of the headline kernel is kept this time for per-instruction stall reasons
set -x
O=gpurun_out/r2o
mkdir -p $O
for v in "" _cenlin _cenlin256 ""; do
  CRCL_LIB_PATH=$PWD/caracal_b200/libcaracal_gpu$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench${v}_$RANDOM.json 2> $O/bench$v.err
done
for v in _cenlin _cenlin256; do
  CRCL_LIB_PATH=$PWD/caracal_b200/libcaracal_gpu$v.so python -m pytest tests/test_gpu_recross.py tests/test_gpu_verlet.py tests/test_gpu_rate.py -q -m gpu > $O/pytest$v.log 2>&1; echo "pytest exit $?" >> $O/pytest$v.log
done
NCU="ncu --set full --clock-control none --import-source on -c 1"
timeout 400 $NCU -k regex:recross_kernel --launch-skip 1 -o $O/recross_ch4h_nb16_200 -f python profiles/prof_recross.py 200 512 > $O/recross_ch4h_nb16_200.log 2>&1
python profiles/ncu_summary.py $O/recross_ch4h_nb16_200.ncu-rep 30 > $O/recross_ch4h_nb16_200.txt 2>&1
ls -la $O
