# round 2, capture AH (1 GPU): source-level ncu capture of the umbrella-phase kernel (verlet_kernel<PesCBE4<K6>,16>, constrain 0),
# report kept for per-line analysis
set -x
O=gpurun_out/r2ah
mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on -c 1 -k regex:verlet_kernel --launch-skip 1 -o $O/verlet_umbrella -f python profiles/prof_umbrella.py > $O/verlet_umbrella.log 2>&1
python profiles/ncu_summary.py $O/verlet_umbrella.ncu-rep 40 > $O/verlet_umbrella.txt 2>&1
ls -la $O
