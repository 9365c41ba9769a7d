# round 2, capture AP (1 GPU): the umbrella-phase kernel (verlet_kernel<PesCBE4<K6>,16>, constrain 0) after transrot's sums left
# local memory -- the same source-level ncu capture as r2ah, for the before / after
set -x
O=gpurun_out/r2ap
mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on -c 1 -k regex:verlet_kernel --launch-skip 1 -o $O/verlet_umbrella -f python profiles/prof_umbrella.py > $O/verlet_umbrella.log 2>&1
python profiles/ncu_summary.py $O/verlet_umbrella.ncu-rep 30 > $O/verlet_umbrella.txt 2>&1
ls -la $O
