# round 2, capture H (1 GPU): where a biased step spends its time -- ncu source-level capture of the umbrella kernel, kept
set -x
O=gpurun_out/r2h
mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on -c 1"
timeout 300 $NCU -k regex:verlet_kernel --launch-skip 1 -o $O/verlet_ch4h_umbrella_1110 -f python profiles/prof_umbrella.py 50 100 10 > $O/u1110.log 2>&1
timeout 300 $NCU -k regex:verlet_kernel --launch-skip 1 -o $O/verlet_ch4h_umbrella_111 -f python profiles/prof_umbrella.py 50 100 1 > $O/u111.log 2>&1
ls -la $O
