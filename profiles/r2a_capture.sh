set -x
mkdir -p gpurun_out/r2a
NCU="ncu --set full --clock-control none --import-source on -c 1"
for p in h3 oh3 ch4h brh2 o3 ch4oh; do
  timeout 300 $NCU -k regex:egrad_kernel -o gpurun_out/r2a/egrad_$p -f python profiles/prof_egrad.py $p > gpurun_out/r2a/egrad_$p.log 2>&1
done
timeout 300 $NCU -k regex:verlet_kernel --launch-skip 1 -o gpurun_out/r2a/verlet_h3 -f python profiles/prof_h3.py 16384 50 > gpurun_out/r2a/verlet_h3.log 2>&1
timeout 300 $NCU -k regex:qm_inter -o gpurun_out/r2a/qm_inter -f python profiles/prof_qmdff.py > gpurun_out/r2a/qm_inter.log 2>&1
timeout 300 $NCU -k regex:qm_hb_search -o gpurun_out/r2a/qm_hb_search -f python profiles/prof_qmdff.py > gpurun_out/r2a/qm_hb_search.log 2>&1
timeout 300 $NCU -k regex:recross_kernel --launch-skip 1 -o gpurun_out/r2a/recross_1000 -f python profiles/prof_recross.py 1000 512 > gpurun_out/r2a/recross_1000.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a/bench_n1.json 2> gpurun_out/r2a/bench_n1.err
ls -la gpurun_out/r2a
