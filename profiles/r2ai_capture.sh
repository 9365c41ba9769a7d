# round 2, capture AI (1 GPU): one link of the start-structure chain (one-bead H + H2 trajectory, biased steps):
# wall time per step, then a source-level ncu capture of verlet_kernel<PesH3,1>
set -x
O=gpurun_out/r2ai
mkdir -p $O
timeout 200 python profiles/prof_chain_h3.py 10000 0 > $O/chain_h3_time.log 2>&1
timeout 200 python profiles/prof_chain_h3.py 10000 3 >> $O/chain_h3_time.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -c 1 -k regex:verlet_kernel --launch-skip 1 -o $O/chain_h3 -f python profiles/prof_chain_h3.py 2000 0 > $O/chain_h3.log 2>&1
python profiles/ncu_summary.py $O/chain_h3.ncu-rep 60 > $O/chain_h3.txt 2>&1
ls -la $O
