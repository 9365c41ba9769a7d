# round 2, capture AS (1 GPU): the spread one-bead kernel at HEAD (London term shared, compact branch on copies, transrot sums in registers) under ncu, source level: verlet_kernel<PesSpread<PesH3,16>,1>
set -x
O=gpurun_out/r2as
mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on -c 1 -k regex:verlet_kernel --launch-skip 1 -o $O/chain_h3 -f python profiles/prof_chain_h3.py 2000 0 > $O/chain_h3.log 2>&1
python profiles/ncu_summary.py $O/chain_h3.ncu-rep 70 > $O/chain_h3.txt 2>&1
ls -la $O
