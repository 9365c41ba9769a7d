# round 2, capture AJ (1 GPU): branch-free quotients / roots in xi.cuh and one-bead trajectories of one-lane surfaces
# spread over 16 lanes (PesSpread) -- whole GPU suite, the chain link timing, the H + H2 rate example at full size
set -x
O=gpurun_out/r2aj
mkdir -p $O
python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 200 python profiles/prof_chain_h3.py 10000 0 > $O/chain_h3_time.log 2>&1
timeout 200 python profiles/prof_chain_h3.py 10000 3 >> $O/chain_h3_time.log 2>&1
timeout 300 python profiles/rate_h3.py $O/rate_h3_nb8_exact_norot.json 8 exact norot > $O/rate_h3_exact.log 2>&1
ls -la $O
