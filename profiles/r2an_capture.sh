# round 2, capture AN (1 GPU): spread form for 2 / 4 / 8-bead trajectories of the one-lane surfaces (the recrossing parent of the
# H + H2 example is one 8-bead trajectory) -- GPU suite, the rate example with three seeds (Monte-Carlo spread of kappa)
set -x
O=gpurun_out/r2an
mkdir -p $O
python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 300 python profiles/rate_h3.py $O/rate_h3_nb8_exact_norot.json 8 exact norot > $O/rate_h3_exact.log 2>&1
timeout 300 python profiles/rate_h3.py $O/rate_h3_nb8_exact_norot_seed2.json 8 exact norot 20250102 > $O/rate_h3_exact_seed2.log 2>&1
timeout 300 python profiles/rate_h3.py $O/rate_h3_nb8_exact_norot_seed3.json 8 exact norot 777 > $O/rate_h3_exact_seed3.log 2>&1
ls -la $O
