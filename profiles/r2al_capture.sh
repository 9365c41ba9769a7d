# round 2, capture AL (1 GPU): London term of the H + H2 surface shared by the lanes of a spread one-bead trajectory, the rare
# compact branch on copies, transrot sums without run-time indices -- GPU suite, chain link, biased step by mode, umbrella
# phase of config 2, rate example, default bench line
set -x
O=gpurun_out/r2al
mkdir -p $O
python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 200 python profiles/prof_chain_h3.py 10000 0 > $O/chain_h3_time.log 2>&1
timeout 200 python profiles/prof_chain_h3.py 10000 3 >> $O/chain_h3_time.log 2>&1
python profiles/bench_umbrella_step.py $O/umbrella_step_ch4h.json > $O/umbrella_step_ch4h.log 2>&1
python profiles/umbrella_multi_gpu.py $O/umbrella_n1.json > $O/umbrella_n1.log 2>&1
timeout 300 python profiles/rate_h3.py $O/rate_h3_nb8_exact_norot.json 8 exact norot > $O/rate_h3_exact.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
ls -la $O
