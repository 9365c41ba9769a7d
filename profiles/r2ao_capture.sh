# round 2, capture AO (1 GPU), closing validation at HEAD: spread forms limited to batches of at most 256 trajectories x beads
# (crcl_set_spread_max_beads) -- whole GPU suite, smoke, reference arm, default bench line with its CPU baseline, config 1,
# biased step by mode, umbrella phase of config 2, the H + H2 rate example, chain links
set -x
O=gpurun_out/r2ao
mkdir -p $O
python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_n1.json 2> $O/bench_reference_n1.err
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python bench.py --config c1 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c1.json 2> $O/bench_c1.err
python profiles/bench_umbrella_step.py $O/umbrella_step_ch4h.json > $O/umbrella_step_ch4h.log 2>&1
python profiles/umbrella_multi_gpu.py $O/umbrella_n1.json > $O/umbrella_n1.log 2>&1
timeout 300 python profiles/rate_h3.py $O/rate_h3_nb8_exact_norot.json 8 exact norot > $O/rate_h3_exact.log 2>&1
for pes in h3 ch4h oh3; do timeout 200 python profiles/prof_chain_h3.py 5000 0 $pes >> $O/chain_time.log 2>&1; done
ls -la $O
