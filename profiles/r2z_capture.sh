# round 2, capture Z (1 GPU), closing validation at HEAD: six processes call build_if_needed at once while a translation unit is
# stale (the race that corrupted the library under torchrun in capture r2s; now a file lock + atomic rename), then the whole
# GPU suite, smoke, the default bench line with its CPU baseline, the reference arm, PES-only bench (Br + H2 with three CTAs)
set -x
O=gpurun_out/r2z
mkdir -p $O
# NOTE (found afterwards): there is no comm.cu -- this removed nothing, no unit was stale in this capture; r2ad repeats it
# with a stamp that exists
rm -f caracal_b200/build/comm.o.sha
for i in 1 2 3 4 5 6; do
  (python -c "import caracal_b200; caracal_b200.build_if_needed(); caracal_b200.load(); print('rank-like process $i: library loaded')" > $O/lock_$i.log 2>&1 &)
done
sleep 1
python -c "import caracal_b200, time; caracal_b200.build_if_needed(); caracal_b200.load(); print('main process: library loaded')" > $O/lock_0.log 2>&1
sleep 20
cat $O/lock_*.log > $O/build_lock.log; rm -f $O/lock_*.log
python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_n1.json 2> $O/bench_reference_n1.err
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
python profiles/bench_egrad.py $O/bench_egrad.json > $O/bench_egrad.log 2>&1
ls -la $O
