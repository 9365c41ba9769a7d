# round 2, capture AD (one box, 2 GPUs) at HEAD: a translation unit is made stale (its content stamp removed), then
#  - six plain processes call build_if_needed + dlopen at once,
#  - torchrun starts bench.py at N = 2 with the unit stale again (every rank calls build_if_needed: the race that corrupted the
#    library at N = 8 in capture r2s; now a file lock + atomic rename)
set -x
O=gpurun_out/r2ad
mkdir -p $O
ls -la caracal_b200/build/water_kernels.o.sha
rm -f caracal_b200/build/water_kernels.o.sha
for i in 1 2 3 4 5 6; do
  (python -c "import caracal_b200, ctypes; caracal_b200.build_if_needed(); ctypes.CDLL(caracal_b200.LIB_PATH); print('process $i: library loaded')" > $O/lock_$i.log 2>&1 &)
done
sleep 45
cat $O/lock_*.log > $O/build_lock.log; rm -f $O/lock_*.log
ls -la caracal_b200/build/water_kernels.o.sha >> $O/build_lock.log 2>&1
rm -f caracal_b200/build/water_kernels.o.sha
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 2 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n2.json 2> $O/bench_n2.err
ls -la caracal_b200/build/water_kernels.o.sha >> $O/build_lock.log 2>&1
ls -la $O
