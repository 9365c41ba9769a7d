# round 2, capture Y (one box, 8 GPUs) at HEAD: bench at N = 8 started while one translation unit is stale (the situation that
# corrupted the library in capture r2s: every rank calls build_if_needed; now serialised by a file lock), then N = 1 on the
# same box
set -x
O=gpurun_out/r2y
mkdir -p $O
# NOTE (found afterwards): there is no comm.cu -- this removed nothing, no unit was stale in this capture; r2ad repeats it
# with a stamp that exists
rm -f caracal_b200/build/comm.o.sha
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n1_samebox.json 2> $O/bench_n1_samebox.err
$TR --nproc-per-node 8 --master-port 29532 tests/multi_gpu_comm.py > $O/multi_gpu_comm_n8.log 2>&1; echo "exit $?" >> $O/multi_gpu_comm_n8.log
ls -la $O
