# round 2, capture AB (one box, 2 GPUs) at HEAD: bench at N = 2 under torchrun started while one translation unit is stale (every
# rank calls build_if_needed; the race of capture r2s, now behind a file lock), the collective at 2 ranks, N = 1 on the same box
set -x
O=gpurun_out/r2ab
mkdir -p $O
# NOTE (found afterwards): there is no comm.cu -- this removed nothing, no unit was stale in this capture; r2ad repeats it
# with a stamp that exists
rm -f caracal_b200/build/comm.o.sha
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n2.json 2> $O/bench_n2.err
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n1_samebox.json 2> $O/bench_n1_samebox.err
$TR --nproc-per-node 2 --master-port 29542 tests/multi_gpu_comm.py > $O/multi_gpu_comm_n2.log 2>&1; echo "exit $?" >> $O/multi_gpu_comm_n2.log
ls -la $O
