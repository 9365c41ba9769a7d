# round 2, capture S (one box, 8 GPUs) at HEAD: the collective at 8 ranks, config 2 weak + strong legs at N = 8, 4, 2, 1 on the
# same box, the umbrella phase over 8 GPUs
set -x
O=gpurun_out/r2s
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29521 tests/multi_gpu_comm.py > $O/multi_gpu_comm_n8.log 2>&1; echo "exit $?" >> $O/multi_gpu_comm_n8.log
$TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err
$TR --nproc-per-node 8 --master-port 29523 profiles/umbrella_multi_gpu.py $O/umbrella_n8.json > $O/umbrella_n8.log 2>&1
$TR --nproc-per-node 4 --master-port 29524 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n4.json 2> $O/bench_n4.err
$TR --nproc-per-node 2 --master-port 29525 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n2.json 2> $O/bench_n2.err
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n1_samebox.json 2> $O/bench_n1_samebox.err
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $O/smi.csv
ls -la $O
