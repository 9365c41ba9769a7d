# round 2, capture AQ (one box, 2 GPUs) at HEAD: the bench line under torchrun at N = 2 (weak and strong legs, the collective
# inside the library) after the kernel changes of the third pass
set -x
O=gpurun_out/r2aq
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 120 $TR --nproc-per-node 2 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n2.json 2> $O/bench_n2.err
ls -la $O
