# round 2, capture B (2 GPUs): the round-2 GPU tests incl. the multi-rank communicator, bench at N = 2 (weak + strong legs)
set -x
O=gpurun_out/r2b
mkdir -p $O
python -m pytest tests/test_gpu_round2.py tests/test_gpu_qmdff_examples.py tests/test_gpu_split.py -q -m gpu > $O/pytest_gpu_round2.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_round2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_comm.py > $O/multi_gpu_comm_n2.log 2>&1; echo "exit $?" >> $O/multi_gpu_comm_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
ls -la $O
