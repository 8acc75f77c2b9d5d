#!/bin/bash
# 2-GPU session: NCCL gradient parity test + N=2 bench through torchrun
set -u
mkdir -p gpurun_out
O=gpurun_out
T=r2w
nvidia-smi --query-gpu=index,name --format=csv > $O/${T}_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_graph.py -x -q --timeout 500 > $O/${T}_pytest_graph.log 2>&1
echo "pytest graph exit $?"; tail -n 5 $O/${T}_pytest_graph.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > $O/${T}_bench_n2.json 2> $O/${T}_bench_n2.err
echo "bench n2 exit $?"; tail -n 1 $O/${T}_bench_n2.json | cut -c1-400; tail -n 3 $O/${T}_bench_n2.err | cut -c1-200
