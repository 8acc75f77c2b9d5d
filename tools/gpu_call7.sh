#!/bin/bash
# Round-1 GPU session G (2 GPUs): tests, N=1 bench, N=2 bench through torchrun (NCCL grad all-reduce inside the graph).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $O/g_pytest.log 2>&1
echo "pytest exit $?" >> $O/g_pytest.log
tail -n 4 $O/g_pytest.log | cut -c1-200
timeout 420 python bench.py --steps 10 --warmup 3 > $O/g_bench_n1.json 2> $O/g_bench_n1.err
echo "bench n1 exit $?"
cut -c1-200 $O/g_bench_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 10 --warmup 3 > $O/g_bench_n2.json 2> $O/g_bench_n2.err
echo "bench n2 exit $?"
tail -n 1 $O/g_bench_n2.json | cut -c1-300
tail -n 3 $O/g_bench_n2.err | cut -c1-300
