#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tests/dist_gpu_worker.py > gpurun_out/r2distdbg.log 2>&1
echo "worker exit $?"
grep -v "Warning\|warn\|^  " gpurun_out/r2distdbg.log | tail -25 | cut -c1-250
