#!/bin/bash
# One GPU session on a B200 box (gpurun -- 'bash tools/gpu_session.sh [n_gpus]'): smoke(), the GPU parity tests, the N=1
# bench line and -- with n_gpus > 1 -- the same bench through torchrun.  Outputs land in gpurun_out/ (scratch); copy what
# should be judged into profiles/.
set -u
NG=${1:-1}
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
tail -n 1 $O/smoke.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > $O/pytest.log 2>&1
echo "pytest exit $?" >> $O/pytest.log
tail -n 3 $O/pytest.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench n1 exit $?"
cut -c1-160 $O/bench_n1.json
if [ "$NG" -gt 1 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $NG --steps 10 --warmup 3 > $O/bench_n$NG.json 2> $O/bench_n$NG.err
  echo "bench n$NG exit $?"
  tail -n 1 $O/bench_n$NG.json | cut -c1-200
fi
