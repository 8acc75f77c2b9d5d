#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${1:-r2t}
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > $O/${T}_pytest.log 2>&1
echo "pytest exit $?" | tee -a $O/${T}_pytest.log
tail -n 12 $O/${T}_pytest.log | cut -c1-400
timeout 900 python bench.py --steps 10 --warmup 3 > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err
echo "bench exit $?"; cut -c1-200 $O/${T}_bench_n1.json; tail -n 3 $O/${T}_bench_n1.err | cut -c1-300
python tools/bench_table.py $O/${T}_bench_n1.json | head -50
timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool synccheck --print-limit 5 python -m pytest tests/test_gpu_gemm.py -k "gemm_matches_fp64 and 1000-128-256" -x -q -p no:cacheprovider > $O/${T}_synccheck_gemm.log 2>&1
echo "synccheck rc=$?: $(grep -E 'ERROR SUMMARY|passed|failed' $O/${T}_synccheck_gemm.log | tr '\n' ' ' | cut -c1-300)"
