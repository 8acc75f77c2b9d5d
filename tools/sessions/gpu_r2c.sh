#!/bin/bash
# round 2, session a: TS-mode GEMM validation + A/B bench + ncu
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2c_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -s --timeout 300 > $O/r2c_pytest_gemm.log 2>&1
echo "pytest gemm exit $?" | tee -a $O/r2c_pytest_gemm.log
tail -n 5 $O/r2c_pytest_gemm.log | cut -c1-300
timeout 600 python tools/gemm_bench.py > $O/r2c_gemm.log 2>&1
echo "gemm_bench exit $?"
python - <<'PY'
import json
for l in open('gpurun_out/r2c_gemm.log'):
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print({k:(round(v['us'],1) if isinstance(v,dict) else v) for k,v in d.items()})
PY
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > $O/r2c_pytest.log 2>&1
echo "pytest exit $?" | tee -a $O/r2c_pytest.log
tail -n 5 $O/r2c_pytest.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2c_bench_n1.json 2> $O/r2c_bench_n1.err
echo "bench exit $?"; cut -c1-400 $O/r2c_bench_n1.json; tail -n 3 $O/r2c_bench_n1.err | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tf32_ts|gemm_tn_ts" -s 2 -c 1 -f -o $O/r2c_gemm_ts python tools/gemm_bench.py --ncu > $O/r2c_ncu.log 2>&1
echo "ncu exit $?"; tail -n 2 $O/r2c_ncu.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_ts -s 1 -c 1 -f -o $O/r2c_gemm_tn python tools/gemm_bench.py --ncu > $O/r2c_ncu_tn.log 2>&1
echo "ncu tn exit $?"
grep -h "truncation" $O/r2c_pytest_gemm.log
