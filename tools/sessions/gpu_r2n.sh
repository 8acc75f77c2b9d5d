#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
T=r2n
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_fused_epilogues.py -x -q --timeout 300 > $O/${T}_pytest_gemm.log 2>&1
echo "pytest gemm exit $?"; tail -n 12 $O/${T}_pytest_gemm.log | cut -c1-300
summ() { python -c "
import sys,json
for l in open('$1'):
    if not l.startswith('{'): continue
    d=json.loads(l)
    print(d['M'],d['N'],d['K'],d['R'],'full',d['full'],'b2b',d['full_b2b'],'| noepi',d['no_epilogue'],'b2b',d['no_epilogue_b2b'],'| Aonly',d['only_A_stream'],'b2b',d['only_A_stream_b2b'])
"; }
timeout 600 python tools/gemm_stage_bench.py > $O/${T}_stage_default.log 2>&1; echo default; summ $O/${T}_stage_default.log
SN_GEMM_BSTAGES=3 timeout 600 python tools/gemm_stage_bench.py > $O/${T}_stage_b3.log 2>&1; echo "bstages=3"; summ $O/${T}_stage_b3.log
SN_GEMM_RSLOTS=1 timeout 600 python tools/gemm_stage_bench.py > $O/${T}_stage_r1.log 2>&1; echo "rslots=1"; summ $O/${T}_stage_r1.log
SN_GEMM_RSLOTS=1 SN_GEMM_BSTAGES=3 timeout 600 python tools/gemm_stage_bench.py > $O/${T}_stage_r1b3.log 2>&1; echo "rslots=1 bstages=3"; summ $O/${T}_stage_r1b3.log
timeout 600 python tools/fusion_bench.py > $O/${T}_fusion.json 2> $O/${T}_fusion.err
python -c "
import json
for k,v in json.load(open('$O/${T}_fusion.json')).items(): print('%-50s %s'%(k,v))"
