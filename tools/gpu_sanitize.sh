#!/bin/bash
# compute-sanitizer passes over the hand-written kernels on small shapes (SURVEY.md section 5 "race detection / sanitizers").
# Usage (GPU box): bash tools/gpu_sanitize.sh ; summaries land in gpurun_out/sanitize_*.log
set -u
mkdir -p gpurun_out
O=gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
SPMM='tests/test_gpu_spmm.py -k "rowgroup_long or epilogue_matches or cube_golden or mesh_operators_vs_oracle"'
GEMM='tests/test_gpu_gemm.py -k "(gemm_matches_fp64 and 1000-128-256) or (tn_matches and 3000) or colsum"'
FUSED='tests/test_gpu_fused_epilogues.py -k "(gemm_act and 1000-128-256) or (stats_store_path and shape0) or avg_block or chained"'
for tool in memcheck racecheck initcheck synccheck; do
  for grp in SPMM GEMM FUSED; do
    sel=${!grp}
    log=$O/sanitize_${tool}_${grp}.log
    if [ $tool = initcheck ] && [ $grp = FUSED ]; then continue; fi   # does not finish in 15 min (TMA-heavy kernels under initcheck)
    eval timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 99 python -m pytest $sel -x -q --timeout 850 -p no:cacheprovider > $log 2>&1
    rc=$?
    echo "== $tool $grp rc=$rc: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' $log | tr '\n' ' ' | cut -c1-300)"
  done
done
