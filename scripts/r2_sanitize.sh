#!/usr/bin/env bash
# compute-sanitizer over the kernel tests at small shapes: memcheck on everything small,
# racecheck + synccheck on the shared-memory / mbarrier / TMEM kernels.   scripts/r2_sanitize.sh <tag>
set -uo pipefail
TAG="${1:-r2s}"; OUT=gpurun_out; mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
SMALL='(128-128-32 or 128-128-96 or 132-68-44 or 60-512-60 or 256-256-64 or flags or rbf_epilogue and 64 or falls_back or block_views) and not 1000 and not 2048 and not 3000'
KERN='whiten or step_prologue or likelihood or marginal_kl or rbf_adjoint or tril or chol_reports or (chol_inv_blocked and (mid or 60-30 or 129-3 or 97-3 or 33-4)) or chol_inv_mid'
timeout 900 $CS --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gemm_tc_gpu.py -q -m gpu -x -k "$SMALL" > $OUT/${TAG}_memcheck_gemm.log 2>&1
echo "memcheck gemm rc $?"; grep -E "ERROR SUMMARY|passed|failed" $OUT/${TAG}_memcheck_gemm.log | tail -3
timeout 900 $CS --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "$KERN" > $OUT/${TAG}_memcheck_kernels.log 2>&1
echo "memcheck kernels rc $?"; grep -E "ERROR SUMMARY|passed|failed" $OUT/${TAG}_memcheck_kernels.log | tail -3
timeout 900 $CS --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "whiten or likelihood or chol_inv_mid or (chol_inv_blocked and (mid and 129 or 60-30)) or marginal_kl" > $OUT/${TAG}_racecheck_kernels.log 2>&1
echo "racecheck kernels rc $?"; grep -E "RACECHECK SUMMARY|passed|failed" $OUT/${TAG}_racecheck_kernels.log | tail -3
timeout 900 $CS --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gemm_tc_gpu.py -q -m gpu -x -k "128-128-96 or 132-68-44 or 256-256-64 or flags" > $OUT/${TAG}_racecheck_gemm.log 2>&1
echo "racecheck gemm rc $?"; grep -E "RACECHECK SUMMARY|passed|failed" $OUT/${TAG}_racecheck_gemm.log | tail -3
timeout 600 $CS --tool synccheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gemm_tc_gpu.py tests/test_kernels_gpu.py -q -m gpu -x -k "128-128-96 or 256-256-64 or whiten or chol_inv_mid or likelihood" > $OUT/${TAG}_synccheck.log 2>&1
echo "synccheck rc $?"; grep -E "ERROR SUMMARY|passed|failed" $OUT/${TAG}_synccheck.log | tail -3
