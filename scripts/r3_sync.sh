#!/usr/bin/env bash
set -uo pipefail
TAG="${1:-r3y}"; OUT=gpurun_out; mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 600 $CS --tool synccheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "chol_inv_blocked and cluster and (60-30 or 97-4 or 129-3 or 257-2)" > $OUT/${TAG}_synccheck_cluster.log 2>&1
echo "synccheck rc $?"; grep -E "ERROR SUMMARY|passed|failed" $OUT/${TAG}_synccheck_cluster.log | tail -3
timeout 900 $CS --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "chol_inv_blocked and cluster and (60-30 or 33-3 or 97-4 or 129-3 or 257-2)" > $OUT/${TAG}_racecheck_cluster.log 2>&1
echo "racecheck rc $?"; grep -E "RACECHECK SUMMARY|passed|failed" $OUT/${TAG}_racecheck_cluster.log | tail -3
timeout 600 $CS --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gemm_tc_gpu.py -q -m gpu -x -k "128-128-96 or 132-68-44 or 256-256-64 or 60-512-60 or flags" > $OUT/${TAG}_memcheck_gemm_persist.log 2>&1
echo "memcheck gemm (default routing) rc $?"; grep -E "ERROR SUMMARY|passed|failed" $OUT/${TAG}_memcheck_gemm_persist.log | tail -3
VARGP_TC_PERSIST=2 timeout 600 $CS --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gemm_tc_gpu.py -q -m gpu -x -k "128-128-96 or 132-68-44 or 256-256-64 or 60-512-60 or flags" > $OUT/${TAG}_memcheck_gemm_persist2.log 2>&1
echo "memcheck gemm (persistent forced) rc $?"; grep -E "ERROR SUMMARY|passed|failed" $OUT/${TAG}_memcheck_gemm_persist2.log | tail -3
VARGP_TC_PERSIST=2 timeout 600 $CS --tool synccheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gemm_tc_gpu.py -q -m gpu -x -k "128-128-96 or 256-256-64" > $OUT/${TAG}_synccheck_gemm_persist2.log 2>&1
echo "synccheck gemm (persistent forced) rc $?"; grep -E "ERROR SUMMARY|passed|failed" $OUT/${TAG}_synccheck_gemm_persist2.log | tail -3
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k chol 2>&1 | tail -2
timeout 200 python scripts/bench_chol.py 60 120 300 > $OUT/${TAG}_bench_chol.txt 2>&1; cat $OUT/${TAG}_bench_chol.txt
