#!/usr/bin/env bash
set -uo pipefail
TAG="${1:-r2u}"
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "chol" > $OUT/${TAG}_chol_pytest.log 2>&1
echo "chol pytest rc $?"; tail -5 $OUT/${TAG}_chol_pytest.log
timeout 300 python scripts/bench_chol.py 300 500 1000 2048 > $OUT/${TAG}_bench_chol.txt 2>&1
echo "bench_chol rc $?"; cat $OUT/${TAG}_bench_chol.txt | tail -12
