#!/usr/bin/env bash
# cluster Cholesky: kernel tests, microbench, then the quick round (tests + bench + timeline).  scripts/r2_chol.sh <tag>
set -uo pipefail
TAG="${1:-r2t}"
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "chol" > $OUT/${TAG}_chol_pytest.log 2>&1
echo "chol pytest rc $?"; tail -15 $OUT/${TAG}_chol_pytest.log
timeout 200 python scripts/bench_chol.py 60 120 128 180 240 300 320 > $OUT/${TAG}_bench_chol.txt 2>&1
echo "bench_chol rc $?"; cat $OUT/${TAG}_bench_chol.txt | tail -12
bash scripts/r2_quick.sh $TAG
