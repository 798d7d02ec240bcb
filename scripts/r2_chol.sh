#!/usr/bin/env bash
set -uo pipefail
TAG="${1:-r2g}"; OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k chol > $OUT/${TAG}_pytest_chol.log 2>&1
echo "pytest chol rc $?"; tail -12 $OUT/${TAG}_pytest_chol.log
timeout 200 python scripts/bench_chol.py 129 180 240 300 320 > $OUT/${TAG}_bench_chol.txt 2>&1; cat $OUT/${TAG}_bench_chol.txt
VARGP_CHOL_MID_MIN_N=1 timeout 200 python scripts/bench_chol.py 60 120 128 2>&1 | tee $OUT/${TAG}_bench_chol_small.txt
