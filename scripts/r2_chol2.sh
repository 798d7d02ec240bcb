#!/usr/bin/env bash
set -uo pipefail
TAG="${1:-r2u}"
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "chol" > $OUT/${TAG}_chol_pytest.log 2>&1
echo "chol pytest rc $?"; tail -5 $OUT/${TAG}_chol_pytest.log
timeout 200 python scripts/bench_chol.py 60 120 180 240 300 > $OUT/${TAG}_bench_chol.txt 2>&1
echo "bench_chol rc $?"; cat $OUT/${TAG}_bench_chol.txt | tail -12
timeout 120 python scripts/chol_stamps.py 300 > $OUT/${TAG}_stamps300.txt 2>&1; head -24 $OUT/${TAG}_stamps300.txt
