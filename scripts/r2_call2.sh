#!/usr/bin/env bash
set -uo pipefail
TAG=r2b; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu --durations=10 > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc $?"; tail -16 $OUT/${TAG}_pytest_gpu.log
timeout 200 python scripts/timeline.py --out $OUT/${TAG}_timeline.json > $OUT/${TAG}_timeline.txt 2>&1
echo "timeline rc $?"; sed -n 3,4p $OUT/${TAG}_timeline.txt
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-scaled --detail > $OUT/${TAG}_bench_detail.json 2> $OUT/${TAG}_bench_detail.err
echo "bench rc $?"; cut -c1-300 $OUT/${TAG}_bench_detail.json
timeout 200 python bench.py --workload permuted_mnist --steps 30 --warmup 5 --no-cpu-baseline --detail > $OUT/${TAG}_bench_permuted.json 2> $OUT/${TAG}_bench_permuted.err
echo "permuted rc $?"; cut -c1-300 $OUT/${TAG}_bench_permuted.json
