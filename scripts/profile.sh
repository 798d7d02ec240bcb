#!/usr/bin/env bash
# ncu evidence for profiles/ (run under gpurun on ONE GPU):  scripts/profile.sh <tag>
#   1. launch list of two eager ELBO steps at the Split-MNIST shape (per-launch device time: compare SHARES)
#   2. ncu --set full of the dominant kernels: gemm_tc (Split shape), gemm_tc2 + the streaming kernels (scaled shape)
set -uo pipefail
TAG="${1:-r1}"
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum -s 520 -c 170 --csv --log-file $OUT/${TAG}_launches_split.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-graph > $OUT/${TAG}_launches_split.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:gemm_tc_kernel -s 40 -c 4 -f -o $OUT/${TAG}_gemm_tc_split \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $OUT/${TAG}_gemm_tc_split.log 2>&1
timeout 500 $NCU --set full --import-source on -k regex:gemm_tc2_kernel -s 22 -c 3 -f -o $OUT/${TAG}_gemm_tc2_scaled \
  python bench.py --workload scaled --batch 8192 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_gemm_tc2_scaled.log 2>&1
timeout 500 $NCU --set full --import-source on -k 'regex:marginal_reduce|marginal_bwd_prep|rbf_bwd_prep|scale_rows' -s 10 -c 6 -f \
  -o $OUT/${TAG}_stream_scaled \
  python bench.py --workload scaled --batch 16384 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_stream_scaled.log 2>&1
ls -la $OUT/${TAG}_*
