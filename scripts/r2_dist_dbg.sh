#!/usr/bin/env bash
set -uo pipefail
OUT=gpurun_out; mkdir -p $OUT
for parts in grad graph; do
  VARGP_DIST_PARTS=$parts timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) tests/dist_worker_gpu.py > $OUT/r2k_dist_$parts.out 2> $OUT/r2k_dist_$parts.err
  echo "parts=$parts rc $?"; grep -v "^\*\|OMP_NUM\|^W1\|^$" $OUT/r2k_dist_$parts.err | tail -12; tail -2 $OUT/r2k_dist_$parts.out
done
