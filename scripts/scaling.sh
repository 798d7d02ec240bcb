#!/usr/bin/env bash
# 1 -> N GPU scaling of bench.py on ONE box (run under `gpurun --gpus 8`):  scripts/scaling.sh <tag> [workload]
set -uo pipefail
TAG="${1:-r1}"; WL="${2:-scaled}"
OUT=gpurun_out; mkdir -p $OUT
PORT=29600
for N in 8 4 2 1; do
  PORT=$((PORT + 1))
  if [ "$N" = 1 ]; then
    timeout 240 python bench.py --workload $WL --steps 4 --warmup 3 --no-cpu-baseline 2> $OUT/${TAG}_${WL}_n$N.err | tail -1 > $OUT/${TAG}_${WL}_n$N.json
  else
    timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --workload $WL --steps 4 --warmup 3 --verbose 2> $OUT/${TAG}_${WL}_n$N.err | tail -1 > $OUT/${TAG}_${WL}_n$N.json
  fi
  echo "N=$N: $(cut -c1-200 $OUT/${TAG}_${WL}_n$N.json)"
done
