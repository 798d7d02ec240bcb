#!/usr/bin/env bash
# 8-GPU record: the default bench line (weak scaling at the Split shape + the scaled strong-scaling sub-record).
set -uo pipefail
TAG="${1:-r4j}"; N="${2:-8}"; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 50 --warmup 5 > $OUT/${TAG}_bench_n${N}.json 2> $OUT/${TAG}_bench_n${N}.err
echo "rc $?"; python - <<PY
import json
try:
  d=json.loads([l for l in open("$OUT/${TAG}_bench_n${N}.json") if l.startswith("{")][-1])
  print("N=$N", d["value"], "steps/s", d["ms_per_step"], "ms e2e", d["e2e"]["value"], d["clocks"])
  s=d.get("scaled"); print("   scaled", {k: s.get(k) for k in ("value","ms_per_step","ms_per_step_at_max_clock","clocks","error")} if s else None)
except Exception as e:
  print("FAILED", e); print(open("$OUT/${TAG}_bench_n${N}.err").read()[-1500:])
PY
