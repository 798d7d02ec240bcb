#!/usr/bin/env bash
# 2-GPU sanity: NCCL parity test + the default bench line at N=2 (with the scaled sub-record).  scripts/r3_multi2.sh <tag>
set -uo pipefail
TAG="${1:-r3m2}"; N=2; OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m pytest tests/test_dist_gpu.py -x -q -m gpu > $OUT/${TAG}_dist_pytest.log 2>&1
echo "dist pytest rc $?"; tail -3 $OUT/${TAG}_dist_pytest.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 50 --warmup 5 > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2.err
python - <<PY
import json
try:
  d=json.loads([l for l in open("$OUT/${TAG}_bench_n2.json") if l.startswith("{")][-1])
  print("N=2", d["value"], "steps/s", d["ms_per_step"], "ms e2e", d["e2e"]["value"], d["clocks"])
  s=d.get("scaled") or {}; print("scaled", {k: s.get(k) for k in ("ms_per_step","value","ms_per_step_at_max_clock","clocks")})
except Exception as e:
  print("FAILED", e); print(open("$OUT/${TAG}_bench_n2.err").read()[-1500:])
PY
