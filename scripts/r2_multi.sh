#!/usr/bin/env bash
# multi-GPU check: scripts/r2_multi.sh <tag> <ngpus>
set -uo pipefail
TAG="${1:-r2m}"; N="${2:-2}"; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_dist_gpu.py -x -q -m gpu > $OUT/${TAG}_pytest_dist.log 2>&1
echo "pytest dist rc $?"; tail -15 $OUT/${TAG}_pytest_dist.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 30 --warmup 5 --verbose > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
echo "bench rc $?"; tail -4 $OUT/${TAG}_bench_n$N.err; python - <<PY
import json
try:
  d=json.loads([l for l in open("$OUT/${TAG}_bench_n$N.json") if l.startswith("{")][-1])
  print("N=$N", d["value"], "steps/s", d["ms_per_step"], "ms e2e", d["e2e"]["value"], "graph", d["cuda_graph"], d["collectives_in_graph"])
  s=d.get("scaled"); print("scaled", {k: s.get(k) for k in ("value","ms_per_step","clocks","error")} if s else None)
except Exception as e:
  print("FAILED", e)
PY
VARGP_GRAPH_NCCL=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 30 --warmup 5 --no-scaled > $OUT/${TAG}_bench_n${N}_eager_tail.json 2> $OUT/${TAG}_bench_n${N}_eager_tail.err
python - <<PY
import json
try:
  d=json.loads([l for l in open("$OUT/${TAG}_bench_n${N}_eager_tail.json") if l.startswith("{")][-1])
  print("N=$N eager tail", d["value"], "steps/s", d["ms_per_step"], "ms")
except Exception as e:
  print("FAILED", e)
PY
