#!/usr/bin/env bash
set -uo pipefail
TAG="${1:-r2m}"; N="${2:-2}"; OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_dist_gpu.py -x -q -m gpu 2>&1 | tail -2
run() { name=$1; shift; extra="$1"; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 50 --warmup 5 $extra > $OUT/${TAG}_bench_n${N}_$name.json 2> $OUT/${TAG}_bench_n${N}_$name.err
  python - <<PY
import json
try:
  d=json.loads([l for l in open("$OUT/${TAG}_bench_n${N}_$name.json") if l.startswith("{")][-1])
  print("N=$N $name", d["value"], "steps/s", d["ms_per_step"], "ms e2e", d["e2e"]["value"], "graph", d["cuda_graph"], d["collectives_in_graph"])
except Exception as e:
  print("FAILED", e); print(open("$OUT/${TAG}_bench_n${N}_$name.err").read()[-1500:])
PY
}
run peer "--no-scaled" X=1
run nccl_eager "--no-scaled" VARGP_PEER_ALLREDUCE=0 VARGP_GRAPH_NCCL=0
run peer2 "--no-scaled" X=1
run nosync "--no-scaled" VARGP_NO_ALLREDUCE=1
