#!/usr/bin/env bash
# NVLS (multimem.ld_reduce) form of the fused gradient all-reduce + Yogi: parity test, then the tails side by side.  scripts/r4_nvls.sh <tag> <N>
set -uo pipefail
TAG="${1:-r4k}"; N="${2:-2}"; OUT=gpurun_out; mkdir -p $OUT
if [ "$N" = "2" ]; then timeout 400 python -m pytest tests/test_dist_gpu.py -x -q -m gpu > $OUT/${TAG}_dist_pytest.log 2>&1; fi
echo "dist pytest rc $?"; tail -3 $OUT/${TAG}_dist_pytest.log
run() { name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 50 --warmup 5 --no-scaled --no-cpu-baseline > $OUT/${TAG}_bench_n${N}_$name.json 2> $OUT/${TAG}_bench_n${N}_$name.err
  python - <<PY
import json
try:
  d=json.loads([l for l in open("$OUT/${TAG}_bench_n${N}_$name.json") if l.startswith("{")][-1])
  print("N=$N $name", d["value"], "steps/s", d["ms_per_step"], "ms e2e", d["e2e"]["value"], "graph", d["cuda_graph"], d["collectives_in_graph"])
except Exception as e:
  print("$name FAILED", e); print(open("$OUT/${TAG}_bench_n${N}_$name.err").read()[-1500:])
PY
}
run nccl_eager X=1
run peer_nvls VARGP_PEER_ALLREDUCE=1
run peer_loads VARGP_PEER_ALLREDUCE=1 VARGP_PEER_NVLS=0
run nosync VARGP_NO_ALLREDUCE=1


