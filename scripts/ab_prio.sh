#!/usr/bin/env bash
# A/B of the scheduling knobs at the default workload (one GPU): scripts/ab_prio.sh <tag>
set -uo pipefail
TAG="${1:-r1c}"; OUT=gpurun_out; mkdir -p $OUT
run() { name=$1; shift; env "$@" timeout 100 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2> $OUT/${TAG}_ab_$name.err | tail -1 > $OUT/${TAG}_ab_$name.json
  python - <<PY
import json; d=json.load(open("$OUT/${TAG}_ab_$name.json")); print("$name", d["value"], "steps/s", d["ms_per_step"], "ms  e2e", d["e2e"]["value"], d["clocks"]["sm_mhz"])
PY
}
run torch_replay VARGP_NODE_PRIO=0
run node_prio VARGP_NODE_PRIO=1
run node_prio_attr VARGP_NODE_PRIO=1 VARGP_PRIO_ATTR=1
