#!/usr/bin/env bash
# Round-2 A/B of the opt-in schedule knobs at the default workload (one GPU):  scripts/ab_sched.sh <tag>
#   1. parity of the model tests with the knobs on, 2. bench lines per knob, 3. block size of the blocked factorisation
set -uo pipefail
TAG="${1:-r2}"; OUT=gpurun_out; mkdir -p $OUT
VARGP_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_experimental_gpu.py -q -m gpu 2>&1 | tail -8
VARGP_STACK_CLASSES=1 VARGP_V_SIDE=1 timeout 200 python -m pytest tests/test_model_gpu.py tests/test_train_gpu.py -x -q -m gpu 2>&1 | tail -3
run() { name=$1; shift; env "$@" timeout 100 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2> $OUT/${TAG}_ab_$name.err | tail -1 > $OUT/${TAG}_ab_$name.json
  python - <<PY
import json; d=json.load(open("$OUT/${TAG}_ab_$name.json")); print("$name", d["value"], "steps/s", d["ms_per_step"], "ms  e2e", d["e2e"]["value"], d["clocks"]["sm_mhz"])
PY
}
run base VARGP_NOOP=1
run stack VARGP_STACK_CLASSES=1
run vside VARGP_V_SIDE=1
run both VARGP_STACK_CLASSES=1 VARGP_V_SIDE=1
run nb96 VARGP_CHOL_BLOCK=96
run nb64 VARGP_CHOL_BLOCK=64
run tcs300 VARGP_TCS_MAX_CTAS=300
run tcs450 VARGP_TCS_MAX_CTAS=450
run tcs700 VARGP_TCS_MAX_CTAS=700
run tcs450_all VARGP_TCS_MAX_CTAS=450 VARGP_STACK_CLASSES=1 VARGP_V_SIDE=1
