#!/usr/bin/env bash
# Round-2 GPU call 1: new parity tests, new bench line (with the scaled sub-record), A/B of the opt-in knobs.
set -uo pipefail
TAG=r2a; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > $OUT/${TAG}_gpu.txt 2>&1
timeout 600 python -m pytest tests/test_large_gpu.py tests/test_train_gpu.py tests/test_model_gpu.py -q -m gpu --durations=8 > $OUT/${TAG}_pytest_new.log 2>&1
echo "pytest new rc $?"; tail -15 $OUT/${TAG}_pytest_new.log
VARGP_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_experimental_gpu.py -q -m gpu > $OUT/${TAG}_pytest_exp.log 2>&1
echo "pytest exp rc $?"; tail -8 $OUT/${TAG}_pytest_exp.log
timeout 400 python bench.py --steps 50 --warmup 5 --verbose > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc $?"; cut -c1-600 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
run() { name=$1; shift; env "$@" timeout 100 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-scaled 2> $OUT/${TAG}_ab_$name.err | tail -1 > $OUT/${TAG}_ab_$name.json
  python - <<PY
import json
try:
  d=json.load(open("$OUT/${TAG}_ab_$name.json")); print("$name", d["value"], "steps/s", d["ms_per_step"], "ms  e2e", d["e2e"]["value"], d["clocks"]["sm_mhz"])
except Exception as e:
  print("$name FAILED", e)
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
