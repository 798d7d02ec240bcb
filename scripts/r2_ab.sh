#!/usr/bin/env bash
# A/B of environment knobs at the default workload (AB_ARGS="--workload permuted_mnist --steps 30 --warmup 5" for another):
#   scripts/r2_ab.sh <tag> "NAME:ENV1=v,ENV2=v" ...
set -uo pipefail
TAG="$1"; shift; OUT=gpurun_out; mkdir -p $OUT
for spec in "$@"; do
  name="${spec%%:*}"; envs="${spec#*:}"
  env ${envs//,/ } timeout 120 python bench.py ${AB_ARGS:---steps 50 --warmup 5} --no-cpu-baseline --no-scaled 2> $OUT/${TAG}_ab_$name.err | tail -1 > $OUT/${TAG}_ab_$name.json
  python - <<PY
import json
try:
  d=json.load(open("$OUT/${TAG}_ab_$name.json")); print("$name", d["value"], "steps/s", d["ms_per_step"], "ms  e2e", d["e2e"]["value"], d["clocks"]["sm_mhz"])
except Exception as e:
  print("$name FAILED", e)
PY
done
