#!/usr/bin/env bash
# quick GPU iteration: kernel + model parity, bench line, timeline.   scripts/r2_quick.sh <tag> [pytest args]
set -uo pipefail
TAG="${1:-r2q}"; shift || true
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest ${@:-tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_large_gpu.py tests/test_train_gpu.py} -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc $?"; tail -6 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-scaled --detail > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc $?"; python - <<PY
import json
try:
  d=json.load(open("$OUT/${TAG}_bench.json")); print(d["value"], "steps/s", d["ms_per_step"], "ms  e2e", d["e2e"]["value"], d["clocks"]["sm_mhz"], "launches/step", d["gpu_launches"]/d["steps"])
except Exception as e:
  print("bench FAILED", e); print(open("$OUT/${TAG}_bench.err").read()[-2000:])
PY
timeout 200 python scripts/timeline.py --out $OUT/${TAG}_timeline.json > $OUT/${TAG}_timeline.txt 2>&1
echo "timeline rc $?"; sed -n 3,3p $OUT/${TAG}_timeline.txt
