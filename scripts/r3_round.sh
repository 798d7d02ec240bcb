#!/usr/bin/env bash
# full GPU suite + default bench line (with scaled sub-record) + Permuted bench.   scripts/r3_round.sh <tag>
set -uo pipefail
TAG="${1:-r3}"
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc $?"; tail -4 $OUT/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 --detail > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err
echo "bench rc $?"; python - <<PY
import json
try:
  d=json.load(open("$OUT/${TAG}_bench_default.json")); print(d["value"], "steps/s", d["ms_per_step"], "ms  e2e", d["e2e"]["value"], d["clocks"], "roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "cpu", d.get("cpu_baseline",{}).get("value"))
  s=d.get("scaled") or {}; print("scaled", {k: s.get(k) for k in ("ms_per_step","value","ms_per_step_at_max_clock","clocks")})
except Exception as e:
  print("bench FAILED", e); print(open("$OUT/${TAG}_bench_default.err").read()[-2000:])
PY
timeout 300 python bench.py --workload permuted_mnist --steps 30 --warmup 5 --no-cpu-baseline --no-scaled --detail > $OUT/${TAG}_bench_permuted.json 2> $OUT/${TAG}_bench_permuted.err
echo "permuted rc $?"; python - <<PY
import json
try:
  d=json.load(open("$OUT/${TAG}_bench_permuted.json")); print(d["value"], "steps/s", d["ms_per_step"], "ms  e2e", d["e2e"]["value"], d["clocks"]["sm_mhz"])
except Exception as e:
  print("permuted FAILED", e); print(open("$OUT/${TAG}_bench_permuted.err").read()[-1500:])
PY
