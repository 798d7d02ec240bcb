#!/usr/bin/env bash
# full single-GPU evidence run: scripts/r2_full.sh <tag>
set -uo pipefail
TAG="${1:-r2p}"; OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -q -m gpu --durations=15 > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc $?"; tail -22 $OUT/${TAG}_pytest_gpu.log
timeout 300 python scripts/parity_report.py > $OUT/${TAG}_parity_report.txt 2> $OUT/${TAG}_parity_report.err
echo "parity rc $?"; tail -3 $OUT/${TAG}_parity_report.err
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
echo "smoke rc $?"; tail -1 $OUT/${TAG}_smoke.log
