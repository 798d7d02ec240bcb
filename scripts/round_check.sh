#!/usr/bin/env bash
# One-GPU round check (run under gpurun):  scripts/round_check.sh <tag>
#   bench line at the default workload, GPU parity suite, smoke(), factor-shard check (2 ranks on one GPU over gloo),
#   ncu launch list of the bench command.  Everything lands in gpurun_out/<tag>_*.
set -uo pipefail
TAG="${1:-r1c}"
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > $OUT/${TAG}_gpu.txt 2>&1
timeout 300 python bench.py --steps 50 --warmup 5 > $OUT/${TAG}_bench_split.json 2> $OUT/${TAG}_bench_split.err
echo "bench rc $?"; cut -c1-400 $OUT/${TAG}_bench_split.json
timeout 700 python -m pytest tests -x -q -m gpu --durations=12 > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc $?"; tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
echo "smoke rc $?"; tail -1 $OUT/${TAG}_smoke.log
VARGP_CHECK_BACKEND=gloo timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
  --master-port 29511 scripts/check_shard_gpu.py 640 1024 > $OUT/${TAG}_shard_gloo.log 2>&1
echo "shard rc $?"; tail -1 $OUT/${TAG}_shard_gloo.log
timeout 120 python bench.py --workload permuted_mnist --steps 30 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_permuted.json 2> $OUT/${TAG}_bench_permuted.err
echo "permuted rc $?"; cut -c1-200 $OUT/${TAG}_bench_permuted.json
timeout 200 ncu --clock-control none --metrics gpu__time_duration.sum -s 520 -c 170 --csv --log-file $OUT/${TAG}_launches_split.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-graph > $OUT/${TAG}_launches_split.log 2>&1
echo "ncu rc $?"
