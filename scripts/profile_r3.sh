#!/usr/bin/env bash
# Round-3 (second half of round 2) evidence for the cluster Cholesky, one GPU, under gpurun:  scripts/profile_r3.sh <tag>
#   1. launch list (gpu__time_duration) of eager steps of the default bench command: compare SHARES
#   2. ncu --set full of potrf_inv_cluster_kernel (Split shape, P=300) and of gemm_tc in the same step
#   3. compute-sanitizer memcheck / racecheck / synccheck over the cluster-kernel tests
set -uo pipefail
TAG="${1:-r3p}"; OUT=gpurun_out; mkdir -p $OUT
NCU="ncu --clock-control none"
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-scaled --no-graph"
timeout 300 $NCU --metrics gpu__time_duration.sum -s 200 -c 110 --csv --log-file $OUT/${TAG}_launches_split.csv $BENCH > $OUT/${TAG}_launches_split.log 2>&1
echo "launch list rc $?"; python scripts/ncu_summary.py launches $OUT/${TAG}_launches_split.csv > $OUT/${TAG}_launches_split.txt 2>&1; head -14 $OUT/${TAG}_launches_split.txt
cap() { name=$1; regex=$2; skip=$3; cnt=$4; shift 4
  timeout 600 $NCU --set full --import-source on -k "regex:$regex" -s $skip -c $cnt -f -o $OUT/${TAG}_$name "$@" > $OUT/${TAG}_$name.log 2>&1
  echo "$name rc $?"
  python scripts/ncu_summary.py rep $OUT/${TAG}_$name.ncu-rep > $OUT/${TAG}_${name}_ncu_full.txt 2>&1
  ls -la $OUT/${TAG}_$name.ncu-rep | awk '{print $5}'
}
cap potrf_cluster_split 'potrf_inv_cluster' 6 2 $BENCH
cap gemm_tc_split 'gemm_tc_kernel|gemm_tcp_kernel' 72 12 $BENCH
CS=/usr/local/cuda/bin/compute-sanitizer
K='chol_inv_cluster or (chol_inv_blocked and cluster and (60-30 or 33-3 or 97-4 or 129-3 or 200-5 or 257-2 or 320-3 or 600-4))'
timeout 900 $CS --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "$K" > $OUT/${TAG}_memcheck_cluster.log 2>&1
echo "memcheck rc $?"; grep -E "ERROR SUMMARY|passed|failed" $OUT/${TAG}_memcheck_cluster.log | tail -3
timeout 900 $CS --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "chol_inv_blocked and cluster and (60-30 or 33-3 or 97-4 or 129-3 or 257-2)" > $OUT/${TAG}_racecheck_cluster.log 2>&1
echo "racecheck rc $?"; grep -E "RACECHECK SUMMARY|passed|failed" $OUT/${TAG}_racecheck_cluster.log | tail -3
timeout 600 $CS --tool synccheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "chol_inv_blocked and cluster and (60-30 or 97-4 or 129-3 or 257-2)" > $OUT/${TAG}_synccheck_cluster.log 2>&1
echo "synccheck rc $?"; grep -E "ERROR SUMMARY|passed|failed" $OUT/${TAG}_synccheck_cluster.log | tail -3
ls -la $OUT/${TAG}_* | awk '{print $5, $9}'
