#!/usr/bin/env bash
# Round-2 ncu evidence (one GPU, under gpurun):  scripts/profile_r2.sh <tag>
#   1. launch list (gpu__time_duration) of the default bench command's steps -- eager launches so that every kernel is its own
#      ncu launch; cold-cache and serialised: compare SHARES
#   2. ncu --set full of the kernels of the Split-MNIST-shaped step, summarised with scripts/ncu_summary.py
#   3. ncu --set full of gemm_tc2 + the streaming kernels at a slice of the scaled shape
set -uo pipefail
TAG="${1:-r2}"; OUT=gpurun_out; mkdir -p $OUT
NCU="ncu --clock-control none"
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-scaled --no-graph"
# 8 eager steps before the timed ones (3 + 3 warm-up, 2 timed follow): skip ~6 steps of ~50 launches, keep ~2.5 steps
timeout 300 $NCU --metrics gpu__time_duration.sum -s 300 -c 160 --csv --log-file $OUT/${TAG}_launches_split.csv $BENCH > $OUT/${TAG}_launches_split.log 2>&1
echo "launch list rc $?"; python scripts/ncu_summary.py launches $OUT/${TAG}_launches_split.csv > $OUT/${TAG}_launches_split.txt 2>&1; head -12 $OUT/${TAG}_launches_split.txt
cap() { name=$1; regex=$2; skip=$3; cnt=$4; shift 4
  timeout 600 $NCU --set full --import-source on -k "regex:$regex" -s $skip -c $cnt -f -o $OUT/${TAG}_$name "$@" > $OUT/${TAG}_$name.log 2>&1
  echo "$name rc $?"
  python scripts/ncu_summary.py rep $OUT/${TAG}_$name.ncu-rep > $OUT/${TAG}_${name}_ncu_full.txt 2>&1
  ls -la $OUT/${TAG}_$name.ncu-rep | awk '{print $5}'
}
cap gemm_tc_split 'gemm_tc_kernel' 102 17 $BENCH
cap whiten_split 'whiten_' 12 2 $BENCH
cap potrf_small_split 'potrf_inv_small' 18 3 $BENCH
cap stream_split 'rbf_bwd_finish|rbf_bwd_prep_rows|marginal_reduce_split|softmax_nll|scale_rows|sym_phi|marginal_bwd_prep|step_grad_finish|step_assemble' 84 14 $BENCH
cap gemm_tc2_split 'gemm_tc2_kernel' 12 2 $BENCH
cap gemm_tc2_scaled 'gemm_tc2_kernel' 22 3 python bench.py --workload scaled --batch 8192 --steps 1 --warmup 3 --no-cpu-baseline
cap stream_scaled 'marginal_reduce|marginal_bwd_prep|rbf_bwd_prep|scale_rows|sym_phi|softmax_nll' 9 9 python bench.py --workload scaled --batch 16384 --steps 1 --warmup 3 --no-cpu-baseline
rm -f $OUT/${TAG}_stream_scaled.ncu-rep $OUT/${TAG}_gemm_tc2_scaled.ncu-rep $OUT/${TAG}_stream_split.ncu-rep
ls -la $OUT/${TAG}_* | awk '{print $5, $9}'
