#!/usr/bin/env bash
set -uo pipefail
TAG="${1:-r4g}"; OUT=gpurun_out; mkdir -p $OUT
VARGP_TC_PERSIST=2 timeout 300 python -m pytest tests/test_gemm_tc_gpu.py -x -q -m gpu > $OUT/${TAG}_gemm_pytest_p2.log 2>&1
echo "gemm pytest (persist=2) rc $?"; tail -2 $OUT/${TAG}_gemm_pytest_p2.log
VARGP_TC_PERSIST=1 timeout 400 python -m pytest tests/test_model_gpu.py tests/test_large_gpu.py tests/test_train_gpu.py -x -q -m gpu > $OUT/${TAG}_model_pytest.log 2>&1
echo "model pytest (persist=1) rc $?"; tail -2 $OUT/${TAG}_model_pytest.log
for m in 0 1; do VARGP_TC_PERSIST=$m timeout 120 python scripts/gemm_persist_ab.py 2>&1 | grep persist; done | tee $OUT/${TAG}_gemm_persist_ab.txt
bash scripts/r2_ab.sh $TAG p0:VARGP_TC_PERSIST=0 p1:VARGP_TC_PERSIST=1 p3:VARGP_TC_PERSIST=3 p0b:VARGP_TC_PERSIST=0 p1b:VARGP_TC_PERSIST=1
AB_ARGS="--workload permuted_mnist --steps 30 --warmup 5" bash scripts/r2_ab.sh $TAG perm_p0:VARGP_TC_PERSIST=0 perm_p1:VARGP_TC_PERSIST=1
