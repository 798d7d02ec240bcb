#!/usr/bin/env bash
# persistent 1-CTA GEMM (gemm_tcp.cu): parity of the GEMM suite with the persistent form forced on, model tests, A/B.
set -uo pipefail
TAG="${1:-r3t}"; OUT=gpurun_out; mkdir -p $OUT
VARGP_TC_PERSIST=2 timeout 300 python -m pytest tests/test_gemm_tc_gpu.py -x -q -m gpu > $OUT/${TAG}_gemm_pytest_p2.log 2>&1
echo "gemm pytest (persist=2) rc $?"; tail -3 $OUT/${TAG}_gemm_pytest_p2.log
timeout 300 python -m pytest tests/test_gemm_tc_gpu.py tests/test_kernels_gpu.py -x -q -m gpu > $OUT/${TAG}_gemm_pytest_p1.log 2>&1
echo "gemm+kernels pytest (default) rc $?"; tail -2 $OUT/${TAG}_gemm_pytest_p1.log
timeout 400 python -m pytest tests/test_model_gpu.py tests/test_large_gpu.py tests/test_train_gpu.py -x -q -m gpu > $OUT/${TAG}_model_pytest.log 2>&1
echo "model pytest rc $?"; tail -2 $OUT/${TAG}_model_pytest.log
bash scripts/r2_ab.sh $TAG p0:VARGP_TC_PERSIST=0 p1:VARGP_TC_PERSIST=1 p2:VARGP_TC_PERSIST=2
AB_ARGS="--workload permuted_mnist --steps 30 --warmup 5" bash scripts/r2_ab.sh $TAG perm_p0:VARGP_TC_PERSIST=0 perm_p1:VARGP_TC_PERSIST=1
