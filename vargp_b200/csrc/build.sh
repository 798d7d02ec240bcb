#!/usr/bin/env bash
# Build libvargp_sm100.so in-tree (sm_100a only).  Usage: vargp_b200/csrc/build.sh [-v]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libvargp_sm100.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
EXTRA=""
if [[ "${1:-}" == "-v" ]]; then EXTRA="-Xptxas -v"; fi
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC -shared $EXTRA \
  "$HERE/api.cu" "$HERE/gemm_simt.cu" "$HERE/gemm_tc.cu" "$HERE/gemm_tcp.cu" "$HERE/gemm_tc2.cu" "$HERE/rbf.cu" "$HERE/marginal.cu" \
  "$HERE/likelihood.cu" "$HERE/chol.cu" "$HERE/potrf_blocked.cu" "$HERE/potrf_small.cu" "$HERE/potrf_cluster.cu" "$HERE/optim.cu" "$HERE/hyper.cu" "$HERE/step.cu" "$HERE/whiten.cu" "$HERE/peer.cu" \
  -o "$OUT"
echo "built $OUT"
