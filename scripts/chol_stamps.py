#!/usr/bin/env python
"""Phase timeline of the cluster Cholesky (potrf_cluster.cu) from its clock64 stamps: matrix 0, per CTA rank and block step.
    python scripts/chol_stamps.py [n]"""
import os, sys, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vargp_b200 import ops as vops

ops = vops.get_ops()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
g = torch.Generator().manual_seed(n)
X = torch.randn(30, n, n + 5, generator=g, dtype=torch.float64)
A = (X @ X.transpose(-1, -2) / (n + 5) + 0.05 * torch.eye(n, dtype=torch.float64)).float().cuda()
L, W = torch.empty_like(A), torch.empty_like(A)
info = torch.zeros(30, device='cuda', dtype=torch.int32)
buf = torch.zeros(4 * 256, device='cuda', dtype=torch.int64)
ops.lib.vargp_chol_cluster_debug.argtypes = [ctypes.c_void_p]
for _ in range(3):
  ops.chol_inv_cluster(A, L, W, 1e-4, info)
torch.cuda.synchronize()
ops.lib.vargp_chol_cluster_debug(buf.data_ptr())
ops.chol_inv_cluster(A, L, W, 1e-4, info)
torch.cuda.synchronize()
ops.lib.vargp_chol_cluster_debug(None)
b = buf.cpu().view(4, 16, 16)
nblk = (n + 31) // 32
GHZ = 1.965
for r in range(4):
  t0 = int(b[r, 0, 15]) or int(b[r, 0, 0])
  if t0 == 0:
    continue
  print(f'rank {r}: (us relative to the CTA start; w0 / w1 = warps 0 and 1)')
  for k in range(nblk):
    f = lambda s: (int(b[r, k, s]) - t0) / GHZ / 1e3 if int(b[r, k, s]) else float('nan')
    print(f'  k={k}: sync1 {f(0):7.2f} | B done w0 {f(1):7.2f} w1 {f(5):7.2f} | sync2 {f(2):7.2f} | C done w0 {f(3):7.2f} w1 {f(7):7.2f}'
          f' | diag start {f(8):7.2f} factored {f(9):7.2f} inverted {f(10):7.2f} pushed {f(11):7.2f}'
          f' | B(w0): item start {f(12):7.2f} product {f(13):7.2f} staged {f(14):7.2f}')
  print(f'  final sync {(int(b[r, 15, 14]) - t0) / GHZ / 1e3:7.2f}  end {(int(b[r, 15, 15]) - t0) / GHZ / 1e3:7.2f}')
