#!/usr/bin/env python
"""Cholesky + inverse microbenchmark (batch 30 = H*C of the Split / Permuted-MNIST shapes): routes of vargp_chol_inv.
    python scripts/bench_chol.py [n ...]"""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vargp_b200 import ops as vops

ops = vops.get_ops()
ns = [int(a) for a in sys.argv[1:]] or [60, 120, 128, 180, 240, 300, 320]
for n in ns:
  g = torch.Generator().manual_seed(n)
  X = torch.randn(30, n, n + 5, generator=g, dtype=torch.float64)
  A64 = X @ X.transpose(-1, -2) / (n + 5) + 0.05 * torch.eye(n, dtype=torch.float64)
  L64 = torch.linalg.cholesky(A64 + 1e-4 * torch.eye(n, dtype=torch.float64))
  A = A64.float().cuda()
  L, W = torch.empty_like(A), torch.empty_like(A)
  info = torch.zeros(30, device='cuda', dtype=torch.int32)
  res = {}
  routes = (('cluster', (33, 320), 0), ('blocked_or_small', (0, 0), 0))
  if n > 320:
    routes = (('blocked128', (0, 0), 128), ('blocked256_cluster', (33, 320), 256), ('blocked320_cluster', (33, 320), 320))
  for route, cl, blk in routes:
    old_cl = ops.chol_cluster_config(*cl)
    old_blk = ops.chol_config()
    if blk:
      ops.chol_config(blk, blk + 1)
    for _ in range(5):
      ops.chol_inv(A, L, W, 1e-4, info)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
      ops.chol_inv(A, L, W, 1e-4, info)
    e1.record()
    torch.cuda.synchronize()
    res[route] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)
    res[route + '_err'] = float(f'{((L.double().cpu() - L64).norm() / L64.norm()).item():.2e}')
    ops.chol_cluster_config(*old_cl)
    ops.chol_config(*old_blk)
  print(json.dumps(dict(n=n, batch=30, us=res)), flush=True)
