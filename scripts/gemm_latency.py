#!/usr/bin/env python
"""Fixed cost vs per-slab cost of the 1-CTA tcgen05 GEMM: one full wave (148 CTAs of one 128x128 tile), K swept."""
import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vargp_b200 import ops as vops
ops = vops.get_ops()
ops.tc2_config(-1)
def t(fn, it=20):
  for _ in range(3): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(it): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / it * 1e3
for batch in (148, 296):
  for K in (32, 64, 128, 320, 800, 1600):
    A = torch.randn(batch, 128, K, device='cuda'); B = torch.randn(batch, K, 128, device='cuda'); C = torch.empty(batch, 128, 128, device='cuda')
    us = t(lambda: ops.gemm(A, B, C))
    us_t = t(lambda: ops.gemm(A, B.transpose(-1, -2).contiguous().transpose(-1, -2), C))
    print(json.dumps(dict(batch=batch, K=K, us=round(us, 2), us_bkmajor_incl_copy=round(us_t, 2), slabs=K // 32)))
# empty kernel launch latency for reference
x = torch.empty(1, device='cuda')
print('fill us', t(lambda: x.zero_()))
# pipeline stamps of CTA 0 (clock64 at 1.965 GHz)
import ctypes
buf = torch.zeros(8, dtype=torch.int64, device='cuda')
ops.lib.vargp_tc_debug.argtypes = [ctypes.c_void_p]
ops.lib.vargp_tc_debug(buf.data_ptr())
for (M, N, K, batch) in ((128, 128, 320, 148), (300, 300, 300, 30), (300, 512, 784, 30)):
  A = torch.randn(batch, M, K, device='cuda'); B = torch.randn(batch, K, N, device='cuda'); C = torch.empty(batch, M, N, device='cuda')
  for _ in range(3): ops.gemm(A, B.transpose(-1, -2).contiguous().transpose(-1, -2), C)
  torch.cuda.synchronize()
  st = buf.cpu().tolist()
  names = ['entry', 'setup', 'slab0 landed', 'slab0 issued', 'sum0 ready', 'mma retired', 'stored', 'exit']
  print((M, N, K, batch), ' '.join(f'{n}=+{(s - st[0]) / 1.965e3:.2f}us' for n, s in zip(names, st)))
ops.lib.vargp_tc_debug(None)
