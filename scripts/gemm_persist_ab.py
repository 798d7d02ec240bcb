#!/usr/bin/env python
"""One-CTA tcgen05 GEMM, one tile per CTA (gemm_tc.cu) vs the persistent form (gemm_tcp.cu), at the Split-MNIST call sites.
    VARGP_TC_PERSIST=0|2 python scripts/gemm_persist_ab.py"""
import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vargp_b200 import ops as vops
ops = vops.get_ops()
def t(fn, it=30):
  for _ in range(5): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(it): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / it * 1e3
res = {}
for (name, M, N, K, batch, kw) in (('V=W*Kzx', 300, 512, 300, 30, dict(a_tri='lower', zeroed=True)), ('NV=N*V', 300, 512, 300, 30, {}),
                                   ('G=Vg*Vt', 300, 300, 512, 30, dict(c_tri='lower')), ('Y=Xi*W', 300, 300, 300, 30, dict(b_tri='lower', zeroed=True)),
                                   ('N=eps*W*Wt', 300, 300, 300, 30, dict(a_tri='lower', b_tri='upper', c_tri='lower', zeroed=True)),
                                   ('Wbar+=2eps*G*W', 300, 300, 300, 30, dict(b_tri='lower', c_tri='lower', beta=1., zeroed=True)),
                                   ('Gz2', 300, 784, 300, 30, {}), ('one wave', 128, 128, 320, 148, {}), ('two waves', 128, 128, 320, 296, {}),
                                   ('2.43 waves', 128, 128, 320, 360, {})):
  A = torch.randn(batch, M, K, device='cuda'); B = torch.randn(batch, K, N, device='cuda'); C = torch.empty(batch, M, N, device='cuda')
  if kw.get('a_tri') == 'lower': A = A.tril()
  if kw.get('b_tri') == 'lower': B = B.tril()
  if kw.get('b_tri') == 'upper': B = B.triu()
  res[name] = round(t(lambda: ops.gemm(A, B, C, **kw)), 2)
print(json.dumps(dict(persist=os.environ.get('VARGP_TC_PERSIST', 'default'), us=res)))

import ctypes
buf = torch.zeros(8, dtype=torch.int64, device='cuda')
ops.lib.vargp_tc_debug.argtypes = [ctypes.c_void_p]
ops.lib.vargp_tc_debug(buf.data_ptr())
for (M, N, K, batch) in ((128, 128, 320, 148), (128, 128, 320, 360), (300, 512, 300, 30)):
  A = torch.randn(batch, M, K, device='cuda'); B = torch.randn(batch, K, N, device='cuda'); C = torch.empty(batch, M, N, device='cuda')
  for _ in range(3): ops.gemm(A, B, C)
  torch.cuda.synchronize()
  st = buf.cpu().tolist()
  names = ['entry', 'setup', 'slab0 landed', 'slab0 issued', 'sum0 ready', 'tile0 drained', 'all stores issued', 'exit']
  print((M, N, K, batch), ' '.join(f'{n}=+{(s - st[0]) / 1.965e3:.2f}us' for n, s in zip(names, st)))
ops.lib.vargp_tc_debug(None)
