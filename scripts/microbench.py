#!/usr/bin/env python
"""Kernel microbench sweep (BASELINE.json configs[4]): Kxz build, batched Cholesky, triangular inverse and the
TRSM-replacement GEMM  V = W Kzx  for M in {64 .. 4096}, C = 10, H = 3 (30 matrices), against the measured
roofline.  One JSON line per (kernel, M) on stdout and appended to gpurun_out/microbench.jsonl.

    python scripts/microbench.py [--M 64 128 ...] [--B 65536] [--iters 5] [--stream]

Timing: CUDA events on the launch stream, 2 warm-ups, then `iters` launches with an L2 flush (256 MB memset)
before each when the working set is below 2x L2.  Peaks: MEASURED_PEAKS.json (3xTF32 peak = bf16 sustained / 6:
TF32 dense runs at half the bf16 rate and every product takes three MMAs).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vargp_b200 import ops as vops   # noqa: E402

H, C, D = 3, 10, 784


def peaks():
  p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    j = json.load(open(p))
    return j['hbm_gbs'], j.get('bf16_tflops_sustained', j['bf16_tflops']) / 6.0, j['bf16_tflops'] / 6.0, 'measured'
  return 6650.0, 1400.0 / 6.0, 1590.0 / 6.0, 'fallback'


_flush = None


def timeit(fn, iters, flush):
  global _flush
  if flush and _flush is None:
    _flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
  for _ in range(2):
    fn()
  torch.cuda.synchronize()
  ts = []
  for _ in range(iters):
    if flush:
      _flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
  ts.sort()
  return ts[len(ts) // 2], ts[0]


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--M', type=int, nargs='+', default=[64, 128, 256, 512, 1024, 2048, 4096])
  ap.add_argument('--B', type=int, default=65536)
  ap.add_argument('--iters', type=int, default=5)
  ap.add_argument('--only', nargs='+', default=None, help='subset of: kxz kzz chol trtri chol_inv trsm stream')
  ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'microbench.jsonl'))
  args = ap.parse_args()
  ops = vops.get_ops()
  hbm, tc_sus, tc_burst, src = peaks()
  dev = 'cuda'
  g = torch.Generator().manual_seed(0)
  os.makedirs(os.path.dirname(args.out), exist_ok=True)
  fout = open(args.out, 'a')
  want = lambda k: args.only is None or k in args.only

  def emit(kernel, M, ms, ms_min, flops=None, nbytes=None, **kw):
    rec = dict(kernel=kernel, M=M, B=args.B, H=H, C=C, D=D, ms=round(ms, 4), ms_min=round(ms_min, 4), peaks=src)
    if flops is not None:
      rec['tflops'] = round(flops / ms / 1e9, 2)
      rec['frac_3xtf32_burst'] = round(flops / ms / 1e9 / tc_burst, 4)
      rec['frac_3xtf32_sustained'] = round(flops / ms / 1e9 / tc_sus, 4)
    if nbytes is not None:
      rec['gbs'] = round(nbytes / ms / 1e6, 1)
      rec['frac_hbm'] = round(nbytes / ms / 1e6 / hbm, 4)
    rec.update(kw)
    line = json.dumps(rec)
    print(line, flush=True)
    fout.write(line + '\n')
    fout.flush()

  theta = torch.cat([torch.log(torch.tensor(10.)) + 0.05 * torch.randn(H, D, generator=g),
                     torch.full((H, 1), -0.7)], dim=1).to(dev)
  x = torch.rand(args.B, D, generator=g).to(dev)
  xs, xn = torch.empty(H, args.B, D, device=dev), torch.empty(H, args.B, device=dev)
  ops.scale_rows(x, theta, xs, xn)
  for M in args.M:
    P = M
    z = torch.rand(C * P, D, generator=g).to(dev)
    zs, zn = torch.empty(H, C * P, D, device=dev), torch.empty(H, C * P, device=dev)
    ops.scale_rows(z, theta, zs, zn)
    zs4, zn3 = zs.view(H, C, P, D), zn.view(H, C, P)
    Kzz = torch.empty(H, C, P, P, device=dev)
    small = H * C * P * P * 4 * 3 < (256 << 20)
    # ---- Kzz (symmetric Gram + exp epilogue)
    if want('kzz'):
      ms, mn = timeit(lambda: ops.rbf_gram(zs4, zn3, zs4, zn3, theta, Kzz, True), args.iters, small)
      emit('kzz_gram', M, ms, mn, flops=2.0 * H * C * P * P * D)
    else:
      ops.rbf_gram(zs4, zn3, zs4, zn3, theta, Kzz, True)
    # ---- Cholesky / triangular inverse
    L, W = torch.empty_like(Kzz), torch.empty_like(Kzz)
    info = torch.zeros(H * C, dtype=torch.int32, device=dev)
    if want('chol'):
      ms, mn = timeit(lambda: ops.chol(Kzz, L, 1e-4, info), args.iters, small)
      emit('chol', M, ms, mn, flops=H * C * P ** 3 / 3.0, nbytes=8.0 * H * C * P * P, info_max=int(info.max()))
    else:
      ops.chol(Kzz, L, 1e-4, info)
    if want('trtri'):
      ms, mn = timeit(lambda: ops.trtri(L, W), args.iters, small)
      emit('trtri', M, ms, mn, flops=H * C * P ** 3 / 3.0, nbytes=8.0 * H * C * P * P)
    else:
      ops.trtri(L, W)
    if want('chol_inv'):
      ms, mn = timeit(lambda: ops.chol_inv(Kzz, L, W, 1e-4, info), args.iters, small)
      emit('chol_inv', M, ms, mn, flops=2.0 * H * C * P ** 3 / 3.0, nbytes=12.0 * H * C * P * P, info_max=int(info.max()),
           block=ops.chol_config()[0], blocked=bool(P >= ops.chol_config()[1]))
    # accuracy spot check of the factorisation on one matrix (fp64 on device)
    if want('chol') or want('trtri') or want('chol_inv'):
      K0 = Kzz[0, 0].double() + 1e-4 * torch.eye(P, device=dev, dtype=torch.float64)
      L0, W0 = L[0, 0].double(), W[0, 0].double()
      e_l = ((L0 @ L0.T - K0).norm() / K0.norm()).item()
      e_w = ((W0 @ L0 - torch.eye(P, device=dev, dtype=torch.float64)).norm() / P ** 0.5).item()
      print(json.dumps(dict(check='factor', M=M, chol_resid=e_l, trtri_resid=e_w)), flush=True)
    # ---- Kxz build and the TRSM replacement at minibatch B
    if want('kxz') or want('trsm') or want('stream'):
      Kzx = torch.empty(H, C, P, args.B, device=dev)
      xs4, xn3 = xs.view(H, 1, args.B, D), xn.view(H, 1, args.B)
      ms, mn = timeit(lambda: ops.rbf_gram(zs4, zn3, xs4, xn3, theta, Kzx, False), args.iters, False)
      if want('kxz'):
        emit('kxz_gram', M, ms, mn, flops=2.0 * H * C * P * args.B * D,
             nbytes=4.0 * (H * args.B * D + H * C * P * D + H * C * P * args.B))
      if want('trsm'):
        V = torch.empty_like(Kzx)
        ms, mn = timeit(lambda: ops.gemm(W, Kzx, V, a_tri='lower', zeroed=True), args.iters, False)
        emit('trsm_gemm(V=W*Kzx)', M, ms, mn, flops=1.0 * H * C * P * P * args.B,
             nbytes=4.0 * (H * C * P * P + 2 * H * C * P * args.B))
        if want('stream'):
          nu = torch.randn(H, C, P, device=dev)
          f_mean, f_var = torch.empty(H, C, args.B, device=dev), torch.empty(H, C, args.B, device=dev)
          A = torch.empty_like(V)
          ops.gemm(W.transpose(-1, -2), V, A, a_tri='upper', zeroed=True)
          ms, mn = timeit(lambda: ops.marginal_reduce(V, A, nu, theta, f_mean, f_var), args.iters, False)
          emit('marginal_reduce', M, ms, mn, nbytes=4.0 * (2 * H * C * P * args.B + 2 * H * C * args.B))
          g_mean, g_var, th_bar = torch.randn_like(f_mean), torch.randn_like(f_var), torch.zeros_like(theta)
          Vg = torch.empty_like(V)
          ms, mn = timeit(lambda: ops.marginal_bwd_prep(V, A, nu, g_mean, g_var, theta, A, Vg, th_bar), args.iters, False)
          emit('marginal_bwd_prep', M, ms, mn, nbytes=4.0 * (4 * H * C * P * args.B + 2 * H * C * args.B))
          rs, cs = torch.empty(H, C, P, device=dev), torch.zeros(H, args.B, device=dev)
          ms, mn = timeit(lambda: ops.rbf_bwd_prep(Vg, Kzx, rs, cs), args.iters, False)
          emit('rbf_bwd_prep', M, ms, mn, nbytes=4.0 * 3 * H * C * P * args.B)
          del Vg
          del A
        del V
      del Kzx
    del Kzz, L, W, zs, zn
    torch.cuda.empty_cache()


if __name__ == '__main__':
  main()
