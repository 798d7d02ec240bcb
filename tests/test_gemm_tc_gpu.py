"""tcgen05 / TMA 3xTF32 GEMM (vargp_gemm_tc) against an fp64 product: all operand majors, ragged shapes,
batch broadcasting, triangular k-skipping, output masks, alpha/beta and the fused RBF epilogue.
The 3xTF32 split must hold fp32-grade accuracy (norm-relative error < 2e-6)."""
import math

import pytest
import torch

from tests.emu_ops import EmuOps

pytestmark = pytest.mark.gpu
EMU = EmuOps()


def rnd(*s, seed=0):
  g = torch.Generator().manual_seed(seed + sum(s))
  return torch.randn(*s, generator=g, dtype=torch.float64)


def relerr(a, b):
  a, b = a.double().cpu(), b.double().cpu()
  return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def make(b, M, N, K, ta, tb):
  A = rnd(*b, K, M, seed=1).transpose(-1, -2) if ta else rnd(*b, M, K, seed=1)
  B = rnd(*b, N, K, seed=2).transpose(-1, -2) if tb else rnd(*b, K, N, seed=2)
  to = lambda t, tr: (t.transpose(-1, -2).contiguous().to('cuda', torch.float32).transpose(-1, -2) if tr
                      else t.to('cuda', torch.float32))
  return A, B, to(A, ta), to(B, tb)


SHAPES = [(128, 128, 32), (128, 128, 96), (256, 384, 128), (60, 512, 60), (300, 512, 300), (300, 300, 512),
          (132, 68, 44), (3000, 512, 784), (64, 64, 784), (1000, 1000, 1000)]


@pytest.mark.parametrize('M,N,K', SHAPES)
@pytest.mark.parametrize('ta,tb', [(False, True), (False, False), (True, False), (True, True)])
def test_tc_gemm_majors(cuda_ops, M, N, K, ta, tb):
  """(ta, tb) = (False, True): both K-major ("NT"); (False, False): B N-major; (True, False): A M-major, B N-major."""
  A, B, Ad, Bd = make((2,), M, N, K, ta, tb)
  C64 = A @ B
  Cd = torch.full((2, M, N), float('nan'), device='cuda')
  n0 = cuda_ops.tc_calls
  cuda_ops.gemm(Ad, Bd, Cd)
  assert cuda_ops.tc_calls == n0 + 1, 'tensor-core path was not taken'
  err = relerr(Cd, C64)
  assert err < 1e-6, f'{(M, N, K, ta, tb)}: {err:.3e}'


def test_tc_falls_back_when_unalignable(cuda_ops):
  A, B, Ad, Bd = make((2,), 130, 70, 45, False, True)      # K = 45: row stride not a multiple of 16 B
  Cd = torch.empty(2, 130, 70, device='cuda')
  n0 = cuda_ops.tc_calls
  cuda_ops.gemm(Ad, Bd, Cd)
  assert cuda_ops.tc_calls == n0
  assert relerr(Cd, A @ B) < 2e-6


@pytest.mark.parametrize('kw', [
  dict(a_tri='lower'), dict(a_tri='upper', ta=True), dict(b_tri='lower'), dict(a_tri='lower', b_tri='upper', tb=True, c_tri='lower'),
  dict(c_tri='lower', beta=1.0, tb=True), dict(c_tri='upper', alpha=-0.5), dict(beta=0.5, alpha=2.0)])
def test_tc_gemm_flags(cuda_ops, kw):
  kw = dict(kw)
  ta, tb = kw.pop('ta', False), kw.pop('tb', False)
  n = 300
  A, B, Ad, Bd = make((3, 2), n, n, n, ta, tb)
  # physically zero the declared triangles (contract of zeroed=True)
  for t64, td, tri in ((A, Ad, kw.get('a_tri')), (B, Bd, kw.get('b_tri'))):
    if tri:
      mask = torch.ones(n, n).tril(-1).bool() if tri == 'upper' else torch.ones(n, n).triu(1).bool()
      t64.masked_fill_(mask, 0.0)
      td.masked_fill_(mask.cuda(), 0.0)
  C0 = rnd(3, 2, n, n, seed=3)
  C64 = C0.clone()
  EMU.gemm(A, B, C64, **kw)
  Cd = C0.to('cuda', torch.float32)
  n0 = cuda_ops.tc_calls
  cuda_ops.gemm(Ad, Bd, Cd, zeroed=True, **kw)
  assert cuda_ops.tc_calls == n0 + 1
  assert relerr(Cd, C64) < 1e-6, kw


def test_tc_batch_broadcast_and_block_views(cuda_ops):
  H, C, S, M, B = 2, 3, 4, 64, 256
  P = S * M
  W = rnd(H, C, P, P, seed=5).tril()
  V = rnd(H, C, P, B, seed=7)
  Wd, Vd = W.to('cuda', torch.float32), V.to('cuda', torch.float32)
  X = rnd(1, 1, B, 96, seed=8)             # broadcast over both batch dims
  out = torch.empty(H, C, P, 96, device='cuda')
  n0 = cuda_ops.tc_calls
  cuda_ops.gemm(Vd, X.to('cuda', torch.float32), out)
  assert cuda_ops.tc_calls == n0 + 1
  assert relerr(out, V @ X) < 1e-6
  # diagonal-block views (3 batch dims, strided blocks)
  blocks = lambda t: t.as_strided((H, C, S, M, M), (C * P * P, P * P, M * P + M, P, 1))
  rows = lambda t: t.as_strided((H, C, S, M, B), (C * P * B, P * B, M * B, B, 1))
  TV64 = torch.empty(H, C, P, B, dtype=torch.float64)
  EMU.gemm(blocks(W).transpose(-1, -2), rows(V), rows(TV64), a_tri='upper')
  TVd = torch.empty(H, C, P, B, device='cuda')
  cuda_ops.gemm(blocks(Wd).transpose(-1, -2), rows(Vd), rows(TVd), a_tri='upper', zeroed=True)
  assert cuda_ops.tc_calls == n0 + 2
  assert relerr(TVd, TV64) < 1e-6


@pytest.mark.parametrize('H,C,P,B,D', [(3, 10, 300, 512, 784), (2, 3, 64, 100, 64)])
def test_tc_rbf_epilogue(cuda_ops, H, C, P, B, D):
  theta = 0.1 * rnd(H, D + 1, seed=1) + math.log(math.sqrt(D) / 3)
  zs, xs = torch.rand(H, C, P, D, dtype=torch.float64) / 8, torch.rand(H, 1, B, D, dtype=torch.float64) / 8
  zn, xn = (zs * zs).sum(-1), (xs * xs).sum(-1)
  K64, Kzz64 = torch.empty(H, C, P, B, dtype=torch.float64), torch.empty(H, C, P, P, dtype=torch.float64)
  EMU.rbf_gram(zs, zn, xs, xn, theta, K64, False)
  EMU.rbf_gram(zs, zn, zs, zn, theta, Kzz64, True)
  f = lambda t: t.to('cuda', torch.float32)
  Kd, Kzzd = torch.empty(H, C, P, B, device='cuda'), torch.empty(H, C, P, P, device='cuda')
  n0 = cuda_ops.tc_calls
  cuda_ops.rbf_gram(f(zs), f(zn), f(xs), f(xn), f(theta), Kd, False)
  cuda_ops.rbf_gram(f(zs), f(zn), f(zs), f(zn), f(theta), Kzzd, True)
  assert cuda_ops.tc_calls == n0 + 2
  assert relerr(Kd, K64) < 1e-6 and relerr(Kzzd, Kzz64) < 1e-6


# ---------------------------------------------------------------------------------------------------------------
# persistent form of the 1-CTA kernel (gemm_tcp.cu): forced for every TMA-store launch through vargp_tc_persist_config(2)
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture
def force_persist(cuda_ops):
  old = cuda_ops.tc_persist_config(2)
  yield cuda_ops
  cuda_ops.tc_persist_config(old)


@pytest.mark.parametrize('M,N,K', [(128, 128, 96), (132, 68, 44), (300, 512, 300), (300, 784, 300), (1000, 1000, 320)])
@pytest.mark.parametrize('ta,tb', [(False, True), (False, False), (True, False), (True, True)])
def test_tcp_gemm_majors(force_persist, M, N, K, ta, tb):
  ops = force_persist
  A, B, Ad, Bd = make((7,), M, N, K, ta, tb)
  Cd = torch.full((7, M, N), float('nan'), device='cuda')
  n0 = ops.tc_calls
  ops.gemm(Ad, Bd, Cd)
  assert ops.tc_calls == n0 + 1
  assert relerr(Cd, A @ B) < 1e-6, (M, N, K, ta, tb)


@pytest.mark.parametrize('kw', [
  dict(a_tri='lower'), dict(a_tri='upper', ta=True), dict(b_tri='lower'), dict(a_tri='lower', b_tri='upper', tb=True, c_tri='lower'),
  dict(c_tri='lower', beta=1.0, tb=True), dict(c_tri='upper', alpha=-0.5)])
def test_tcp_gemm_flags_match_the_one_tile_form_bitwise(cuda_ops, kw):
  """Triangular operands give tiles of 2 ... 10 slabs (and culled ones): the work-stealing tile loop in longest-first order
  must produce exactly what one tile per CTA produces -- the same instruction sequence per partial sum."""
  kw = dict(kw)
  ta, tb = kw.pop('ta', False), kw.pop('tb', False)
  n = 300
  A, B, Ad, Bd = make((5, 6), n, n, n, ta, tb)
  for t64, td, tri in ((A, Ad, kw.get('a_tri')), (B, Bd, kw.get('b_tri'))):
    if tri:
      mask = torch.ones(n, n).tril(-1).bool() if tri == 'upper' else torch.ones(n, n).triu(1).bool()
      t64.masked_fill_(mask, 0.0)
      td.masked_fill_(mask.cuda(), 0.0)
  C0 = rnd(5, 6, n, n, seed=3)
  C64 = C0.clone()
  EMU.gemm(A, B, C64, **kw)
  outs = []
  for mode in (0, 2):
    old = cuda_ops.tc_persist_config(mode)
    try:
      Cd = C0.to('cuda', torch.float32)
      cuda_ops.gemm(Ad, Bd, Cd, zeroed=True, **kw)
      outs.append(Cd)
    finally:
      cuda_ops.tc_persist_config(old)
  assert relerr(outs[1], C64) < 1e-6, kw
  if kw.get('beta', 0.0) == 0.0:          # beta = 1 goes through cp.reduce adds at L2: same values, no ordering guarantee needed
    assert torch.equal(outs[0], outs[1])


def test_tcp_rbf_epilogue(force_persist):
  """Fused gamma^2 exp(.) epilogue (and the symmetric lower-triangular Kzz form with culled tiles) in the persistent loop."""
  ops = force_persist
  H, C, P, B, D = 3, 10, 300, 512, 784
  theta = 0.1 * rnd(H, D + 1, seed=1) + math.log(math.sqrt(D) / 3)
  zs, xs = torch.rand(H, C, P, D, dtype=torch.float64) / 8, torch.rand(H, 1, B, D, dtype=torch.float64) / 8
  zn, xn = (zs * zs).sum(-1), (xs * xs).sum(-1)
  K64, Kzz64 = torch.empty(H, C, P, B, dtype=torch.float64), torch.empty(H, C, P, P, dtype=torch.float64)
  EMU.rbf_gram(zs, zn, xs, xn, theta, K64, False)
  EMU.rbf_gram(zs, zn, zs, zn, theta, Kzz64, True)
  f = lambda t: t.to('cuda', torch.float32)
  Kd, Kzzd = torch.empty(H, C, P, B, device='cuda'), torch.full((H, C, P, P), float('nan'), device='cuda')
  ops.rbf_gram(f(zs), f(zn), f(xs), f(xn), f(theta), Kd, False)
  ops.rbf_gram(f(zs), f(zn), f(zs), f(zn), f(theta), Kzzd, True, c_tri='lower')
  assert relerr(Kd, K64) < 1e-6 and relerr(Kzzd, Kzz64.tril()) < 1e-6
  assert torch.equal(Kzzd.triu(1), torch.zeros_like(Kzzd))


# ---------------------------------------------------------------------------------------------------------------
# persistent 2-CTA kernel (gemm_tc2.cu): forced for every qualifying shape through vargp_tc2_config(1)
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture
def force_tc2(cuda_ops):
  old = cuda_ops.tc2_config(1)
  yield cuda_ops
  cuda_ops.tc2_config(old)


TC2_SHAPES = [(256, 256, 64), (512, 256, 96), (300, 520, 300), (1000, 1000, 1000), (3000, 512, 784), (260, 1100, 68),
              (2048, 768, 2048)]


@pytest.mark.parametrize('M,N,K', TC2_SHAPES)
@pytest.mark.parametrize('ta,tb', [(False, True), (False, False), (True, False), (True, True)])
def test_tc2_gemm_majors(force_tc2, M, N, K, ta, tb):
  ops = force_tc2
  A, B, Ad, Bd = make((2,), M, N, K, ta, tb)
  C64 = A @ B
  Cd = torch.full((2, M, N), float('nan'), device='cuda')
  n0 = ops.tc2_launch_count()
  ops.gemm(Ad, Bd, Cd)
  assert ops.tc2_launch_count() == n0 + 1, '2-CTA kernel was not taken'
  err = relerr(Cd, C64)
  assert err < 1e-6, f'{(M, N, K, ta, tb)}: {err:.3e}'


def test_tc2_positive_gram_has_no_truncation_drift(force_tc2):
  """All-positive operands (the D = 784 Gram of U[0,1] data) expose accumulation / split bias."""
  ops = force_tc2
  g = torch.Generator().manual_seed(3)
  A = torch.rand(2, 512, 784, generator=g, dtype=torch.float64)
  B = torch.rand(2, 784, 1024, generator=g, dtype=torch.float64)
  Ad, Bd = A.to('cuda', torch.float32), B.to('cuda', torch.float32).transpose(-1, -2).contiguous().transpose(-1, -2)
  Cd = torch.empty(2, 512, 1024, device='cuda')
  n0 = ops.tc2_launch_count()
  ops.gemm(Ad, Bd, Cd)
  assert ops.tc2_launch_count() == n0 + 1
  rel = ((Cd.double().cpu() - A @ B) / (A @ B)).abs().max().item()
  assert rel < 1e-6, rel


@pytest.mark.parametrize('kw', [
  dict(a_tri='lower'), dict(a_tri='upper', ta=True), dict(b_tri='lower'), dict(a_tri='lower', b_tri='upper', tb=True, c_tri='lower'),
  dict(c_tri='lower', beta=1.0, tb=True), dict(c_tri='upper', alpha=-0.5), dict(beta=0.5, alpha=2.0)])
@pytest.mark.parametrize('n', [300, 1000])
def test_tc2_gemm_flags(force_tc2, kw, n):
  ops = force_tc2
  kw = dict(kw)
  ta, tb = kw.pop('ta', False), kw.pop('tb', False)
  A, B, Ad, Bd = make((3, 2), n, n, n, ta, tb)
  for t64, td, tri in ((A, Ad, kw.get('a_tri')), (B, Bd, kw.get('b_tri'))):
    if tri:
      mask = torch.ones(n, n).tril(-1).bool() if tri == 'upper' else torch.ones(n, n).triu(1).bool()
      t64.masked_fill_(mask, 0.0)
      td.masked_fill_(mask.cuda(), 0.0)
  C0 = rnd(3, 2, n, n, seed=3)
  C64 = C0.clone()
  EMU.gemm(A, B, C64, **kw)
  Cd = C0.to('cuda', torch.float32)
  n0 = ops.tc2_launch_count()
  ops.gemm(Ad, Bd, Cd, zeroed=True, **kw)
  assert ops.tc2_launch_count() == n0 + 1
  assert relerr(Cd, C64) < 1e-6, kw


def test_tc2_transposed_output_and_unaligned_rows(force_tc2):
  """C written through a transposed view (c_rs == 1) and through rows that are not 16-byte aligned (scalar stores)."""
  ops = force_tc2
  A, B, Ad, Bd = make((2,), 384, 516, 128, False, True)
  C64 = A @ B
  Ct = torch.empty(2, 516, 384, device='cuda')
  n0 = ops.tc2_launch_count()
  ops.gemm(Ad, Bd, Ct.transpose(-1, -2))
  Cpad = torch.empty(2, 384, 517, device='cuda')
  ops.gemm(Ad, Bd, Cpad[..., :516])
  assert ops.tc2_launch_count() == n0 + 2
  assert relerr(Ct.transpose(-1, -2), C64) < 1e-6 and relerr(Cpad[..., :516], C64) < 1e-6


def test_tc2_batch_broadcast_and_block_views(force_tc2):
  ops = force_tc2
  H, C, S, M, B = 2, 3, 2, 256, 512
  P = S * M
  W = rnd(H, C, P, P, seed=5).tril()
  V = rnd(H, C, P, B, seed=7)
  Wd, Vd = W.to('cuda', torch.float32), V.to('cuda', torch.float32)
  X = rnd(1, 1, B, 256, seed=8)             # broadcast over both batch dims
  out = torch.empty(H, C, P, 256, device='cuda')
  n0 = ops.tc2_launch_count()
  ops.gemm(Vd, X.to('cuda', torch.float32), out)
  assert relerr(out, V @ X) < 1e-6
  blocks = lambda t: t.as_strided((H, C, S, M, M), (C * P * P, P * P, M * P + M, P, 1))
  rows = lambda t: t.as_strided((H, C, S, M, B), (C * P * B, P * B, M * B, B, 1))
  TV64 = torch.empty(H, C, P, B, dtype=torch.float64)
  EMU.gemm(blocks(W).transpose(-1, -2), rows(V), rows(TV64), a_tri='upper')
  TVd = torch.empty(H, C, P, B, device='cuda')
  ops.gemm(blocks(Wd).transpose(-1, -2), rows(Vd), rows(TVd), a_tri='upper', zeroed=True)
  assert ops.tc2_launch_count() == n0 + 2
  assert relerr(TVd, TV64) < 1e-6


def test_tc2_rbf_epilogue(force_tc2):
  ops = force_tc2
  H, C, P, B, D = 3, 4, 512, 1024, 784
  theta = 0.1 * rnd(H, D + 1, seed=1) + math.log(math.sqrt(D) / 3)
  zs, xs = torch.rand(H, C, P, D, dtype=torch.float64) / 8, torch.rand(H, 1, B, D, dtype=torch.float64) / 8
  zn, xn = (zs * zs).sum(-1), (xs * xs).sum(-1)
  K64, Kzz64 = torch.empty(H, C, P, B, dtype=torch.float64), torch.empty(H, C, P, P, dtype=torch.float64)
  EMU.rbf_gram(zs, zn, xs, xn, theta, K64, False)
  EMU.rbf_gram(zs, zn, zs, zn, theta, Kzz64, True)
  f = lambda t: t.to('cuda', torch.float32)
  Kd, Kzzd = torch.empty(H, C, P, B, device='cuda'), torch.empty(H, C, P, P, device='cuda')
  n0 = ops.tc2_launch_count()
  ops.rbf_gram(f(zs), f(zn), f(xs), f(xn), f(theta), Kd, False)
  ops.rbf_gram(f(zs), f(zn), f(zs), f(zn), f(theta), Kzzd, True)
  assert ops.tc2_launch_count() == n0 + 2
  assert relerr(Kd, K64) < 1e-6 and relerr(Kzzd, Kzz64) < 1e-6
  assert torch.equal(Kzzd.diagonal(dim1=-2, dim2=-1), torch.exp(2 * f(theta)[:, D]).view(H, 1, 1).expand(H, C, P))
