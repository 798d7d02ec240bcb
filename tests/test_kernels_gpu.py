"""GPU parity of every kernel of libvargp_sm100.so against the contracts in tests/emu_ops.py
(fp64 evaluation of the same contract as arbiter)."""
import math

import pytest
import torch

from tests.emu_ops import EmuOps

pytestmark = pytest.mark.gpu
EMU = EmuOps()


def rnd(*s, seed=0, scale=1.0):
  g = torch.Generator().manual_seed(seed + sum(s))
  return (scale * torch.randn(*s, generator=g, dtype=torch.float64))


def dev(t):
  return t.to('cuda', torch.float32)


def close(a, b, tol, name=''):
  a, b = a.detach().double().cpu(), b.detach().double().cpu()
  err = ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
  assert math.isfinite(err) and err < tol, f'{name}: rel err {err:.3e} (tol {tol})'


SHAPES = [(2, 3, 60, 48, 784), (3, 10, 33, 17, 37), (1, 4, 20, 100, 2), (3, 2, 130, 200, 64)]


@pytest.mark.parametrize('H,C,P,B,D', SHAPES)
def test_scale_rows_and_gram(cuda_ops, H, C, P, B, D):
  theta = rnd(H, D + 1, scale=0.1) + math.log(math.sqrt(D) / 3)
  z, x = torch.rand(C * P, D, dtype=torch.float64), torch.rand(B, D, dtype=torch.float64)
  zs64, zn64 = torch.empty(H, C * P, D, dtype=torch.float64), torch.empty(H, C * P, dtype=torch.float64)
  xs64, xn64 = torch.empty(H, B, D, dtype=torch.float64), torch.empty(H, B, dtype=torch.float64)
  EMU.scale_rows(z, theta, zs64, zn64)
  EMU.scale_rows(x, theta, xs64, xn64)
  zs, zn = torch.empty(H, C * P, D, device='cuda'), torch.empty(H, C * P, device='cuda')
  xs, xn = torch.empty(H, B, D, device='cuda'), torch.empty(H, B, device='cuda')
  cuda_ops.scale_rows(dev(z), dev(theta), zs, zn)
  cuda_ops.scale_rows(dev(x), dev(theta), xs, xn)
  close(zs, zs64, 1e-6, 'zs'); close(zn, zn64, 1e-6, 'zn'); close(xs, xs64, 1e-6, 'xs'); close(xn, xn64, 1e-6, 'xn')
  Kzz64, Kzx64 = torch.empty(H, C, P, P, dtype=torch.float64), torch.empty(H, C, P, B, dtype=torch.float64)
  EMU.rbf_gram(zs64.view(H, C, P, D), zn64.view(H, C, P), zs64.view(H, C, P, D), zn64.view(H, C, P), theta, Kzz64, True)
  EMU.rbf_gram(zs64.view(H, C, P, D), zn64.view(H, C, P), xs64.view(H, 1, B, D), xn64.view(H, 1, B), theta, Kzx64, False)
  Kzz, Kzx = torch.empty(H, C, P, P, device='cuda'), torch.empty(H, C, P, B, device='cuda')
  cuda_ops.rbf_gram(zs.view(H, C, P, D), zn.view(H, C, P), zs.view(H, C, P, D), zn.view(H, C, P), dev(theta), Kzz, True)
  cuda_ops.rbf_gram(zs.view(H, C, P, D), zn.view(H, C, P), xs.view(H, 1, B, D), xn.view(H, 1, B), dev(theta), Kzx, False)
  close(Kzz, Kzz64, 2e-6, 'Kzz'); close(Kzx, Kzx64, 2e-6, 'Kzx')
  g2 = torch.exp(2 * dev(theta)[:, -1])
  assert torch.equal(Kzz.diagonal(dim1=-2, dim2=-1), g2.view(H, 1, 1).expand(H, C, P))


GEMM_CASES = [
  dict(b=(3, 10), M=60, N=512, K=60, a_tri='lower'),
  dict(b=(3, 10), M=60, N=512, K=60, a_tri='upper', ta=True),
  dict(b=(2, 3, 4), M=12, N=12, K=12, a_tri='lower', b_tri='lower'),
  dict(b=(2, 5), M=130, N=130, K=77, c_tri='lower', beta=1.0, tb=True),
  dict(b=(2, 5), M=130, N=130, K=130, a_tri='lower', b_tri='upper', c_tri='lower', tb=True),
  dict(b=(30,), M=300, N=784, K=512),
  dict(b=(4,), M=257, N=300, K=300, b_tri='lower', alpha=-1.0),
  dict(b=(2, 3), M=33, N=1, K=33, a_tri='upper', ta=True),
  dict(b=(3, 10), M=60, N=60, K=1, beta=1.0, c_tri='lower'),
  dict(b=(1,), M=1000, N=640, K=1000, a_tri='lower'),
]


@pytest.mark.parametrize('case', GEMM_CASES)
def test_gemm(cuda_ops, case):
  b, M, N, K = case['b'], case['M'], case['N'], case['K']
  A = rnd(*b, K, M, seed=1).transpose(-1, -2) if case.get('ta') else rnd(*b, M, K, seed=1)
  Bm = rnd(*b, N, K, seed=2).transpose(-1, -2) if case.get('tb') else rnd(*b, K, N, seed=2)
  C0 = rnd(*b, M, N, seed=3)
  kw = dict(alpha=case.get('alpha', 1.0), beta=case.get('beta', 0.0), a_tri=case.get('a_tri'),
            b_tri=case.get('b_tri'), c_tri=case.get('c_tri'))
  C64 = C0.clone()
  EMU.gemm(A, Bm, C64, **kw)
  Ad = dev(A.transpose(-1, -2).contiguous()).transpose(-1, -2) if case.get('ta') else dev(A)
  Bd = dev(Bm.transpose(-1, -2).contiguous()).transpose(-1, -2) if case.get('tb') else dev(Bm)
  Cd = dev(C0)
  # poison the structurally-zero triangle of A: the kernel must never read it
  if kw['a_tri'] == 'lower':
    Ad.masked_fill_(torch.ones(M, K, device='cuda').triu(1).bool(), float('nan'))
  elif kw['a_tri'] == 'upper':
    Ad.masked_fill_(torch.ones(M, K, device='cuda').tril(-1).bool(), float('nan'))
  if kw['b_tri'] == 'lower':
    Bd.masked_fill_(torch.ones(K, N, device='cuda').triu(1).bool(), float('nan'))
  elif kw['b_tri'] == 'upper':
    Bd.masked_fill_(torch.ones(K, N, device='cuda').tril(-1).bool(), float('nan'))
  cuda_ops.gemm(Ad, Bd, Cd, **kw)
  close(Cd, C64, 3e-6, f'gemm {case}')


def test_gemm_broadcast_and_views(cuda_ops):
  H, C, S, M, B = 2, 3, 4, 8, 50
  P = S * M
  W = rnd(H, C, P, P, seed=5).tril()
  Lu = rnd(S, C, M, M, seed=6).tril()
  T64 = torch.empty(H, C, S, M, M, dtype=torch.float64)
  Wd64 = W.as_strided((H, C, S, M, M), (C * P * P, P * P, M * P + M, P, 1))
  EMU.gemm(Wd64, Lu.permute(1, 0, 2, 3).unsqueeze(0), T64, a_tri='lower', b_tri='lower')
  Wg = dev(W)
  Wd = Wg.as_strided((H, C, S, M, M), (C * P * P, P * P, M * P + M, P, 1))
  T = torch.empty(H, C, S, M, M, device='cuda')
  cuda_ops.gemm(Wd, dev(Lu).permute(1, 0, 2, 3).unsqueeze(0), T, a_tri='lower', b_tri='lower')
  close(T, T64, 3e-6, 'T blocks')
  V = rnd(H, C, P, B, seed=7)
  TV64 = torch.empty(H, C, P, B, dtype=torch.float64)
  rows = lambda t: t.as_strided((H, C, S, M, B), (C * P * B, P * B, M * B, B, 1))
  EMU.gemm(T64.transpose(-1, -2), rows(V), rows(TV64), a_tri='upper')
  Vg, TVg = dev(V), torch.empty(H, C, P, B, device='cuda')
  cuda_ops.gemm(T.transpose(-1, -2), rows(Vg), rows(TVg), a_tri='upper')
  close(TVg, TV64, 3e-6, 'TV')


@pytest.mark.parametrize('n,batch', [(60, 30), (300, 30), (33, 7), (20, 4), (1, 3), (600, 4), (1000, 2)])
def test_chol_trtri(cuda_ops, n, batch):
  X = rnd(batch, n, n + 5, seed=n)
  A = X @ X.transpose(-1, -2) / (n + 5) + 0.05 * torch.eye(n, dtype=torch.float64)
  L64 = torch.linalg.cholesky(A + 1e-4 * torch.eye(n, dtype=torch.float64))
  L = torch.full((batch, n, n), float('nan'), device='cuda')
  info = torch.full((batch,), -1, device='cuda', dtype=torch.int32)
  cuda_ops.chol(dev(A), L, 1e-4, info)
  assert int(info.abs().max()) == 0
  close(L, L64, 2e-5, 'chol')
  assert torch.equal(L.triu(1), torch.zeros_like(L))
  W = torch.full((batch, n, n), float('nan'), device='cuda')
  cuda_ops.trtri(L, W)
  W64 = torch.linalg.solve_triangular(L.double().cpu(), torch.eye(n, dtype=torch.float64).expand(batch, n, n), upper=False)
  close(W, W64, 2e-5, 'trtri')
  assert torch.equal(W.triu(1), torch.zeros_like(W))


def test_chol_reports_non_pd(cuda_ops):
  A = torch.eye(40, dtype=torch.float64).repeat(3, 1, 1)
  A[1, 17, 17] = -1.0
  L = torch.empty(3, 40, 40, device='cuda')
  info = torch.zeros(3, device='cuda', dtype=torch.int32)
  cuda_ops.chol(dev(A), L, 1e-4, info)
  assert info.tolist() == [0, 18, 0]


@pytest.mark.parametrize('n,batch,nb', [(600, 4, 128), (1024, 3, 128), (450, 5, 64), (520, 2, 96), (2048, 2, 128),
                                        (300, 6, 0), (129, 3, 128), (1000, 2, 256), (300, 30, 128), (128, 5, 128),
                                        (60, 30, 128), (97, 3, 128), (1, 2, 128), (33, 4, 128), (20, 7, 128),
                                        (300, 30, 'cluster'), (33, 3, 'cluster'), (60, 30, 'cluster'), (64, 2, 'cluster'),
                                        (65, 2, 'cluster'), (96, 5, 'cluster'), (97, 4, 'cluster'), (120, 30, 'cluster'),
                                        (128, 7, 'cluster'), (129, 3, 'cluster'), (180, 30, 'cluster'), (200, 5, 'cluster'),
                                        (240, 30, 'cluster'), (257, 2, 'cluster'), (289, 40, 'cluster'), (320, 3, 'cluster'),
                                        (1000, 3, 'cluster256'), (600, 4, 'cluster256'), (2048, 2, 'cluster256'),
                                        (700, 2, 'cluster320')])
def test_chol_inv_blocked(cuda_ops, n, batch, nb):
  """vargp_chol_inv: GEMM-driven blocked factorisation + inverse (potrf_blocked.cu) incl. ragged block counts;
  nb = 0 is the small-matrix route through the one-CTA kernels."""
  old = cuda_ops.chol_config()
  cl = isinstance(nb, str) and nb.startswith('cluster')
  old_cl = cuda_ops.chol_cluster_config(*((33, 320) if cl else (0, 0)))   # potrf_cluster.cu only where asked
  try:
    if cl and len(nb) > 7:            # blocked driver with cluster-factored diagonal blocks of 256 / 320
      cuda_ops.chol_config(int(nb[7:]), int(nb[7:]) + 1)
    elif nb == 'cluster':
      assert cuda_ops.chol_cluster_wants(n)
    elif nb:
      cuda_ops.chol_config(nb, nb + 1)
    else:
      cuda_ops.chol_config(0, 1 << 20)
    X = rnd(batch, n, n + 5, seed=n)
    A = X @ X.transpose(-1, -2) / (n + 5) + 0.05 * torch.eye(n, dtype=torch.float64)
    L64 = torch.linalg.cholesky(A + 1e-4 * torch.eye(n, dtype=torch.float64))
    W64 = torch.linalg.solve_triangular(L64, torch.eye(n, dtype=torch.float64).expand(batch, n, n), upper=False)
    # strided views (ld > n) with NaN-poisoned surroundings: nothing outside the n x n blocks may be read or written
    Lbuf = torch.full((batch, n, n + 8), float('nan'), device='cuda')
    Wbuf = torch.full((batch, n + 4, n + 4), float('nan'), device='cuda')
    L, W = Lbuf[:, :, :n], Wbuf[:, :n, :n]
    info = torch.full((batch,), -1, device='cuda', dtype=torch.int32)
    Ad = dev(A)
    A0 = Ad.clone()
    cuda_ops.chol_inv(Ad, L, W, 1e-4, info)
    assert int(info.abs().max()) == 0
    assert torch.equal(Ad, A0)
    close(L, L64, 2e-5, 'chol')
    close(W, W64, 3e-5, 'inverse')
    assert torch.equal(L.triu(1), torch.zeros_like(L)) and torch.equal(W.triu(1), torch.zeros_like(W))
    assert torch.isnan(Lbuf[:, :, n:]).all() and torch.isnan(Wbuf[:, n:]).all() and torch.isnan(Wbuf[:, :, n:]).all()
    # residuals in fp64: L L^T = A + jitter I and W L = I
    Ld, Wd = L.double().cpu(), W.double().cpu()
    Aj = A + 1e-4 * torch.eye(n, dtype=torch.float64)
    assert ((Ld @ Ld.transpose(-1, -2) - Aj).norm() / Aj.norm()).item() < 2e-6
    assert ((Wd @ Ld - torch.eye(n, dtype=torch.float64)).norm() / n ** 0.5).item() < 2e-4
  finally:
    cuda_ops.chol_config(*old)
    cuda_ops.chol_cluster_config(*old_cl)


def test_chol_inv_cluster_reports_first_bad_pivot(cuda_ops):
  """potrf_cluster.cu: the first failing pivot over all CTAs of the cluster is reported, good matrices are untouched by bad ones."""
  n = 300
  A = torch.eye(n, dtype=torch.float64).repeat(4, 1, 1)
  A[1, 170, 170] = -1.0
  A[1, 250, 250] = -1.0
  A[2, 3, 3] = -2.0
  A[3, 299, 299] = -1.0
  L, W = torch.empty(4, n, n, device='cuda'), torch.empty(4, n, n, device='cuda')
  info = torch.full((4,), -7, device='cuda', dtype=torch.int32)
  old = cuda_ops.chol_cluster_config(33, 320)
  try:
    cuda_ops.chol_inv(dev(A), L, W, 1e-4, info)
  finally:
    cuda_ops.chol_cluster_config(*old)
  assert info.tolist() == [0, 171, 4, 300]
  assert torch.allclose(L[0], torch.eye(n, device='cuda') * (1 + 1e-4) ** 0.5)


def test_chol_inv_cluster_in_place(cuda_ops):
  """A may alias W (the blocked driver's calling convention for diagonal blocks)."""
  n, batch = 250, 6
  X = rnd(batch, n, n + 5, seed=77)
  A = X @ X.transpose(-1, -2) / (n + 5) + 0.05 * torch.eye(n, dtype=torch.float64)
  L64 = torch.linalg.cholesky(A + 1e-4 * torch.eye(n, dtype=torch.float64))
  W64 = torch.linalg.solve_triangular(L64, torch.eye(n, dtype=torch.float64).expand(batch, n, n), upper=False)
  W = dev(A)
  L = torch.empty_like(W)
  info = torch.zeros(batch, device='cuda', dtype=torch.int32)
  cuda_ops.chol_inv_cluster(W, L, W, 1e-4, info)
  assert int(info.abs().max()) == 0
  close(L, L64, 2e-5, 'chol')
  close(W, W64, 3e-5, 'inverse')


def test_chol_inv_blocked_reports_first_bad_pivot(cuda_ops):
  old = cuda_ops.chol_config()
  old_cl = cuda_ops.chol_cluster_config(0, 0)
  try:
    cuda_ops.chol_config(64, 65)
    n = 300
    A = torch.eye(n, dtype=torch.float64).repeat(3, 1, 1)
    A[1, 170, 170] = -1.0
    A[1, 250, 250] = -1.0
    A[2, 3, 3] = -2.0
    L, W = torch.empty(3, n, n, device='cuda'), torch.empty(3, n, n, device='cuda')
    info = torch.zeros(3, device='cuda', dtype=torch.int32)
    cuda_ops.chol_inv(dev(A), L, W, 1e-4, info)
    assert info.tolist() == [0, 171, 4]
  finally:
    cuda_ops.chol_config(*old)
    cuda_ops.chol_cluster_config(*old_cl)


def test_chol_strided_block_view(cuda_ops):
  """chol / trtri on diagonal blocks of a bigger matrix (ld > n), as the blocked large-P schedule uses them."""
  n, P = 64, 256
  X = rnd(2, P, P + 3, seed=9)
  A = X @ X.transpose(-1, -2) / P + 0.1 * torch.eye(P, dtype=torch.float64)
  Ad = dev(A)
  blk = Ad[:, 64:128, 64:128]
  L = torch.zeros(2, P, P, device='cuda')
  info = torch.zeros(2, device='cuda', dtype=torch.int32)
  cuda_ops.chol(blk, L[:, 64:128, 64:128], 0.0, info)
  close(L[:, 64:128, 64:128], torch.linalg.cholesky(A[:, 64:128, 64:128]), 2e-5, 'blk chol')
  assert float(L[:, :64].abs().max()) == 0.0


@pytest.mark.parametrize('C,M', [(10, 60), (4, 20), (3, 7), (2, 1)])
def test_tril_unpack(cuda_ops, C, M):
  T = M * (M + 1) // 2
  vec = rnd(C, T, seed=M)
  vec[0, 0] = 30.0   # softplus threshold branch
  out64 = torch.empty(C, M, M, dtype=torch.float64)
  EMU.tril_unpack(vec, out64)
  out = torch.empty(C, M, M, device='cuda')
  cuda_ops.tril_unpack(dev(vec), out)
  close(out, out64, 1e-6, 'tril_unpack')
  g = rnd(C, M, M, seed=M + 1)
  vb64 = torch.empty(C, T, dtype=torch.float64)
  EMU.tril_unpack_bwd(g, vec, vb64)
  vb = torch.empty(C, T, device='cuda')
  cuda_ops.tril_unpack_bwd(dev(g), dev(vec), vb)
  close(vb, vb64, 1e-6, 'tril_unpack_bwd')


@pytest.mark.parametrize('H,C,S,M,B,D', [(3, 10, 3, 20, 200, 784), (2, 3, 1, 7, 33, 37), (1, 4, 2, 9, 21, 16),
                                         (2, 2, 1, 300, 40, 16), (3, 10, 5, 60, 512, 784), (1, 2, 2, 100, 5000, 16)])
def test_marginal_kl_kernels(cuda_ops, H, C, S, M, B, D):
  P = S * M
  W = (rnd(H, C, P, P, seed=1, scale=0.1).tril() + torch.eye(P, dtype=torch.float64))
  T, nu = rnd(H, C, S, M, M, seed=2).tril(), rnd(H, C, P, seed=3)
  Lu = rnd(C, M, M, seed=4, scale=0.1).tril() + torch.eye(M, dtype=torch.float64)
  theta = rnd(H, D + 1, seed=5, scale=0.1)
  V, NV = rnd(H, C, P, B, seed=6), rnd(H, C, P, B, seed=7)
  kl64 = torch.zeros((), dtype=torch.float64)
  EMU.kl_fwd(W, T, nu, Lu, M, kl64)
  kl = torch.zeros((), device='cuda')
  cuda_ops.kl_fwd(dev(W), dev(T), dev(nu), dev(Lu), M, kl)
  close(kl, kl64, 1e-5, 'kl_fwd')
  fm64, fv64 = torch.empty(H, C, B, dtype=torch.float64), torch.empty(H, C, B, dtype=torch.float64)
  EMU.marginal_reduce(V, NV, nu, theta, fm64, fv64)
  fm, fv = torch.empty(H, C, B, device='cuda'), torch.empty(H, C, B, device='cuda')
  cuda_ops.marginal_reduce(dev(V), dev(NV), dev(nu), dev(theta), fm, fv)
  close(fm, fm64, 1e-5, 'f_mean'); close(fv, fv64, 1e-5, 'f_var')
  # adjoint helpers
  g_kl = torch.tensor([0.7], dtype=torch.float64)
  Wb64, Tb64, nb64 = rnd(H, C, P, P, seed=9), rnd(H, C, S, M, M, seed=10), rnd(H, C, P, seed=11)
  Wb, Tb, nb = dev(Wb64), dev(Tb64), dev(nb64)
  EMU.kl_bwd(W, T, nu, M, g_kl[0], Wb64, Tb64, nb64)
  cuda_ops.kl_bwd(dev(W), dev(T), dev(nu), M, dev(g_kl), Wb, Tb, nb)
  close(Wb, Wb64, 1e-6, 'kl_bwd W'); close(Tb, Tb64, 1e-6, 'kl_bwd T'); close(nb, nb64, 1e-6, 'kl_bwd nu')
  Lb64 = rnd(C, M, M, seed=12)
  Lb = dev(Lb64)
  EMU.kl_bwd_lu(Lu, g_kl[0], Lb64)
  cuda_ops.kl_bwd_lu(dev(Lu), dev(g_kl), Lb)
  close(Lb, Lb64, 1e-6, 'kl_bwd_lu')
  gm, gv = rnd(H, C, B, seed=13), rnd(H, C, B, seed=14)
  Vb64, Vg64, thb64 = torch.empty(H, C, P, B, dtype=torch.float64), torch.empty(H, C, P, B, dtype=torch.float64), rnd(H, D + 1, seed=15)
  thb = dev(thb64)
  EMU.marginal_bwd_prep(V, NV, nu, gm, gv, theta, Vb64, Vg64, thb64)
  NVd, Vg = dev(NV), torch.empty(H, C, P, B, device='cuda')
  cuda_ops.marginal_bwd_prep(dev(V), NVd, dev(nu), dev(gm), dev(gv), dev(theta), NVd, Vg, thb)     # Vbar aliases NV
  close(NVd, Vb64, 1e-6, 'Vbar'); close(Vg, Vg64, 1e-6, 'Vg'); close(thb, thb64, 1e-5, 'thbar')
  X64 = rnd(H, C, P, P, seed=16)
  Xd = dev(X64)
  EMU.sym_phi(X64)
  cuda_ops.sym_phi(Xd)
  close(Xd, X64, 1e-6, 'sym_phi')
  EMU.sym_phi(X64, mirror=True)
  cuda_ops.sym_phi(Xd, mirror=True)
  close(Xd, X64, 1e-6, 'sym_mirror')


@pytest.mark.parametrize('H,C,P,B,D', [(3, 10, 60, 200, 784), (2, 3, 21, 33, 37), (1, 4, 18, 21, 2), (3, 10, 300, 512, 784),
                                       (5, 2, 70, 1028, 20), (2, 2, 44, 1100, 8)])
def test_rbf_adjoint_kernels(cuda_ops, H, C, P, B, D):
  theta = rnd(H, D + 1, seed=1, scale=0.1)
  zs, xs = rnd(H, C, P, D, seed=2), rnd(H, B, D, seed=3)
  K, Kbar = rnd(H, C, P, B, seed=4), rnd(H, C, P, B, seed=5)
  r64, c64 = torch.empty(H, C, P, dtype=torch.float64), rnd(H, B, seed=6)
  cs = dev(c64)
  Kb2 = Kbar.clone()
  EMU.rbf_bwd_prep(Kb2, K, r64, c64)
  Kbd, r = dev(Kbar), torch.empty(H, C, P, device='cuda')
  cuda_ops.rbf_bwd_prep(Kbd, dev(K), r, cs)
  close(Kbd, Kb2, 1e-6, 'Wk'); close(r, r64, 1e-5, 'rsum'); close(cs, c64, 1e-5, 'csum')
  # symmetric variant: diagonal moved out
  Ks, Ksb = rnd(H, C, P, P, seed=20), rnd(H, C, P, P, seed=21)
  rs64, ds64 = torch.empty(H, C, P, dtype=torch.float64), torch.empty(H, C, P, dtype=torch.float64)
  Ksb2 = Ksb.clone()
  EMU.rbf_bwd_prep(Ksb2, Ks, rs64, None, ds64)
  Ksbd, rs, ds = dev(Ksb), torch.empty(H, C, P, device='cuda'), torch.empty(H, C, P, device='cuda')
  cuda_ops.rbf_bwd_prep(Ksbd, dev(Ks), rs, None, ds)
  close(Ksbd, Ksb2, 1e-6, 'Wk sym'); close(rs, rs64, 1e-5, 'rsum sym'); close(ds, ds64, 1e-6, 'dsum')
  assert float(Ksbd.diagonal(dim1=-2, dim2=-1).abs().max()) == 0.0
  G1, G2, r1, r2 = rnd(H, C, P, D, seed=7), rnd(H, C, P, D, seed=8), rnd(H, C, P, seed=9), rnd(H, C, P, seed=10)
  for use1, use2 in ((True, True), (False, True), (True, False)):
    Zb64, thb64 = torch.empty(C, P, D, dtype=torch.float64), rnd(H, D + 1, seed=11)
    thb, Zb = dev(thb64), torch.empty(C, P, D, device='cuda')
    dg = rnd(H, C, P, seed=30) if use2 else None
    EMU.rbf_bwd_finish(zs, G1 if use1 else None, G2 if use2 else None, r1, r2, theta, Zb64, thb64, dg)
    cuda_ops.rbf_bwd_finish(dev(zs), dev(G1) if use1 else None, dev(G2) if use2 else None,
                            dev(r1), dev(r2), dev(theta), Zb, thb, None if dg is None else dev(dg))
    close(Zb, Zb64, 1e-5, 'Zbar'); close(thb, thb64, 1e-5, 'theta_bar finish')
  Gx = rnd(H, C, B, D, seed=12)
  for want_x in (False, True):
    thb64, xb64 = rnd(H, D + 1, seed=13), torch.empty(B, D, dtype=torch.float64)
    thb, xb = dev(thb64), torch.empty(B, D, device='cuda')
    EMU.rbf_bwd_xside(xs, c64, Gx if want_x else None, theta, thb64, xb64 if want_x else None)
    cuda_ops.rbf_bwd_xside(dev(xs), dev(c64), dev(Gx) if want_x else None, dev(theta), thb, xb if want_x else None)
    close(thb, thb64, 1e-5, 'theta_bar xside')
    if want_x:
      close(xb, xb64, 1e-5, 'x_bar')


@pytest.mark.parametrize('H,F,C,B', [(3, 10, 10, 512), (2, 5, 3, 33), (1, 7, 4, 100), (20, 50, 10, 64), (2, 3, 17, 40)])
def test_likelihood_kernels(cuda_ops, H, F, C, B):
  fm, fv = rnd(H, C, B, seed=1), rnd(H, C, B, seed=2).abs() + 0.05
  eps = rnd(H, F, C, B, seed=3)
  y = torch.randint(0, C, (B,), generator=torch.Generator().manual_seed(4))
  nll64, gm64, gv64 = torch.zeros((), dtype=torch.float64), torch.empty_like(fm), torch.empty_like(fv)
  EMU.nll_fwd_bwd(fm, fv, eps, y, nll64, gm64, gv64)
  nll, gm, gv = torch.zeros((), device='cuda'), torch.empty(H, C, B, device='cuda'), torch.empty(H, C, B, device='cuda')
  cuda_ops.nll_fwd_bwd(dev(fm), dev(fv), dev(eps), y.cuda(), nll, gm, gv)
  close(nll, nll64, 1e-5, 'nll'); close(gm, gm64, 1e-5, 'g_mean'); close(gv, gv64, 1e-5, 'g_var')
  p64 = torch.empty(B, C, dtype=torch.float64)
  EMU.predict(fm, fv, eps, p64)
  p = torch.empty(B, C, device='cuda')
  cuda_ops.predict(dev(fm), dev(fv), dev(eps), p)
  assert (p.double().cpu() - p64).abs().max().item() < 1e-6


@pytest.mark.parametrize('H,C,M,S,D', [(3, 10, 60, 5, 784), (2, 3, 7, 1, 37), (1, 4, 9, 2, 2)])
def test_step_prologue_epilogue_kernels(cuda_ops, H, C, M, S, D):
  """ops.step_assemble / ops.step_grad_finish (the parameter plumbing of the fused training step) vs their contracts."""
  P, T = S * M, M * (M + 1) // 2
  z, um, ut = rnd(C, M, D, seed=1), rnd(C, M, 1, seed=2), rnd(C, T, seed=3)
  Zc64, ml64, Ll64 = rnd(C, P, D, seed=4), torch.empty(C, M, dtype=torch.float64), torch.empty(C, M, M, dtype=torch.float64)
  Zc, ml, Ll = dev(Zc64), torch.empty(C, M, device='cuda'), torch.empty(C, M, M, device='cuda')
  EMU.step_assemble(z, um, ut, Zc64, ml64, Ll64)
  cuda_ops.step_assemble(dev(z), dev(um), dev(ut), Zc, ml, Ll)
  close(Zc, Zc64, 1e-7, 'Zcat'); close(ml, ml64, 1e-7, 'm_last'); close(Ll, Ll64, 1e-6, 'Lu_last')
  Zb, mb, Lb = rnd(C, P, D, seed=5), rnd(H, C, M, 1, seed=6), rnd(H, C, M, M, seed=7)
  lm, llv, pm, plv = (rnd(D + 1, seed=8 + i, scale=0.3) for i in range(4))
  eps, thb = rnd(H, D + 1, seed=12), rnd(H, D + 1, seed=13)
  gu, gh = torch.tensor([0.7], dtype=torch.float64), torch.tensor([1.3], dtype=torch.float64)
  for use_gu in (True, False):
    outs64 = [torch.empty(C, M, D, dtype=torch.float64), torch.empty(C, M, 1, dtype=torch.float64),
              torch.empty(C, T, dtype=torch.float64), torch.empty(D + 1, dtype=torch.float64), torch.empty(D + 1, dtype=torch.float64)]
    outs = [torch.full(o.shape, float('nan'), device='cuda') for o in outs64]
    EMU.step_grad_finish(Zb, mb, Lb, Ll64, ut, gu if use_gu else None, lm, llv, pm, plv, eps, thb, gh, *outs64)
    cuda_ops.step_grad_finish(dev(Zb), dev(mb), dev(Lb), Ll, dev(ut), dev(gu) if use_gu else None, dev(lm), dev(llv), dev(pm),
                              dev(plv), dev(eps), dev(thb), dev(gh), *outs)
    for a, b, nm in zip(outs, outs64, ('z_g', 'um_g', 'ut_g', 'lm_g', 'llv_g')):
      close(a, b, 1e-5, nm)


@pytest.mark.parametrize('H,C,S,M,rect', [(3, 10, 5, 60, None), (2, 3, 1, 7, None), (1, 4, 2, 9, None), (2, 2, 2, 96, None),
                                          (3, 4, 3, 33, (1, 2, 1, 4)), (2, 5, 2, 64, (0, 2, 2, 5)), (1, 2, 1, 128, None)])
def test_whiten_block_kernels(cuda_ops, H, C, S, M, rect):
  """whiten.cu (T, nu, N += T T^T, KL forward; Tbar / KL adjoint / whitening adjoint backward on the M x M task blocks in one
  shared-memory kernel each) vs the contracts in tests/emu_ops.py, incl. (h, c) sub-rectangles (factor sharding)."""
  P = S * M
  W = (rnd(H, C, P, P, seed=1, scale=0.1).tril() + torch.eye(P, dtype=torch.float64))
  Lu = rnd(S, C, M, M, seed=2, scale=0.2).tril() + torch.eye(M, dtype=torch.float64)
  m = rnd(S, C, M, seed=3)
  N0 = rnd(H, C, P, P, seed=4)
  T64, nu64, N64, kl64 = torch.zeros(H, C, S, M, M, dtype=torch.float64), torch.zeros(H, C, P, dtype=torch.float64), N0.clone(), \
      torch.zeros((), dtype=torch.float64)
  EMU.whiten_fwd(W, Lu, m, T64, nu64, N64, kl64, rect=rect)
  T, nu, N, kl = torch.zeros(H, C, S, M, M, device='cuda'), torch.zeros(H, C, P, device='cuda'), dev(N0), torch.zeros((), device='cuda')
  for rep in range(2):            # twice: the ticket workspace must come back zeroed
    cuda_ops.whiten_fwd(dev(W), dev(Lu), dev(m), T, nu, N, kl, rect=rect)
    if rep == 0:
      N.copy_(dev(N0)); kl.zero_()
  close(T, T64, 1e-6, 'T'); close(nu, nu64, 1e-6, 'nu'); close(N, N64, 1e-6, 'N'); close(kl, kl64, 1e-5, 'kl')
  if M > 96:
    return
  G = rnd(H, C, P, P, seed=5); G = G + G.transpose(-1, -2)
  nubar, Wb0 = rnd(H, C, P, seed=6), rnd(H, C, P, P, seed=7)
  g_kl = torch.tensor([0.7], dtype=torch.float64)
  for s0, use_kl in ((0, True), (S - 1, True), (0, False)):
    Sg = S - s0
    Wb64, Lb64, mb64 = Wb0.clone(), torch.zeros(H, Sg, C, M, M, dtype=torch.float64), torch.zeros(H, Sg, C, M, 1, dtype=torch.float64)
    EMU.whiten_bwd(W, T64, nu64, Lu, m, G, nubar, g_kl if use_kl else None, Wb64, Lb64, mb64, s_grad0=s0, rect=rect)
    Wb, Lb, mb = dev(Wb0), torch.zeros(H, Sg, C, M, M, device='cuda'), torch.zeros(H, Sg, C, M, 1, device='cuda')
    cuda_ops.whiten_bwd(dev(W), dev(T64), dev(nu64), dev(Lu), dev(m), dev(G), dev(nubar), dev(g_kl) if use_kl else None, Wb, Lb, mb,
                        s_grad0=s0, rect=rect)
    close(Wb, Wb64, 1e-6, 'Wbar'); close(Lb, Lb64, 1e-6, 'Lubar'); close(mb, mb64, 1e-6, 'mbar')
