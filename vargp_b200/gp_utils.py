"""Drop-in for ``var_gp/gp_utils.py``: same function names, signatures and ``cache=`` protocol
(keys ``Lz``, ``Lz_Kzx``), every batch dimension the reference accepts.

All arithmetic runs on libvargp_sm100.so through three differentiable primitives
(`cholesky`, `tri_solve`, `matmul`) whose backward passes are hand-written on the same kernels:
a triangular solve is "invert the factor once (vargp_trtri), then GEMM", which is what lets the solve
chain run as tensor-core GEMMs.  The fused training path (``elbo.py``) does not go through these
wrappers; they exist for API parity (``forward(x, loss_cache=...)``, the block-diagonal ablation,
``VARGPRetrain``-style callers).
"""
import torch
import torch.nn.functional as F

from . import ops as _ops_mod
from .functional import TrilUnpackFn


def _ops():
  return _ops_mod.get_ops()


def _flat(t):
  """(..., r, c) -> contiguous (b, r, c) and the batch shape."""
  return t.reshape(-1, *t.shape[-2:]).contiguous(), t.shape[:-2]


def _bcast(a, b):
  """Expand the batch dims of a and b to their broadcast shape (views, stride 0 where broadcast)."""
  bs = torch.broadcast_shapes(a.shape[:-2], b.shape[:-2])
  if len(bs) > 3:
    raise ValueError('at most 3 batch dimensions are supported')
  return a.expand(*bs, *a.shape[-2:]), b.expand(*bs, *b.shape[-2:]), bs


class _MatmulFn(torch.autograd.Function):
  @staticmethod
  def forward(ctx, A, B):
    A, B = A.detach(), B.detach()
    Ae, Be, bs = _bcast(A, B)
    C = torch.empty(*bs, A.shape[-2], B.shape[-1], device=A.device, dtype=A.dtype)
    _ops().gemm(Ae, Be, C)
    ctx.save_for_backward(A, B)
    return C

  @staticmethod
  def backward(ctx, g):
    A, B = ctx.saved_tensors
    g = g.contiguous()
    Ae, Be, bs = _bcast(A, B)
    gA = torch.empty(*bs, *A.shape[-2:], device=A.device, dtype=A.dtype)
    gB = torch.empty(*bs, *B.shape[-2:], device=A.device, dtype=A.dtype)
    _ops().gemm(g, Be.transpose(-1, -2), gA)
    _ops().gemm(Ae.transpose(-1, -2), g, gB)
    return gA.sum_to_size(A.shape), gB.sum_to_size(B.shape)


def matmul(A, B):
  """Batched A @ B on the library GEMM (broadcasting batch dims)."""
  return _MatmulFn.apply(A, B)


class _CholFn(torch.autograd.Function):
  @staticmethod
  def forward(ctx, A, eps):
    Af, bshape = _flat(A.detach())
    L = torch.empty_like(Af)
    info = torch.zeros(Af.shape[0], device=A.device, dtype=torch.int32)
    _ops().chol(Af, L, eps, info)
    bad = int(info.max().item())          # same synchronous failure mode as torch.cholesky
    if bad:
      raise torch.linalg.LinAlgError(
        f'linalg.cholesky: the input is not positive-definite (leading minor of order {bad} is not positive-definite)')
    ctx.save_for_backward(L)
    ctx.bshape = bshape
    return L.view(*bshape, *L.shape[-2:])

  @staticmethod
  def backward(ctx, g):
    # Abar = W^T [(Phi(L^T Lbar) + Phi(L^T Lbar)^T) / 2] W,  W = L^-1
    L, = ctx.saved_tensors
    ops = _ops()
    gf, _ = _flat(g)
    W = torch.empty_like(L)
    ops.trtri(L, W)
    S = torch.empty_like(L)
    ops.gemm(L.transpose(-1, -2), gf, S, a_tri='upper', b_tri='lower', c_tri='lower')
    ops.sym_phi(S)
    Y = torch.empty_like(L)
    ops.gemm(S, W, Y, b_tri='lower')
    Abar = torch.empty_like(L)
    ops.gemm(W.transpose(-1, -2), Y, Abar, a_tri='upper')
    return Abar.view(*ctx.bshape, *L.shape[-2:]), None


class _TriSolveFn(torch.autograd.Function):
  """X = L^-1 B (left, lower, no transpose)."""

  @staticmethod
  def forward(ctx, L, B):
    ops = _ops()
    Lf, lshape = _flat(L.detach())
    W = torch.empty_like(Lf)
    ops.trtri(Lf, W)
    W = W.view(*lshape, *Lf.shape[-2:])
    B = B.detach()
    We, Be, bs = _bcast(W, B)
    X = torch.empty(*bs, *B.shape[-2:], device=B.device, dtype=B.dtype)
    ops.gemm(We, Be, X, a_tri='lower')
    ctx.save_for_backward(W, X)
    ctx.shapes = (L.shape, B.shape)
    return X

  @staticmethod
  def backward(ctx, g):
    W, X = ctx.saved_tensors
    ops = _ops()
    g = g.contiguous()
    We = W.expand(*X.shape[:-2], *W.shape[-2:])
    Bbar = torch.empty_like(X)
    ops.gemm(We.transpose(-1, -2), g, Bbar, a_tri='upper')
    Lbar = torch.empty(*X.shape[:-2], *W.shape[-2:], device=X.device, dtype=X.dtype)
    ops.gemm(Bbar, X.transpose(-1, -2), Lbar, alpha=-1., c_tri='lower')
    return Lbar.sum_to_size(ctx.shapes[0]), Bbar.sum_to_size(ctx.shapes[1])


def tri_solve(L, B):
  """torch.triangular_solve(B, L, upper=False)[0] on the library (W = L^-1 once, then a GEMM)."""
  return _TriSolveFn.apply(L, B)


# ---------------------------------------------------------------------------------------------
# reference API
# ---------------------------------------------------------------------------------------------
def cholesky(M, eps=1e-4):
  """L with M + eps I = L L^T   (var_gp/gp_utils.py:5-11)."""
  return _CholFn.apply(M, float(eps))


def rev_cholesky(L):
  """M = L L^T   (var_gp/gp_utils.py:14-19)."""
  return matmul(L, L.transpose(-1, -2))


def vec2tril(vec, m=None):
  """Packed row-major lower triangle -> (..., m, m), softplus on the diagonal (var_gp/gp_utils.py:22-49)."""
  if m is None:
    D = vec.size(-1)
    m = int(((torch.tensor(8. * D + 1).sqrt() - 1.) / 2.).long().item())
  if vec.is_cuda:
    flat = vec.reshape(-1, vec.size(-1))
    return TrilUnpackFn.apply(flat, m).view(*vec.shape[:-1], m, m)
  # host-side bookkeeping (parameter init / checkpoint conversion happens on CPU before .to(device))
  idx = torch.tril_indices(m, m)
  tril = torch.zeros(*vec.shape[:-1], m, m, dtype=vec.dtype)
  tril[..., idx[0], idx[1]] = vec
  return torch.where(torch.eye(m).bool(), F.softplus(tril), tril)


def mat2trilvec(mat):
  """(..., m, m) -> packed lower triangle (var_gp/gp_utils.py:52-65); pure indexing."""
  m = mat.size(-1)
  idx = torch.tril_indices(m, m, device=mat.device)
  return mat[..., idx[0], idx[1]]


def _atb(A, B):
  """einsum('...ij,...ik->...jk', A, B) = A^T B."""
  return matmul(A.transpose(-1, -2), B)


def gp_cond(u, Kzz, Kzx, Kxx, Lz=None, Lz_Kzx=None):
  """mu = Kxz Kzz^-1 u, Sigma = Kxx - Kxz Kzz^-1 Kzx   (var_gp/gp_utils.py:68-98)."""
  if Lz is None:
    Lz = cholesky(Kzz)
  Lz_u = tri_solve(Lz, u)
  if Lz_Kzx is None:
    Lz_Kzx = tri_solve(Lz, Kzx)
  mu = _atb(Lz_Kzx, Lz_u)
  Sigma = Kxx - _atb(Lz_Kzx, Lz_Kzx)
  return mu, Sigma


def linear_joint(m, S, Kzx, Kzz, V, b, cache=None):
  """Joint of N(z; m, S) N(x; Az + b, V), A = Kxz Kzz^-1   (var_gp/gp_utils.py:101-147)."""
  Lz = cholesky(Kzz)
  Lz_m = tri_solve(Lz, m)
  Lz_Kzx = tri_solve(Lz, Kzx)
  Am = _atb(Lz_Kzx, Lz_m)
  Lz_S = tri_solve(Lz, S)
  AS = _atb(Lz_Kzx, Lz_S)
  SAt = AS.transpose(-1, -2)
  Lz_SAt = tri_solve(Lz, SAt)
  ASAt = _atb(Lz_SAt, Lz_Kzx)
  mu = torch.cat([m, Am + b], dim=-2)
  Sigma = torch.cat([torch.cat([S, SAt], dim=-1), torch.cat([AS, V + ASAt], dim=-1)], dim=-2)
  if isinstance(cache, dict):
    cache.update(dict(Lz_Kzx=Lz_Kzx, Lz=Lz))
  return mu, Sigma


def linear_marginal_diag(m, S, Kzz, Kzx, Kxx_diag, cache=None):
  """Diagonal of the marginal of N(z; m, S) N(y; Az, V)   (var_gp/gp_utils.py:150-191)."""
  Lz = cholesky(Kzz)
  Lz_m = tri_solve(Lz, m)
  Lz_Kzx = tri_solve(Lz, Kzx)
  mu = _atb(Lz_Kzx, Lz_m).squeeze(-1)
  diag1 = Lz_Kzx.pow(2).sum(dim=-2)
  Lz_LS = tri_solve(Lz, cholesky(S))
  diag2 = _atb(Lz_LS, Lz_Kzx).pow(2).sum(dim=-2)
  Sigma = Kxx_diag - diag1 + diag2
  if isinstance(cache, dict):
    cache.update(dict(Lz=Lz, Lz_Kzx=Lz_Kzx))
  return mu, Sigma
