"""Fused sparse-variational marginal + KL(u) for VAR-GP, with a hand-derived backward.

This is the host-side schedule of the hot path: it owns no arithmetic.  Every number is produced by a
kernel of ``libvargp_sm100.so`` reached through ``vargp_b200.ops`` (C-ABI, raw pointers + strides).

Formulation.  For each hyper sample h and class c, with
Z = [z_0; ...; z_t] (P = (t+1) M rows) and K = K_h(Z, Z) + eps I:

    W   = chol(K)^-1                         one Cholesky + one triangular inverse   (replaces the t+2
                                             Choleskys of var_gp/vargp.py:61-80,108,155)
    T_s = W_ss Lu_s ,  nu_s = W_ss m_s       whitened variational factor / mean: the autoregressive joint
                                             of var_gp/gp_utils.py:101-147 is block-diagonal here
    V   = W Kzx ,  N = blockdiag(T_s T_s^T) + eps W W^T   (P x P, symmetric, once per step)
    f_mean_b = nu . V_b
    f_var_b  = gamma^2 - |V_b|^2 + V_b^T N V_b                                (gp_utils.py:150-191)
             = gamma^2 - |V_b|^2 + sum_s |T_s^T V_sb|^2 + eps |W^T V_b|^2
    KL_hc    = -sum_{i in t} log W_ii - sum_i log Lu_t,ii + (|T_t|_F^2 + |nu_t|^2 - M) / 2
    kl_u     = (1/H) sum_hc KL_hc                                             (vargp.py:182-190)

(the KL line holds for ``ep_var_mean=True``; the block-diagonal ablation goes through
``vargp_b200.gp_utils``).  All O(P^2 B) and O(P^3) work is expressed as batched GEMMs so it can run on
the tensor cores; nothing of size B x B is ever formed.
"""
import math
import os

import torch

from . import ops as _ops_mod

JITTER = 1e-4   # var_gp/gp_utils.py:5


def _ops():
  return _ops_mod.get_ops()


class _Ctx:
  """Plain container for tensors saved between forward and backward."""
  pass


_SIDE = {}
USE_SIDE_STREAM = os.environ.get('VARGP_STREAMS', '1') != '0'


class _Fork:
  """`with fork:` queues the enclosed launches on a side stream, ordered after everything already queued on the
  current stream; `fork.join()` makes the current stream wait for them.  The 30-matrix Cholesky and the chain of
  P x P adjoint GEMMs leave most SMs idle, so the minibatch-sized Kzx branch runs beside them (both branches are
  captured into the step's CUDA graph as parallel paths).  Rules that keep the caching allocator safe: every tensor
  the side branch touches is allocated on the main stream BEFORE the fork and stays referenced until after join()."""

  def __init__(self, dev):
    self.side = None
    if USE_SIDE_STREAM and dev.type == 'cuda':
      key = dev.index if dev.index is not None else torch.cuda.current_device()
      if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
      self.side = _SIDE[key]
    self._ctx = None

  def __enter__(self):
    if self.side is not None:
      self.side.wait_stream(torch.cuda.current_stream())
      self._ctx = torch.cuda.stream(self.side)
      self._ctx.__enter__()
    return self

  def __exit__(self, *exc):
    if self._ctx is not None:
      self._ctx.__exit__(*exc)
      self._ctx = None
    return False

  def join(self):
    if self.side is not None:
      torch.cuda.current_stream().wait_stream(self.side)


def _zeros_many(dev, dt, *shapes):
  """Zero-filled tensors of the given shapes carved out of ONE buffer (one fill launch; 128 B aligned segments)."""
  sizes = [math.prod(sh) for sh in shapes]
  offs, tot = [], 0
  for n in sizes:
    offs.append(tot)
    tot += (n + 31) // 32 * 32
  flat = torch.zeros(tot, device=dev, dtype=dt)
  return [flat[o:o + n].view(sh) for o, n, sh in zip(offs, sizes, shapes)]


def _blocks(mat, S, M):
  """(H, C, P, P) -> view (H, C, S, M, M) of the S diagonal M x M blocks (no copy)."""
  H, C, P, _ = mat.shape
  sH, sC, sR, sC2 = mat.stride()
  return mat.as_strided((H, C, S, M, M), (sH, sC, M * sR + M * sC2, sR, sC2))


def _rows(mat, S, M):
  """(H, C, P, B) -> view (H, C, S, M, B) of the S row blocks."""
  H, C, P, B = mat.shape
  sH, sC, sR, sB = mat.stride()
  return mat.as_strided((H, C, S, M, B), (sH, sC, M * sR, sR, sB))


def marginal_forward(theta, Zcat, x, m_all, Lu_all, M, want_kl, ctx=None):
  """theta (H, D+1); Zcat (C, P, D); x (B, D); m_all (S, C, M); Lu_all (S, C, M, M) lower.

  Returns f_mean, f_var (H, C, B), kl_u (0-d tensor or None) and fills ``ctx`` for the backward.
  """
  ops = _ops()
  H, D1 = theta.shape
  D = D1 - 1
  C, P, _ = Zcat.shape
  B = x.shape[0]
  S = P // M
  dev, dt = x.device, x.dtype
  new = lambda *s: torch.empty(*s, device=dev, dtype=dt)

  # (1) scaled operands and their squared norms                               [kernels.py:41-44,50]
  zs, zn = new(H, C * P, D), new(H, C * P)
  xs, xn = new(H, B, D), new(H, B)
  Kzz, Kzx = new(H, C, P, P), new(H, C, P, B)
  ops.scale_rows(Zcat.reshape(C * P, D), theta, zs, zn)
  zs4, zn3 = zs.view(H, C, P, D), zn.view(H, C, P)

  # (2) Gram matrices                                                          [kernels.py:45-56]
  #     the x side and Kzx (minibatch-sized) run on the side stream next to Kzz -> Cholesky -> whitening -> KL -> N
  fork = _Fork(dev)
  with fork:
    ops.scale_rows(x, theta, xs, xn)
    ops.rbf_gram(zs4, zn3, xs.view(H, 1, B, D), xn.view(H, 1, B), theta, Kzx, False, tag='Kzx')
  ops.rbf_gram(zs4, zn3, zs4, zn3, theta, Kzz, True, tag='Kzz')

  # (3) W = chol(Kzz + eps I)^-1                                               [gp_utils.py:5-11]
  L, W = new(H, C, P, P), new(H, C, P, P)
  if dt == torch.float32:       # Cholesky status words and the KL accumulator share one zero-filled buffer
    zb = torch.zeros(H * C + 1, device=dev, dtype=dt)
    info, kl0 = zb[:H * C].view(torch.int32), zb[H * C]
  else:
    info, kl0 = torch.zeros(H * C, device=dev, dtype=torch.int32), torch.zeros((), device=dev, dtype=dt)
  ops.chol_inv(Kzz, L, W, JITTER, info)

  # (4) whitened variational parameters (block diagonal)
  T, nu = new(H, C, S, M, M), new(H, C, P)
  Wd = _blocks(W, S, M)
  LuB = Lu_all.permute(1, 0, 2, 3).unsqueeze(0)                 # (1, C, S, M, M), broadcast over h
  ops.gemm(Wd, LuB, T, a_tri='lower', b_tri='lower', tag='T=Wss*Lu', zeroed=True)
  mB = m_all.permute(1, 0, 2).unsqueeze(0).unsqueeze(-1)        # (1, C, S, M, 1)
  ops.gemm(Wd, mB, nu.view(H, C, S, M, 1), a_tri='lower', tag='nu=Wss*m', zeroed=True)

  # (5) KL(q(u_t | u_<t) || p(u_t | u_<t))                                     [vargp.py:182-190]
  kl = None
  if want_kl:
    kl = kl0
    ops.kl_fwd(W, T, nu, Lu_all[S - 1], M, kl)

  # (6) predictive marginal                                                    [gp_utils.py:150-191]
  #     N = blockdiag(T_s T_s^T) + eps W W^T collects everything quadratic in V, so the minibatch-sized work is
  #     two GEMMs (V, N V) and one streaming reduction instead of three GEMMs here and four more in the backward
  N = new(H, C, P, P)
  ops.gemm(W, W.transpose(-1, -2), N, alpha=JITTER, a_tri='lower', b_tri='upper', tag='N=eps*W*Wt', zeroed=True)
  ops.gemm(T, T.transpose(-1, -2), _blocks(N, S, M), beta=1., a_tri='lower', b_tri='upper', tag='N+=T*Tt',
           zeroed=True)
  V, NV = new(H, C, P, B), new(H, C, P, B)
  fork.join()
  ops.gemm(W, Kzx, V, a_tri='lower', tag='V=W*Kzx', zeroed=True)
  ops.gemm(N, V, NV, tag='NV=N*V')
  f_mean, f_var = new(H, C, B), new(H, C, B)
  ops.marginal_reduce(V, NV, nu, theta, f_mean, f_var)

  if ctx is not None:
    ctx.dims = (H, C, P, B, D, S, M)
    ctx.saved = dict(theta=theta, zs=zs4, xs=xs, Kzz=Kzz, Kzx=Kzx, W=W, T=T, nu=nu, V=V, NV=NV,
                     m_all=m_all, Lu_all=Lu_all)
  return f_mean, f_var, kl, info, L


def marginal_backward(ctx, g_mean, g_var, g_kl, need_x_grad=False):
  """Adjoint of `marginal_forward`.  g_mean, g_var (H, C, B) or None; g_kl 0-d tensor or None.

  Returns grads (theta (H, D+1), Zcat (C, P, D), x (B, D) or None, m_all (S, C, M), Lu_all (S, C, M, M)).
  """
  ops = _ops()
  H, C, P, B, D, S, M = ctx.dims
  sv = ctx.saved
  theta, zs, xs, Kzz, Kzx, W, T, nu = (sv[k] for k in ('theta', 'zs', 'xs', 'Kzz', 'Kzx', 'W', 'T', 'nu'))
  V, NV, m_all, Lu_all = (sv[k] for k in ('V', 'NV', 'm_all', 'Lu_all'))
  dev, dt = V.device, V.dtype
  new = lambda *s: torch.empty(*s, device=dev, dtype=dt)
  zeros = lambda *s: torch.zeros(*s, device=dev, dtype=dt)
  Wd = _blocks(W, S, M)
  LuB = Lu_all.permute(1, 0, 2, 3).unsqueeze(0)                 # (1, C, S, M, M)
  mB = m_all.permute(1, 0, 2).unsqueeze(0)                      # (1, C, S, M)

  Wbar, Tbar, nubar, theta_bar, r1z, csumz = _zeros_many(dev, dt, (H, C, P, P), (H, C, S, M, M), (H, C, P), (H, D + 1),
                                                         (H, C, P), (H, B))
  Kxbar = Gz1 = Gx = r1 = csum = None
  fork = _Fork(dev)
  have_data = g_mean is not None or g_var is not None
  if have_data:
    if g_mean is None:
      g_mean = torch.zeros_like(g_var)
    if g_var is None:
      g_var = torch.zeros_like(g_mean)
    g_mean, g_var = g_mean.contiguous(), g_var.contiguous()
    # Vbar = nu gm^T + 2 gv (N V - V)  (overwrites NV) ;  Vg = gv V ;  theta_bar[:, D] += 2 gamma^2 sum_cb gv
    Vbar, Vg = NV, new(H, C, P, B)
    ops.marginal_bwd_prep(V, NV, nu, g_mean, g_var, theta, Vbar, Vg, theta_bar)
    # Kzx side of the adjoint on the side stream: Kzx_bar = W^T Vbar -> (.) Kzx, row / column sums -> Gz1 = Wk1 xs
    Kxbar, Gz1 = new(H, C, P, B), new(H, C, P, D)
    Gx = new(H, C, B, D) if need_x_grad else None
    r1, csum = r1z, csumz
    with fork:
      ops.gemm(W.transpose(-1, -2), Vbar, Kxbar, a_tri='upper', tag='Kxbar=Wt*Vbar', zeroed=True)
      ops.rbf_bwd_prep(Kxbar, Kzx, r1, csum)                    # Kxbar <- Kxbar * Kzx ; col sums over (c, i)
      ops.gemm(Kxbar, xs.view(H, 1, B, D), Gz1, tag='Gz1=Wk1*xs')
      if need_x_grad:
        ops.gemm(Kxbar.transpose(-1, -2), zs, Gx, tag='Gx=Wk1t*zs')
    # Wbar = tril(Vbar Kzx^T)
    ops.gemm(Vbar, Kzx.transpose(-1, -2), Wbar, c_tri='lower', tag='Wbar=Vbar*Kzxt')
    # Nbar = G = sum_b gv_b V_b V_b^T (symmetric: lower triangle by GEMM, then mirrored)
    G = new(H, C, P, P)
    ops.gemm(Vg, V.transpose(-1, -2), G, c_tri='lower', tag='G=Vg*Vt')
    ops.sym_phi(G, mirror=True)
    # N = blockdiag(T_s T_s^T) + eps W W^T  =>  Tbar_s = tril(2 G_ss T_s),  Wbar += tril(2 eps G W)
    ops.gemm(_blocks(G, S, M), T, Tbar, alpha=2., b_tri='lower', c_tri='lower', tag='Tbar=2*Gss*T', zeroed=True)
    ops.gemm(G, W, Wbar, alpha=2. * JITTER, beta=1., b_tri='lower', c_tri='lower', tag='Wbar+=2eps*G*W', zeroed=True)
    # nubar = V gm
    ops.gemm(V, g_mean.unsqueeze(-1), nubar.unsqueeze(-1), tag='nubar=V*gm')
  if g_kl is not None:
    # adds (g_kl/H) T_t, (g_kl/H) nu_t and -(g_kl/H)/W_ii on the last block
    ops.kl_bwd(W, T, nu, M, g_kl, Wbar, Tbar, nubar)

  # whitening adjoint: Wbar_ss += tril(Tbar_s Lu_s^T + nubar_s m_s^T); Lu_bar_s = sum_h W_ss^T Tbar_s ; m_bar_s = sum_h W_ss^T nubar_s
  Wbd = _blocks(Wbar, S, M)
  ops.gemm(Tbar, LuB.transpose(-1, -2), Wbd, beta=1., a_tri='lower', b_tri='upper', c_tri='lower', tag='whiten_adj',
           zeroed=True)
  ops.gemm(nubar.view(H, C, S, M, 1), mB.unsqueeze(-2), Wbd, beta=1., c_tri='lower', tag='whiten_adj')
  #   (the per-h products are written task-major, so that the sum over h is already in parameter layout)
  Lubar_h = new(H, S, C, M, M)
  ops.gemm(Wd.transpose(-1, -2), Tbar, Lubar_h.permute(0, 2, 1, 3, 4), a_tri='upper', b_tri='lower', c_tri='lower',
           tag='whiten_adj', zeroed=True)
  mbar_h = new(H, S, C, M, 1)
  ops.gemm(Wd.transpose(-1, -2), nubar.view(H, C, S, M, 1), mbar_h.permute(0, 2, 1, 3, 4), a_tri='upper',
           tag='whiten_adj', zeroed=True)
  Lu_bar = Lubar_h.sum(0)                                        # (S, C, M, M)
  m_bar = mbar_h.sum(0).squeeze(-1)                              # (S, C, M)
  if g_kl is not None:
    # d/dLu_t of -sum_i log Lu_t,ii  (mean over h of H identical terms)
    ops.kl_bwd_lu(Lu_all[S - 1], g_kl, Lu_bar[S - 1])

  # Cholesky-inverse adjoint:  Kbar = -W^T Xi W,  Xi = (Phi(X) + Phi(X)^T)/2,  X = tril(Wbar W^T)
  X = new(H, C, P, P)
  ops.gemm(Wbar, W.transpose(-1, -2), X, a_tri='lower', b_tri='upper', c_tri='lower', tag='X=Wbar*Wt', zeroed=True)
  ops.sym_phi(X)                                                 # in place -> Xi (full symmetric)
  Y = new(H, C, P, P)
  ops.gemm(X, W, Y, b_tri='lower', tag='Y=Xi*W', zeroed=True)
  Kzzbar = new(H, C, P, P)
  ops.gemm(W.transpose(-1, -2), Y, Kzzbar, alpha=-1., a_tri='upper', tag='Kzzbar=-Wt*Y', zeroed=True)

  # RBF adjoint                                                                 
  r2, dg = new(H, C, P), new(H, C, P)
  ops.rbf_bwd_prep(Kzzbar, Kzz, r2, None, dg)                   # Kzzbar <- Kzzbar * Kzz (diag -> dg) ; row sums
  Gz2 = new(H, C, P, D)
  ops.gemm(Kzzbar, zs, Gz2, tag='Gz2=Wk2*zs')
  Z_bar = new(C, P, D)
  fork.join()
  ops.rbf_bwd_finish(zs, Gz1, Gz2, r1, r2, theta, Z_bar, theta_bar, dg)
  x_bar = None
  if have_data:
    x_bar = new(B, D) if need_x_grad else None
    ops.rbf_bwd_xside(xs, csum, Gx, theta, theta_bar, x_bar)
  return theta_bar, Z_bar, x_bar, m_bar, Lu_bar
