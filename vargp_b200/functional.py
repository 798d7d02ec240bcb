"""torch.autograd glue: each Function launches kernels of libvargp_sm100.so in forward and the hand-derived
adjoint kernels in backward (no autograd tape through the numerics)."""
import torch

from . import elbo
from . import ops as _ops_mod


def _ops():
  return _ops_mod.get_ops()


class MarginalFn(torch.autograd.Function):
  """(theta, Zcat, x, m_all, Lu_all) -> (f_mean, f_var, kl_u, info, L).

  One launch sequence replaces compute_q + compute_pf_diag + the KL ingredients of
  var_gp/vargp.py:35-175 (see elbo.py for the algebra).  `info` (int32, H*C) is the Cholesky status,
  `L` the Cholesky factor of K(z_<=t) + eps I (both non-differentiable by-products).
  """

  @staticmethod
  def forward(ctx, theta, Zcat, x, m_all, Lu_all, M, want_kl):
    c = elbo._Ctx()
    f_mean, f_var, kl, info, L = elbo.marginal_forward(
      theta.detach().contiguous(), Zcat.detach().contiguous(), x.detach().contiguous(),
      m_all.detach().contiguous(), Lu_all.detach().contiguous(), M, want_kl, c)
    ctx.c = c
    ctx.need_x = x.requires_grad
    ctx.want_kl = want_kl
    ctx.set_materialize_grads(False)
    if kl is None:
      kl = torch.zeros((), device=x.device, dtype=x.dtype)
    ctx.mark_non_differentiable(info, L)
    return f_mean, f_var, kl, info, L

  @staticmethod
  def backward(ctx, g_mean, g_var, g_kl, _g_info, _g_L):
    if not ctx.want_kl:
      g_kl = None
    if g_kl is not None:
      g_kl = g_kl.detach().reshape(1).contiguous()
    th_bar, Z_bar, x_bar, m_bar, Lu_bar = elbo.marginal_backward(ctx.c, g_mean, g_var, g_kl, need_x_grad=ctx.need_x)
    ctx.c = None
    return th_bar, Z_bar, x_bar, m_bar, Lu_bar, None, None


class TrilUnpackFn(torch.autograd.Function):
  """vec (C, M(M+1)/2) -> lower-triangular (C, M, M) with softplus diagonal (var_gp/gp_utils.py:22-49)."""

  @staticmethod
  def forward(ctx, vec, M):
    vec = vec.detach().contiguous()
    out = torch.empty(vec.shape[0], M, M, device=vec.device, dtype=vec.dtype)
    _ops().tril_unpack(vec, out)
    ctx.save_for_backward(vec)
    return out

  @staticmethod
  def backward(ctx, g):
    vec, = ctx.saved_tensors
    gv = torch.empty_like(vec)
    _ops().tril_unpack_bwd(g.contiguous(), vec, gv)
    return gv, None


class SoftmaxNllFn(torch.autograd.Function):
  """(f_mean, f_var, y, eps) -> nll; forward and adjoint come out of one kernel pass
  (var_gp/likelihoods.py:13-47)."""

  @staticmethod
  def forward(ctx, f_mean, f_var, y, eps):
    f_mean, f_var = f_mean.detach().contiguous(), f_var.detach().contiguous()
    nll = torch.zeros((), device=f_mean.device, dtype=f_mean.dtype)
    gm, gv = torch.empty_like(f_mean), torch.empty_like(f_var)
    _ops().nll_fwd_bwd(f_mean, f_var, eps.contiguous(), y.contiguous(), nll, gm, gv)
    ctx.save_for_backward(gm, gv)
    return nll

  @staticmethod
  def backward(ctx, g):
    gm, gv = ctx.saved_tensors
    return gm * g, gv * g, None, None


def softmax_predict(f_mean, f_var, eps):
  """probs (B, C) = mean_{h,f} softmax_C(f_mean + sqrt(f_var) eps)   (var_gp/likelihoods.py:49-63)."""
  H, F, C, B = eps.shape
  probs = torch.empty(B, C, device=f_mean.device, dtype=f_mean.dtype)
  _ops().predict(f_mean.detach().contiguous(), f_var.detach().contiguous(), eps.contiguous(), probs)
  return probs
