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
  def forward(ctx, theta, Zcat, x, m_all, Lu_all, M, want_kl, shard=None):
    c = elbo._Ctx()
    f_mean, f_var, kl, info, L = elbo.marginal_forward(
      theta.detach().contiguous(), Zcat.detach().contiguous(), x.detach().contiguous(),
      m_all.detach().contiguous(), Lu_all.detach().contiguous(), M, want_kl, c, shard=shard)
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
    return th_bar, Z_bar, x_bar, m_bar, Lu_bar, None, None, None


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


class HyperFn(torch.autograd.Function):
  """(log_mean, log_logvar, prior_log_mean, prior_log_logvar, eps) -> (theta (H, D+1), kl_hypers): the
  reparameterised hyper sample and KL(q(theta) || p(theta)) of var_gp/kernels.py:62-77, one launch forward and one
  backward instead of ~30 elementwise launches."""

  @staticmethod
  def forward(ctx, log_mean, log_logvar, prior_log_mean, prior_log_logvar, eps):
    lm, lv = log_mean.detach().contiguous(), log_logvar.detach().contiguous()
    pm, plv, eps = prior_log_mean.detach().contiguous(), prior_log_logvar.detach().contiguous(), eps.contiguous()
    theta = torch.empty_like(eps)
    kl = torch.empty((), device=eps.device, dtype=eps.dtype)
    _ops().hyper_fwd(lm, lv, pm, plv, eps, theta, kl)
    ctx.save_for_backward(lm, lv, pm, plv, eps)
    ctx.set_materialize_grads(False)
    return theta, kl

  @staticmethod
  def backward(ctx, g_theta, g_kl):
    lm, lv, pm, plv, eps = ctx.saved_tensors
    m_bar, lv_bar = torch.empty_like(lm), torch.empty_like(lv)
    _ops().hyper_bwd(lm, lv, pm, plv, eps, None if g_theta is None else g_theta.contiguous(),
                     None if g_kl is None else g_kl.detach().reshape(1).contiguous(), m_bar, lv_bar)
    return m_bar, lv_bar, None, None, None


class Combine3Fn(torch.autograd.Function):
  """loss = coef . (kl_hypers, kl_u, nll)  (experiments/vargp.py:34) with 3 launches for forward + backward."""

  @staticmethod
  def forward(ctx, kl_h, kl_u, nll, coef):
    ctx.save_for_backward(coef)
    return torch.dot(torch.stack((kl_h.detach(), kl_u.detach(), nll.detach())), coef)

  @staticmethod
  def backward(ctx, g):
    coef, = ctx.saved_tensors
    g3 = g * coef
    return g3[0], g3[1], g3[2], None


class SoftmaxNllFn(torch.autograd.Function):
  """(f_mean, f_var, y, eps) -> nll; forward and adjoint come out of one kernel pass
  (var_gp/likelihoods.py:13-47)."""

  @staticmethod
  def forward(ctx, f_mean, f_var, y, eps):
    f_mean, f_var = f_mean.detach().contiguous(), f_var.detach().contiguous()
    ops = _ops()
    nw = ops.nll_work(eps.shape[0], eps.shape[-1]) if hasattr(ops, 'nll_work') else 0
    buf = torch.zeros(nw + 1, device=f_mean.device, dtype=f_mean.dtype)      # [workspace | nll]: one fill launch
    nll = buf[nw]
    gmv = torch.empty((2,) + tuple(f_mean.shape), device=f_mean.device, dtype=f_mean.dtype)
    if nw:
      ops.nll_fwd_bwd(f_mean, f_var, eps.contiguous(), y.contiguous(), nll, gmv[0], gmv[1], work=buf[:nw])
    else:
      ops.nll_fwd_bwd(f_mean, f_var, eps.contiguous(), y.contiguous(), nll, gmv[0], gmv[1])
    ctx.save_for_backward(gmv)
    return nll

  @staticmethod
  def backward(ctx, g):
    gmv, = ctx.saved_tensors
    gmv = gmv * g                       # one launch for both adjoints
    return gmv[0], gmv[1], None, None


def softmax_predict(f_mean, f_var, eps):
  """probs (B, C) = mean_{h,f} softmax_C(f_mean + sqrt(f_var) eps)   (var_gp/likelihoods.py:49-63)."""
  H, F, C, B = eps.shape
  probs = torch.empty(B, C, device=f_mean.device, dtype=f_mean.dtype)
  _ops().predict(f_mean.detach().contiguous(), f_var.detach().contiguous(), eps.contiguous(), probs)
  return probs
