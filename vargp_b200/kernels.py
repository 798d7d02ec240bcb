"""Drop-in for ``var_gp/kernels.py``: ARD-RBF kernel with a log-normal variational hyper-posterior.

Same class names, constructor arguments, parameter / buffer names (``log_mean``, ``log_logvar``,
``prior_log_mean``, ``prior_log_logvar``, ``phi.*``) and RNG draw as the reference, so state dicts are
interchangeable.  ``compute`` runs on libvargp_sm100.so (scale -> Gram GEMM with fused exp epilogue); its
backward is the hand-written adjoint, not an autograd tape.
"""
import torch
import torch.nn as nn

from . import ops as _ops_mod


def _ops():
  return _ops_mod.get_ops()


def _is_broadcast_batch(t):
  """True when every batch dim of (..., N, D) `t` is size 1 or stride 0 (e.g. x.unsqueeze(0).expand(C, -1, -1))."""
  return all(sz == 1 or st == 0 for sz, st in zip(t.shape[:-2], t.stride()[:-2]))


class RbfComputeFn(torch.autograd.Function):
  """K[h, c, i, j] = gamma_h^2 exp(-1/2 sum_d (x_cid - y_cjd)^2 / sigma_hd^2)   (var_gp/kernels.py:24-56).

  theta (H, D+1); x (C, Pa, D); y None | (C, Pb, D) | (1, Pb, D) shared by all classes.
  """

  @staticmethod
  def forward(ctx, theta, x, y, shared_y):
    ops = _ops()
    theta, x = theta.detach().contiguous(), x.detach().contiguous()
    H, D = theta.shape[0], theta.shape[1] - 1
    C, Pa, _ = x.shape
    new = lambda *s: torch.empty(*s, device=x.device, dtype=x.dtype)
    xs, xn = new(H, C * Pa, D), new(H, C * Pa)
    ops.scale_rows(x.reshape(C * Pa, D), theta, xs, xn)
    xs4, xn3 = xs.view(H, C, Pa, D), xn.view(H, C, Pa)
    if y is None:
      K = new(H, C, Pa, Pa)
      ops.rbf_gram(xs4, xn3, xs4, xn3, theta, K, True)
      ys4 = None
    else:
      y = y.detach().contiguous()
      Cy, Pb, _ = y.shape
      ys, yn = new(H, Cy * Pb, D), new(H, Cy * Pb)
      ops.scale_rows(y.reshape(Cy * Pb, D), theta, ys, yn)
      ys4, yn3 = ys.view(H, Cy, Pb, D), yn.view(H, Cy, Pb)
      K = new(H, C, Pa, Pb)
      ops.rbf_gram(xs4, xn3, ys4, yn3, theta, K, False)
    ctx.sym = y is None
    ctx.shared_y = shared_y
    ctx.save_for_backward(theta, xs4, ys4 if ys4 is not None else xs4, K)
    return K

  @staticmethod
  def backward(ctx, g):
    ops = _ops()
    theta, xs, ys, K = ctx.saved_tensors
    H, C, Pa, D = xs.shape
    new = lambda *s: torch.empty(*s, device=K.device, dtype=K.dtype)
    theta_bar = torch.zeros(H, D + 1, device=K.device, dtype=K.dtype)
    x_bar = new(C, Pa, D)
    if ctx.sym:
      Wk = (0.5 * (g + g.transpose(-1, -2))).contiguous()     # only the symmetric part of Kbar acts
      r, dg = new(H, C, Pa), new(H, C, Pa)
      ops.rbf_bwd_prep(Wk, K, r, None, dg)
      G = new(H, C, Pa, D)
      ops.gemm(Wk, xs, G)
      ops.rbf_bwd_finish(xs, None, G, None, r, theta, x_bar, theta_bar, dg)
      return theta_bar, x_bar, None, None
    Wk = g.contiguous().clone()
    Cy, Pb = ys.shape[1], ys.shape[2]
    r = new(H, C, Pa)
    if ctx.shared_y:
      csum = torch.zeros(H, Pb, device=K.device, dtype=K.dtype)
      ops.rbf_bwd_prep(Wk, K, r, csum)
      G = new(H, C, Pa, D)
      ops.gemm(Wk, ys, G)                                      # ys (H, 1, Pb, D) broadcasts over classes
      ops.rbf_bwd_finish(xs, G, None, r, None, theta, x_bar, theta_bar)
      Gx = new(H, C, Pb, D)
      ops.gemm(Wk.transpose(-1, -2), xs, Gx)
      y_bar = new(Pb, D)
      ops.rbf_bwd_xside(ys.reshape(H, Pb, D), csum, Gx, theta, theta_bar, y_bar)
      return theta_bar, x_bar, y_bar.unsqueeze(0), None
    # per-class y: both sides take the "symmetric" form of the finish kernel with half weights, so the
    # cross term of theta_bar and the gamma term are each counted once per side
    ops.rbf_bwd_prep(Wk, K, r, None)
    csum = Wk.sum(-2)
    G = new(H, C, Pa, D)
    ops.gemm(Wk, ys, G, alpha=0.5)
    ops.rbf_bwd_finish(xs, None, G, None, (0.5 * r).contiguous(), theta, x_bar, theta_bar)
    Gy = new(H, C, Pb, D)
    ops.gemm(Wk.transpose(-1, -2), xs, Gy, alpha=0.5)
    y_bar = new(C, Pb, D)
    ops.rbf_bwd_finish(ys, None, Gy, None, (0.5 * csum).contiguous(), theta, y_bar, theta_bar)
    return theta_bar, x_bar, y_bar, None


class RBFKernel(nn.Module):
  """var_gp/kernels.py:7-77."""

  def __init__(self, in_size, prior_log_mean=None, prior_log_logvar=None, map_est=False):
    super().__init__()
    self.map_est = map_est
    # variational parameters: log lengthscales (in_size) and log scale factor (1); same init as the reference
    init = torch.tensor(.5).log() * torch.ones(in_size + 1) + .05 * torch.randn(in_size + 1)
    self.log_mean = nn.Parameter(init)
    self.log_logvar = nn.Parameter(-2 * torch.ones(in_size + 1))
    self.register_buffer('prior_log_mean',
                         prior_log_mean if prior_log_mean is not None else torch.zeros_like(self.log_mean))
    self.register_buffer('prior_log_logvar',
                         prior_log_logvar if prior_log_logvar is not None else torch.zeros_like(self.log_logvar))

  def features(self, x):
    """Input map applied before the RBF (identity here; an MLP in DeepRBFKernel)."""
    return x

  def compute(self, kern_samples, x, y=None):
    """kern_samples (H, D+1); x (..., M, D); y (..., N, D) or None -> (H, ..., M, N)."""
    x, y = self.features(x), (None if y is None else self.features(y))
    return rbf_compute(kern_samples, x, y)

  def compute_diag(self, kern_samples):
    return (kern_samples[..., -1:] * 2.).exp().unsqueeze(-2)

  def sample_hypers(self, n_hypers, eps=None):
    """theta = log_mean + sqrt(exp(log_logvar)) * eps, eps ~ N(0, I) of shape (n_hypers, D+1)
    (same draw as Normal.rsample in var_gp/kernels.py:62-68); `eps` may be supplied to pin the noise."""
    if self.map_est:
      return self.log_mean.unsqueeze(0)
    if eps is None:
      eps = torch.empty((n_hypers,) + self.log_mean.shape, dtype=self.log_mean.dtype,
                        device=self.log_mean.device).normal_()
    return self.log_mean + self.log_logvar.exp().sqrt() * eps

  def sample_hypers_with_kl(self, n_hypers, eps=None):
    """(theta, kl_hypers) from ONE fused launch (+ one for the backward); same draw and same values as
    `sample_hypers` followed by `kl_hypers`."""
    if self.map_est:
      return self.sample_hypers(n_hypers, eps), self.kl_hypers()
    if eps is None:
      eps = torch.empty((n_hypers,) + self.log_mean.shape, dtype=self.log_mean.dtype,
                        device=self.log_mean.device).normal_()
    from .functional import HyperFn
    return HyperFn.apply(self.log_mean, self.log_logvar, self.prior_log_mean, self.prior_log_logvar, eps)

  def kl_hypers(self):
    """sum_{D+1} KL(N(m_q, s_q^2) || N(m_p, s_p^2))   (var_gp/kernels.py:70-77; 785 elements: stays PyTorch)."""
    if self.map_est:
      return torch.tensor(0.0, device=self.log_mean.device)
    s_q = self.log_logvar.exp().sqrt()
    s_p = self.prior_log_logvar.exp().sqrt()
    var_ratio = (s_q / s_p).pow(2)
    t1 = ((self.log_mean - self.prior_log_mean) / s_p).pow(2)
    return (0.5 * (var_ratio + t1 - 1 - var_ratio.log())).sum(dim=0)


class DeepRBFKernel(RBFKernel):
  """var_gp/kernels.py:80-96: 784 -> 256 -> 256 -> 64 MLP, then the RBF kernel on the features."""

  def __init__(self, in_size, feature_size=64, **kwargs):
    super().__init__(feature_size, **kwargs)
    self.phi = nn.Sequential(
      nn.Linear(in_size, 256), nn.ReLU(),
      nn.Linear(256, 256), nn.ReLU(),
      nn.Linear(256, feature_size),
    )

  def features(self, x):
    return self.phi(x)


def rbf_compute(theta, x, y=None):
  """Functional form of RBFKernel.compute for arbitrary leading batch dims."""
  batch = x.shape[:-2]
  Pa, D = x.shape[-2:]
  x3 = x.reshape(-1, Pa, D)
  if y is None:
    K = RbfComputeFn.apply(theta, x3, None, False)
    return K.reshape(theta.shape[0], *batch, Pa, Pa)
  Pb = y.shape[-2]
  if tuple(y.shape[:-2]) != tuple(batch):
    y = y.expand(*batch, Pb, D)
  shared = _is_broadcast_batch(y) or x3.shape[0] == 1
  if shared:
    y3 = y[(0,) * len(batch)].unsqueeze(0) if len(batch) else y.unsqueeze(0)
  else:
    y3 = y.reshape(-1, Pb, D)
  K = RbfComputeFn.apply(theta, x3, y3, shared)
  return K.reshape(theta.shape[0], *batch, Pa, Pb)
