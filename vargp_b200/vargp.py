"""Drop-in for ``var_gp/vargp.py``: the VAR-GP model (same constructor, ``forward`` / ``loss`` /
``predict`` / ``create_clf``, parameter names and state-dict layout as the reference).

``loss`` and ``predict`` run the fused schedule of ``elbo.py`` on libvargp_sm100.so: one Cholesky + one
triangular inverse per (hyper sample, class) instead of the reference's t+2 nested Choleskys, batched
GEMMs for everything else, a single fused likelihood pass, and a hand-derived backward.  The block-diagonal
ablation (``ep_var_mean=False``) and ``forward(x, loss_cache=dict)`` go through the composable ops of
``vargp_b200.gp_utils`` (same kernels, autograd-composed).

RNG parity: draws are issued with the same calls, shapes and order as the reference (theta -> u_<t ->
likelihood), so on the same device and seed both consume the global generator identically
(SURVEY.md section 8c).  All three can also be pinned explicitly through ``noise=dict(eps_theta, eps_u, eps_f)``.
"""
import torch
import torch.nn as nn

from .functional import MarginalFn, TrilUnpackFn
from .gp_utils import vec2tril, mat2trilvec
from .kernels import RBFKernel, DeepRBFKernel
from .likelihoods import MulticlassSoftmax


class CholeskyError(torch.linalg.LinAlgError):
  pass


class VARGP(nn.Module):
  def __init__(self, z_init, kernel, likelihood, n_var_samples=1, ep_var_mean=True, prev_params=None):
    super().__init__()
    self.var_mean_mask = float(ep_var_mean)
    self.M = z_init.size(-2)
    self.kernel = kernel
    self.n_v = n_var_samples
    self.likelihood = likelihood

    # previous tasks' variational parameters are constants of this task (var_gp/vargp.py:17-20);
    # registered as non-persistent buffers so .to(device) moves them and the state dict keeps the
    # reference's keys {z, u_mean, u_tril_vec, kernel.*}.
    prev_params = list(prev_params or [])
    self.n_prev = len(prev_params)
    if self.n_prev:
      for p in prev_params:
        if p['z'].size(-2) != self.M:
          raise ValueError('all tasks must use the same number of inducing points per class')
      self.register_buffer('prev_z', torch.cat([p['z'].detach() for p in prev_params], dim=-2), persistent=False)
      self.register_buffer('prev_u_mean', torch.stack([p['u_mean'].detach().squeeze(-1) for p in prev_params]),
                           persistent=False)
      self.register_buffer('prev_u_tril_vec', torch.stack([p['u_tril_vec'].detach() for p in prev_params]),
                           persistent=False)
    self._prev_Lu = None          # (t, C, M, M), unpacked lazily on the compute device

    self.z = nn.Parameter(z_init.detach())
    out_size = self.z.size(0)
    self.u_mean = nn.Parameter(torch.Tensor(out_size, self.M, 1).normal_(0., .5))
    self.u_tril_vec = nn.Parameter(mat2trilvec(torch.eye(self.M).unsqueeze(0).expand(out_size, -1, -1)))

    self.factor_shard = None      # elbo.FactorShard while a data-parallel training step shards the O(P^3) work
    self.sync_errors = True       # raise LinAlgError from loss()/predict() like torch.cholesky would
    self._last_info = None

  # -- reference-compatible view of the previous tasks -----------------------------------------
  @property
  def prev_params(self):
    out = []
    for s in range(self.n_prev):
      sl = slice(s * self.M, (s + 1) * self.M)
      out.append(dict(z=self.prev_z[:, sl], u_mean=self.prev_u_mean[s].unsqueeze(-1),
                      u_tril=vec2tril(self.prev_u_tril_vec[s], self.M)))
    return out

  def _apply(self, fn, *args, **kwargs):
    self._prev_Lu = None
    return super()._apply(fn, *args, **kwargs)

  def _prev_factors(self):
    if self._prev_Lu is None or self._prev_Lu.device != self.z.device:
      t, C, T = self.prev_u_tril_vec.shape
      with torch.no_grad():
        self._prev_Lu = TrilUnpackFn.apply(self.prev_u_tril_vec.reshape(t * C, T), self.M).view(t, C, self.M, self.M)
    return self._prev_Lu

  # -- fused path -----------------------------------------------------------------------------
  def _assemble(self):
    """Zcat (C, P, D) inputs of the kernel (features for DKL), m_all (S, C, M), Lu_all (S, C, M, M)."""
    Lu_t = TrilUnpackFn.apply(self.u_tril_vec, self.M)
    if self.n_prev:
      Zcat = torch.cat([self.prev_z, self.z], dim=-2)
      m_all = torch.cat([self.prev_u_mean, self.u_mean.squeeze(-1).unsqueeze(0)], dim=0)
      Lu_all = torch.cat([self._prev_factors(), Lu_t.unsqueeze(0)], dim=0)
    else:
      Zcat, m_all, Lu_all = self.z, self.u_mean.squeeze(-1).unsqueeze(0), Lu_t.unsqueeze(0)
    return self.kernel.features(Zcat), m_all, Lu_all

  def _marginal(self, x, theta, want_kl):
    Zf, m_all, Lu_all = self._assemble()
    xf = self.kernel.features(x)
    f_mean, f_var, kl, info, L = MarginalFn.apply(theta, Zf, xf, m_all, Lu_all, self.M, want_kl, self.factor_shard)
    self._last_info = info
    if self.sync_errors:
      self.check_errors()
    return f_mean, f_var, kl, L

  def check_errors(self):
    """Raise torch.linalg.LinAlgError if the last Cholesky hit a non-positive pivot (device sync)."""
    if self._last_info is not None:
      bad = int(self._last_info.max().item())
      if bad:
        raise CholeskyError(f'linalg.cholesky: the input is not positive-definite '
                            f'(leading minor of order {bad} is not positive-definite)')

  # -- reference API --------------------------------------------------------------------------
  def compute_q(self, theta, cache=None):
    """Autoregressive variational distributions q(u_<t | theta), q(u_<=t | theta)   (var_gp/vargp.py:35-88).
    theta (n_hypers, D+1) -> mu_lt, S_lt, mu_leq_t, S_leq_t, z_leq_t; `cache` receives Lz_lt, Lz_lt_Kz_lt_z_t.
    Reference-order composed path (the training step never materialises these: elbo.py works in whitened space)."""
    if not self.n_prev:
      raise ValueError('compute_q needs prev_params (the reference indexes prev_params[0], vargp.py:52)')
    from .composed import _compute_q
    return _compute_q(self, theta, cache=cache)

  def compute_pf_diag(self, theta, x, mu_leq_t, S_leq_t, z_leq_t, cache=None):
    """Diagonal of p(f) = int p(f | u_<=t) q(u_<=t)   (var_gp/vargp.py:90-113) -> f_mean, f_var (n_hypers, C, B)."""
    from .composed import _compute_pf_diag
    return _compute_pf_diag(self, theta, x, mu_leq_t, S_leq_t, z_leq_t, cache=cache)

  def forward(self, x, loss_cache=False, noise=None):
    """x (B, in_size) -> pred_mu, pred_var (n_hypers, out_size, B)   (var_gp/vargp.py:115-175).

    With ``loss_cache`` a dict it is filled with var_mu_t, var_L_cov_t, prior_mu_t, prior_L_cov_t like the
    reference (composed path, consumes the u_<t draw)."""
    noise = noise or {}
    theta = self.kernel.sample_hypers(self.n_v, eps=noise.get('eps_theta'))
    if isinstance(loss_cache, dict):
      from .composed import forward_with_cache
      return forward_with_cache(self, x, theta, loss_cache, noise)
    f_mean, f_var, _, _ = self._marginal(x, theta, want_kl=False)
    return f_mean, f_var

  def loss(self, x, y, noise=None):
    """-> (kl_hypers, kl_u, nll)   (var_gp/vargp.py:177-194)."""
    noise = noise or {}
    if self.var_mean_mask != 1.0:
      from .composed import loss_composed
      return loss_composed(self, x, y, noise)
    theta, kl_hypers = self.kernel.sample_hypers_with_kl(self.n_v, eps=noise.get('eps_theta'))
    f_mean, f_var, kl_u, _ = self._marginal(x, theta, want_kl=True)
    if self.n_prev and 'eps_u' not in noise:
      # the reference draws u_<t here (vargp.py:138); with ep_var_mean=True the KL does not depend on it
      #  -- the draw is still issued so the generator stays in step with the reference
      torch.empty((self.n_v, theta.size(0), self.z.size(0), self.n_prev * self.M),
                  dtype=self.z.dtype, device=self.z.device).normal_()
    nll = self.likelihood.loss(f_mean, f_var, y, eps=noise.get('eps_f'))
    return kl_hypers, kl_u, nll

  def predict(self, x, noise=None):
    """-> class probabilities (B, out_size)   (var_gp/vargp.py:196-198)."""
    noise = noise or {}
    pred_mu, pred_var = self(x, noise=noise)
    return self.likelihood.predict(pred_mu, pred_var, eps=noise.get('eps_f'))

  @staticmethod
  def create_clf(dataset, M=20, n_f=10, n_var_samples=3, prev_params=None,
                 ep_var_mean=True, map_est_hypers=False, dkl=False):
    """var_gp/vargp.py:200-243.  Unlike the reference, the caller's `prev_params` dicts are not mutated."""
    N = len(dataset)
    out_size = torch.unique(dataset.targets).size(0)
    # inducing points initialised at random data points, an independent subset per class
    z = torch.stack([dataset[torch.randperm(N)[:M]][0] for _ in range(out_size)])

    prior_log_mean, prior_log_logvar, phi_params = None, None, None
    if prev_params:
      # hyper-prior := previous task's hyper-posterior
      prior_log_mean = prev_params[-1].get('kernel.log_mean')
      prior_log_logvar = prev_params[-1].get('kernel.log_logvar')
      if dkl:
        phi_params = {k[11:]: v for k, v in prev_params[-1].items() if k.startswith('kernel.phi.')}
      prev_params = [{k: v for k, v in p.items() if not k.startswith('kernel')} for p in prev_params]

    if dkl:
      kernel = DeepRBFKernel(z.size(-1), prior_log_mean=prior_log_mean,
                             prior_log_logvar=prior_log_logvar, map_est=map_est_hypers)
      if phi_params is not None:
        kernel.phi.load_state_dict(phi_params)
    else:
      kernel = RBFKernel(z.size(-1), prior_log_mean=prior_log_mean,
                         prior_log_logvar=prior_log_logvar, map_est=map_est_hypers)
    likelihood = MulticlassSoftmax(n_f=n_f)
    return VARGP(z, kernel, likelihood, n_var_samples=n_var_samples,
                 ep_var_mean=ep_var_mean, prev_params=prev_params)
