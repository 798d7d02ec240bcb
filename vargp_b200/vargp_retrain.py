"""Drop-in for ``var_gp/vargp_retrain.py``: the "retrain" ablation of VAR-GP, in which the variational parameters
of every previous task are trained again next to the current task's (same constructor, ``forward`` / ``loss`` /
``predict`` / ``create_clf``, parameter names and state-dict layout ``retrain_params.<s>.{z,u_mean,u_tril_vec}``).

It runs on the differentiable library primitives of ``vargp_b200.gp_utils`` (Cholesky, inverse-then-GEMM solves,
GEMMs, the RBF Gram with its hand-written adjoint) in the reference's op order, with the redundancy taken out:

* K(z_<=t, z_<=t) and K(z~_<t, z~_<t) are built once and shared by the prior factors, the conditional
  p(u~_<t | u_<=t) and the predictive marginal (the reference rebuilds each up to three times,
  vargp_retrain.py:111,149,162-164);
* the conditional is factored once per (hyper sample, class) instead of once per variational sample (the reference
  expands Kzz / Kzx / Kxx over n_v and factors n_v identical copies, vargp_retrain.py:162-165), and all samples
  ride through the solves as columns of one right-hand side;
* ``log_prob`` of the (n_v, n_v, H, C) samples is one solve with n_v^2 columns per (hyper sample, class).

RNG parity: the four draws are issued with the reference's shapes and order (hypers (H, D+1) -> q_leq_t.sample
(n_v, H, C, P) -> p_lt_tilde.sample (n_v, n_v, H, C, Q) -> likelihood (H, F, C, B)); each can be pinned through
``noise=dict(eps_theta, eps_q, eps_p, eps_f)``.

Deliberate difference: the reference wraps the tensors of ``prev_params`` into the trainable ``retrain_params``
without copying (vargp_retrain.py:16-24), so on CPU every optimizer step also moves the "frozen" posteriors; here
the frozen posteriors are separate buffers, which is what the derivation (and the reference on any other device)
does.  A single loss / gradient evaluation is identical either way.
"""
import math

import torch
import torch.nn as nn

from .composed import mvn_kl
from .gp_utils import cholesky, rev_cholesky, vec2tril, linear_joint, linear_marginal_diag, tri_solve, matmul
from .kernels import RBFKernel
from .likelihoods import MulticlassSoftmax

_KEYS = ('z', 'u_mean', 'u_tril_vec')


def _normal(shape, like):
  return torch.empty(shape, dtype=like.dtype, device=like.device).normal_()


def mvn_log_prob_cols(value, mu, L):
  """log N(value_s; mu, L L^T) for the columns of value (..., n, S); mu (..., n), L (..., n, n) lower -> (..., S).
  What MultivariateNormal(mu, scale_tril=L).log_prob evaluates (vargp_retrain.py:218), one solve for all samples."""
  n = mu.size(-1)
  maha = tri_solve(L, value - mu.unsqueeze(-1)).pow(2).sum(-2)
  return -0.5 * (n * math.log(2. * math.pi) + maha) - L.diagonal(dim1=-2, dim2=-1).log().sum(-1).unsqueeze(-1)


class VARGPRetrain(nn.Module):
  def __init__(self, z_init, kernel, likelihood, n_var_samples=1, prev_params=None):
    super().__init__()
    prev_params = list(prev_params or [])
    self.n_prev = len(prev_params)
    self.retrain_params = None
    if self.n_prev:
      # frozen posteriors q(u~_<t) (non-persistent buffers: .to(device) moves them, the state dict keeps the
      # reference's keys) and their trainable copies
      for k in _KEYS:
        for s, p in enumerate(prev_params):
          self.register_buffer(f'prev{s}_{k}', p[k].detach().clone(), persistent=False)
      self.retrain_params = nn.ModuleList([
        nn.ParameterDict({k: nn.Parameter(p[k].detach().clone()) for k in _KEYS}) for p in prev_params])

    self.M = z_init.size(-2)
    self.kernel = kernel
    self.n_v = n_var_samples
    self.likelihood = likelihood

    self.z = nn.Parameter(z_init.detach())
    out_size = self.z.size(0)
    self.u_mean = nn.Parameter(torch.Tensor(out_size, self.M, 1).normal_(0., .5))
    self.u_tril_vec = nn.Parameter(torch.ones(out_size, (self.M * (self.M + 1)) // 2))

  @property
  def prev_params(self):
    """The frozen previous posteriors as the reference's list of dicts."""
    return [{k: getattr(self, f'prev{s}_{k}') for k in _KEYS} for s in range(self.n_prev)]

  def compute_q(self, theta, prev_params, cache=None):
    """Autoregressive joint over `prev_params` followed by the current task (vargp_retrain.py:40-95).
    -> mu_lt, S_lt, mu_leq_t, S_leq_t, z_lt, z_leq_t."""
    H = theta.size(0)
    ex = lambda a: a.unsqueeze(0).expand(H, -1, -1, -1)
    z_lt = prev_params[0]['z']
    mu_lt = ex(prev_params[0]['u_mean'])
    S_lt = ex(rev_cholesky(vec2tril(prev_params[0]['u_tril_vec'])))
    for p in prev_params[1:]:
      Kzx = self.kernel.compute(theta, z_lt, p['z'])
      Kzz = self.kernel.compute(theta, z_lt)
      mu_lt, S_lt = linear_joint(mu_lt, S_lt, Kzx, Kzz, ex(rev_cholesky(vec2tril(p['u_tril_vec']))), ex(p['u_mean']))
      z_lt = torch.cat([z_lt, p['z']], dim=-2)
    Kzx = self.kernel.compute(theta, z_lt, self.z)
    Kzz = self.kernel.compute(theta, z_lt)
    c = dict()
    mu_leq_t, S_leq_t = linear_joint(mu_lt, S_lt, Kzx, Kzz, ex(rev_cholesky(vec2tril(self.u_tril_vec))),
                                     ex(self.u_mean), cache=c)
    z_leq_t = torch.cat([z_lt, self.z], dim=-2)
    if isinstance(cache, dict):
      cache['Lz_lt'] = c['Lz']
      cache['Lz_lt_Kz_lt_z_t'] = c['Lz_Kzx']
    return mu_lt, S_lt, mu_leq_t, S_leq_t, z_lt, z_leq_t

  def compute_pf_diag(self, theta, x, mu_leq_t, S_leq_t, z_leq_t, cache=None, Kzz=None):
    """Diagonal of p(f) = int p(f | u_<=t) q(u_<=t)   (vargp_retrain.py:97-123) -> f_mean, f_var (H, C, B)."""
    xf = x.unsqueeze(0).expand(z_leq_t.size(0), -1, -1)
    if Kzz is None:
      Kzz = self.kernel.compute(theta, z_leq_t)
    Kzx = self.kernel.compute(theta, z_leq_t, xf)
    return linear_marginal_diag(mu_leq_t, S_leq_t, Kzz, Kzx, self.kernel.compute_diag(theta), cache=cache)

  def forward(self, x, loss_cache=False, noise=None):
    """x (B, in_size) -> pred_mu, pred_var (n_hypers, out_size, B); fills `loss_cache` (a dict) with the reference's
    keys (vargp_retrain.py:126-195)."""
    noise = noise or {}
    n_v = self.n_v
    theta = self.kernel.sample_hypers(n_v, eps=noise.get('eps_theta'))
    if not self.n_prev:
      cpf = dict()
      L_u = vec2tril(self.u_tril_vec, self.M)
      pred_mu, pred_var = self.compute_pf_diag(theta, x, self.u_mean, rev_cholesky(L_u), self.z, cache=cpf)
      if isinstance(loss_cache, dict):
        mu_t = self.u_mean.squeeze(-1).unsqueeze(0).unsqueeze(0)
        loss_cache.update(dict(var_mu_t=mu_t, var_L_cov_t=L_u.unsqueeze(0).unsqueeze(0),
                               prior_mu_t=torch.zeros_like(mu_t), prior_L_cov_t=cpf.pop('Lz').unsqueeze(0)))
      return pred_mu, pred_var

    _, _, mu_leq_t, S_leq_t, _, z_leq_t = self.compute_q(theta, list(self.retrain_params))
    K_leq = self.kernel.compute(theta, z_leq_t)              # prior covariance of u_<=t, Kzz of both conditionals
    pred_mu, pred_var = self.compute_pf_diag(theta, x, mu_leq_t, S_leq_t, z_leq_t, Kzz=K_leq)
    if not isinstance(loss_cache, dict):
      return pred_mu, pred_var

    H, C, P = mu_leq_t.shape[:3]
    mu_lt_tilde, S_lt_tilde, _, _, z_lt_tilde, _ = self.compute_q(theta, self.prev_params)
    K_tilde = self.kernel.compute(theta, z_lt_tilde)         # prior covariance of u~_<t
    Q = z_lt_tilde.size(-2)
    L_q = cholesky(S_leq_t)
    L_prior = cholesky(K_leq)
    with torch.no_grad():
      # u_<=t ~ q(u_<=t | theta): n_v samples as columns                             [:159-160]
      eps_q = noise.get('eps_q')
      if eps_q is None:
        eps_q = _normal((n_v, H, C, P), x)
      u_leq = mu_leq_t + matmul(L_q, eps_q.permute(1, 2, 3, 0).contiguous())          # (H, C, P, n_v)
      # p(u~_<t | u_<=t, theta): one factorisation per (h, c), all samples at once   [:162-166]
      Kzx = self.kernel.compute(theta, z_leq_t, z_lt_tilde)
      Lz_Kzx = tri_solve(L_prior, Kzx)
      p_mu = matmul(Lz_Kzx.transpose(-1, -2), tri_solve(L_prior, u_leq))               # (H, C, Q, n_v)
      L_p = cholesky(K_tilde - matmul(Lz_Kzx.transpose(-1, -2), Lz_Kzx))
      eps_p = noise.get('eps_p')
      if eps_p is None:
        eps_p = _normal((n_v, n_v, H, C, Q), x)
      # column s * n_v + v holds sample s of the conditional given u_<=t sample v     [:168]
      u_tilde = (p_mu.unsqueeze(-2) + matmul(L_p, eps_p.permute(2, 3, 4, 0, 1).reshape(H, C, Q, n_v * n_v))
                 .view(H, C, Q, n_v, n_v)).reshape(H, C, Q, n_v * n_v)
    loss_cache.update(dict(
      var_mu_leq_t=mu_leq_t.squeeze(-1), var_L_leq_t=L_q,
      prior_mu_leq_t=torch.zeros_like(mu_leq_t.squeeze(-1)), prior_L_leq_t=L_prior,
      var_mu_lt_tilde=mu_lt_tilde.squeeze(-1), var_L_lt_tilde=cholesky(S_lt_tilde),
      prior_mu_lt_tilde=torch.zeros_like(mu_lt_tilde.squeeze(-1)), prior_L_lt_tilde=cholesky(K_tilde),
      u_lt_tilde=u_tilde.view(H, C, Q, n_v, n_v).permute(3, 4, 0, 1, 2)))             # (n_v, n_v, H, C, Q) view
    return pred_mu, pred_var

  def loss(self, x, y, noise=None):
    """-> (kl_hypers, kl_u, nll)   (vargp_retrain.py:197-237)."""
    noise = noise or {}
    lc = dict()
    pred_mu, pred_var = self(x, loss_cache=lc, noise=noise)
    nll = self.likelihood.loss(pred_mu, pred_var, y, eps=noise.get('eps_f'))
    if self.n_prev:
      kl_u = mvn_kl(lc['var_mu_leq_t'], lc['var_L_leq_t'], lc['prior_mu_leq_t'], lc['prior_L_leq_t']) \
          .sum(dim=-1).mean(dim=0)
      u = lc['u_lt_tilde']
      n_s, n_v, H, C, Q = u.shape
      cols = u.permute(2, 3, 4, 0, 1).reshape(H, C, Q, n_s * n_v)
      ratio = mvn_log_prob_cols(cols, lc['prior_mu_lt_tilde'], lc['prior_L_lt_tilde']) \
          - mvn_log_prob_cols(cols, lc['var_mu_lt_tilde'], lc['var_L_lt_tilde'])      # (H, C, n_s * n_v)
      kl_u = kl_u + ratio.sum(dim=1).mean()
    else:
      kl_u = mvn_kl(lc['var_mu_t'], lc['var_L_cov_t'], lc['prior_mu_t'], lc['prior_L_cov_t']) \
          .sum(dim=-1).mean(dim=0).mean(dim=0)
    return self.kernel.kl_hypers(), kl_u, nll

  def predict(self, x, noise=None):
    """-> class probabilities (B, out_size)   (vargp_retrain.py:239-241)."""
    noise = noise or {}
    pred_mu, pred_var = self(x, noise=noise)
    return self.likelihood.predict(pred_mu, pred_var, eps=noise.get('eps_f'))

  @staticmethod
  def create_clf(dataset, M=20, n_f=10, n_var_samples=3, prev_params=None):
    """vargp_retrain.py:243-267.  Unlike the reference, the caller's `prev_params` dicts are not mutated."""
    N = len(dataset)
    out_size = torch.unique(dataset.targets).size(0)
    z = torch.stack([dataset[torch.randperm(N)[:M]][0] for _ in range(out_size)])
    prior_log_mean, prior_log_logvar = None, None
    if prev_params:
      prior_log_mean = prev_params[-1].get('kernel.log_mean')
      prior_log_logvar = prev_params[-1].get('kernel.log_logvar')
      prev_params = [{k: v for k, v in p.items() if k in _KEYS} for p in prev_params]
    kernel = RBFKernel(z.size(-1), prior_log_mean=prior_log_mean, prior_log_logvar=prior_log_logvar)
    return VARGPRetrain(z, kernel, MulticlassSoftmax(n_f=n_f), n_var_samples=n_var_samples, prev_params=prev_params)
