"""Autograd-composed variant of the VAR-GP forward / loss, built from the differentiable primitives of
``vargp_b200.gp_utils`` (same kernels as the fused path, but one launch group per reference op).

Used where the fused schedule of ``elbo.py`` does not apply:
  * ``VARGP.forward(x, loss_cache=dict)``: the reference's loss-cache protocol (keys var_mu_t, var_L_cov_t,
    prior_mu_t, prior_L_cov_t), which needs the sampled u_<t;
  * the block-diagonal ablation ``ep_var_mean=False``, whose KL depends on that sample.
It follows the reference's op order (var_gp/vargp.py:35-194) rather than the whitened single-Cholesky
algebra, so it doubles as an on-device cross-check of the fused path.
"""
import torch

from .gp_utils import (cholesky, rev_cholesky, vec2tril, gp_cond, linear_joint, linear_marginal_diag,
                       tri_solve, matmul)


def _compute_q(model, theta, cache=None):
  """q(u_<t | theta) and q(u_<=t | theta) by folding `linear_joint` over the tasks (vargp.py:35-88)."""
  H = theta.size(0)
  prev = model.prev_params
  kern = model.kernel
  z_lt = prev[0]['z']
  mu_lt = prev[0]['u_mean'].unsqueeze(0).expand(H, -1, -1, -1)
  S_lt = rev_cholesky(prev[0]['u_tril']).unsqueeze(0).expand(H, -1, -1, -1)
  for p in prev[1:]:
    Kzx = kern.compute(theta, z_lt, p['z'])
    Kzz = kern.compute(theta, z_lt)
    V = rev_cholesky(p['u_tril']).unsqueeze(0).expand(H, -1, -1, -1)
    b = p['u_mean'].unsqueeze(0).expand(H, -1, -1, -1)
    mu_lt, S_lt = linear_joint(mu_lt, S_lt, Kzx, Kzz, V, b)
    z_lt = torch.cat([z_lt, p['z']], dim=-2)
  Kzx = kern.compute(theta, z_lt, model.z)
  Kzz = kern.compute(theta, z_lt)
  V = rev_cholesky(vec2tril(model.u_tril_vec, model.M)).unsqueeze(0).expand(H, -1, -1, -1)
  b = model.u_mean.unsqueeze(0).expand(H, -1, -1, -1)
  c = dict()
  mu_leq, S_leq = linear_joint(mu_lt, S_lt, Kzx, Kzz, V, b, cache=c)
  z_leq = torch.cat([z_lt, model.z], dim=-2)
  if isinstance(cache, dict):
    cache['Lz_lt'] = c['Lz']
    cache['Lz_lt_Kz_lt_z_t'] = c['Lz_Kzx']
  return mu_lt, S_lt, mu_leq, S_leq, z_leq


def _compute_pf_diag(model, theta, x, mu_leq, S_leq, z_leq, cache=None):
  """vargp.py:90-113."""
  xf = x.unsqueeze(0).expand(z_leq.size(0), -1, -1)
  Kzz = model.kernel.compute(theta, z_leq)
  Kzx = model.kernel.compute(theta, z_leq, xf)
  return linear_marginal_diag(mu_leq, S_leq, Kzz, Kzx, model.kernel.compute_diag(theta), cache=cache)


def forward_with_cache(model, x, theta, loss_cache, noise):
  """vargp.py:115-175 with `loss_cache` a dict."""
  n_v, M = model.n_v, model.M
  if model.n_prev:
    cq = dict()
    mu_lt, S_lt, mu_leq, S_leq, z_leq = _compute_q(model, theta, cache=cq)
    pred_mu, pred_var = _compute_pf_diag(model, theta, x, mu_leq, S_leq, z_leq)
    # MultivariateNormal(mu_<t, covariance_matrix=S_<t).rsample([n_v]): Cholesky WITHOUT jitter
    eps_u = noise.get('eps_u')
    if eps_u is None:
      eps_u = torch.empty((n_v,) + tuple(mu_lt.shape[:-1]), dtype=x.dtype, device=x.device).normal_()
    L_S = cholesky(S_lt, eps=0.)
    u_lt = mu_lt.unsqueeze(0) + matmul(L_S.unsqueeze(0), eps_u.unsqueeze(-1))      # (n_v, H, C, Q, 1)
    Lz = cq['Lz_lt'].unsqueeze(0)
    Lz_Kzx = cq['Lz_lt_Kz_lt_z_t'].unsqueeze(0).expand(n_v, *([-1] * (Lz.dim() - 1)))
    Kzz_t = model.kernel.compute(theta, model.z).unsqueeze(0)
    prior_mu, prior_cov = gp_cond(u_lt, None, None, Kzz_t, Lz=Lz, Lz_Kzx=Lz_Kzx)
    var_mu = prior_mu * model.var_mean_mask + model.u_mean.unsqueeze(0).unsqueeze(0)
    var_L = vec2tril(model.u_tril_vec, M).unsqueeze(0).unsqueeze(0)
    loss_cache.update(dict(var_mu_t=var_mu.squeeze(-1), var_L_cov_t=var_L,
                           prior_mu_t=prior_mu.squeeze(-1), prior_L_cov_t=cholesky(prior_cov)))
  else:
    cpf = dict()
    L_u = vec2tril(model.u_tril_vec, M)
    pred_mu, pred_var = _compute_pf_diag(model, theta, x, model.u_mean, rev_cholesky(L_u), model.z, cache=cpf)
    mu_t = model.u_mean.squeeze(-1).unsqueeze(0).unsqueeze(0)
    loss_cache.update(dict(var_mu_t=mu_t, var_L_cov_t=L_u.unsqueeze(0).unsqueeze(0),
                           prior_mu_t=torch.zeros_like(mu_t), prior_L_cov_t=cpf['Lz'].unsqueeze(0)))
  return pred_mu, pred_var


def mvn_kl(mu_q, L_q, mu_p, L_p):
  """KL(N(mu_q, L_q L_q^T) || N(mu_p, L_p L_p^T)) for lower-triangular factors (what
  torch.distributions.kl_divergence evaluates at vargp.py:182-190), solves on the library."""
  n = mu_q.size(-1)
  half_logdet = L_p.diagonal(dim1=-2, dim2=-1).log().sum(-1) - L_q.diagonal(dim1=-2, dim2=-1).log().sum(-1)
  tr = tri_solve(L_p, L_q).pow(2).sum((-2, -1))
  maha = tri_solve(L_p, (mu_p - mu_q).unsqueeze(-1)).pow(2).sum((-2, -1))
  return half_logdet + 0.5 * (tr + maha - n)


def loss_composed(model, x, y, noise):
  """vargp.py:177-194 on the composed path -> (kl_hypers, kl_u, nll)."""
  theta = model.kernel.sample_hypers(model.n_v, eps=noise.get('eps_theta'))
  lc = dict()
  pred_mu, pred_var = forward_with_cache(model, x, theta, lc, noise)
  nll = model.likelihood.loss(pred_mu, pred_var, y, eps=noise.get('eps_f'))
  kl = mvn_kl(lc['var_mu_t'], lc['var_L_cov_t'], lc['prior_mu_t'], lc['prior_L_cov_t'])
  kl_u = kl.sum(dim=-1).mean(dim=0).mean(dim=0)
  return model.kernel.kl_hypers(), kl_u, nll
