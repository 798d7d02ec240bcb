"""CPU oracle for the VAR-GP ELBO hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  Nothing under ``vargp_b200/`` does.

What it is: a plain-PyTorch, CPU, functional restatement of the reference algorithm, op for op
(including the reference's wasteful parts: the ``B x B`` Gram that is built only for its diagonal,
the ``t+2`` nested Choleskys, the ``n_v``-fold redundant prior Cholesky), so that
  (a) it can be timed as "the reference's own CPU implementation" of the path, and
  (b) it is the arbiter for parity (run it in fp64 for the arbiter, fp32 for the like-for-like check).
All noise is passed in explicitly (``noise`` dict) instead of being drawn from the global RNG.

Pinning: the reference ships no tests / golden vectors ("parity unpinned" there).
This oracle is therefore pinned against the *live reference* executed in the build container:
``tests/golden/make_golden.py`` imports ``/root/reference/var_gp`` with patched RNG, runs the seeded
cases, and writes ``tests/golden/*.pt``; ``tests/test_oracle_golden.py`` checks this file against those
fixtures on every CPU test run.

Reference citations are ``file:line`` relative to ``/root/reference``.
"""
import math

import torch

JITTER = 1e-4  # var_gp/gp_utils.py:5 (default eps of `cholesky`)


# ----------------------------------------------------------------------------------------------
# var_gp/kernels.py
# ----------------------------------------------------------------------------------------------
def rbf_compute(theta, x, y=None):
  """ARD-RBF Gram, var_gp/kernels.py:24-56.  theta (H, D+1); x (..., M, D); y (..., N, D) or None.

  Returns (H, ..., M, N).  Follows the reference's three-Gram construction (xx, yy, xy and the two
  diagonals) so rounding matches; `y is None` re-uses xx for all three (diagonal is exactly gamma^2).
  """
  n_h = theta.size(0)
  th = theta.reshape(n_h, 1, *([1] * (x.dim() - 2)), theta.size(-1))
  ell = th[..., :-1].exp()
  amp2 = (2. * th[..., -1:]).exp()

  xs = x.unsqueeze(0) / ell
  g_xx = xs @ xs.transpose(-1, -2)
  if y is None:
    g_yy = g_xy = g_xx
  else:
    ys = y.unsqueeze(0) / ell
    g_yy = ys @ ys.transpose(-1, -2)
    g_xy = xs @ ys.transpose(-1, -2)
  d2 = -2. * g_xy + g_xx.diagonal(dim1=-2, dim2=-1).unsqueeze(-1) \
       + g_yy.diagonal(dim1=-2, dim2=-1).unsqueeze(-2)
  return amp2 * (-.5 * d2).exp()


def rbf_diag(theta):
  """var_gp/kernels.py:58-60 -> (H, 1, 1)."""
  return (2. * theta[..., -1:]).exp().unsqueeze(-2)


def sample_hypers(log_mean, log_logvar, eps_theta):
  """var_gp/kernels.py:62-68 with the N(0,1) draw made explicit: eps_theta (H, D+1)."""
  return log_mean + log_logvar.exp().sqrt() * eps_theta


def kl_hypers(log_mean, log_logvar, prior_log_mean, prior_log_logvar):
  """var_gp/kernels.py:70-77: sum over D+1 of KL(N(m_q, s_q^2) || N(m_p, s_p^2))
  (closed form of torch.distributions.kl._kl_normal_normal)."""
  s_q = log_logvar.exp().sqrt()
  s_p = prior_log_logvar.exp().sqrt()
  var_ratio = (s_q / s_p).pow(2)
  t1 = ((log_mean - prior_log_mean) / s_p).pow(2)
  return (0.5 * (var_ratio + t1 - 1. - var_ratio.log())).sum(0)


# ----------------------------------------------------------------------------------------------
# var_gp/gp_utils.py
# ----------------------------------------------------------------------------------------------
def chol_jitter(mat, eps=JITTER):
  """var_gp/gp_utils.py:5-11."""
  eye = torch.eye(mat.size(-1), dtype=mat.dtype)
  return torch.linalg.cholesky(mat + eps * eye)


def llt(L):
  """var_gp/gp_utils.py:14-19."""
  return L @ L.transpose(-1, -2)


def vec2tril(vec, m=None):
  """var_gp/gp_utils.py:22-49: packed row-major lower triangle -> matrix, softplus on the diagonal."""
  if m is None:
    m = int((math.sqrt(8. * vec.size(-1) + 1.) - 1.) / 2.)
  rows, cols = torch.tril_indices(m, m)
  out = torch.zeros(*vec.shape[:-1], m, m, dtype=vec.dtype)
  out[..., rows, cols] = vec
  on_diag = torch.eye(m, dtype=torch.bool)
  return torch.where(on_diag, torch.nn.functional.softplus(out), out)


def mat2trilvec(mat):
  """var_gp/gp_utils.py:52-65."""
  rows, cols = torch.tril_indices(mat.size(-1), mat.size(-1))
  return mat[..., rows, cols]


def _lsolve(L, rhs):
  """left, lower, no-transpose triangular solve (torch.triangular_solve(rhs, L, upper=False))."""
  shape = torch.broadcast_shapes(L.shape[:-2], rhs.shape[:-2])
  return torch.linalg.solve_triangular(L.expand(*shape, *L.shape[-2:]),
                                       rhs.expand(*shape, *rhs.shape[-2:]), upper=False)


def _atb(a, b):
  """einsum('...ij,...ik->...jk')."""
  return a.transpose(-1, -2) @ b


def gp_cond(u, Kzz, Kzx, Kxx, Lz=None, Lz_Kzx=None):
  """var_gp/gp_utils.py:68-98."""
  if Lz is None:
    Lz = chol_jitter(Kzz)
  Lz_u = _lsolve(Lz, u)
  if Lz_Kzx is None:
    Lz_Kzx = _lsolve(Lz, Kzx)
  mu = _atb(Lz_Kzx, Lz_u)
  cov = Kxx - _atb(Lz_Kzx, Lz_Kzx)
  return mu, cov


def linear_joint(m, S, Kzx, Kzz, V, b, cache=None):
  """var_gp/gp_utils.py:101-147."""
  Lz = chol_jitter(Kzz)
  Lz_m = _lsolve(Lz, m)
  Lz_Kzx = _lsolve(Lz, Kzx)
  Am = _atb(Lz_Kzx, Lz_m)
  Lz_S = _lsolve(Lz, S)
  AS = _atb(Lz_Kzx, Lz_S)
  SAt = AS.transpose(-1, -2)
  Lz_SAt = _lsolve(Lz, SAt)
  ASAt = _atb(Lz_SAt, Lz_Kzx)
  mu = torch.cat([m, Am + b], dim=-2)
  cov = torch.cat([torch.cat([S, SAt], dim=-1),
                   torch.cat([AS, V + ASAt], dim=-1)], dim=-2)
  if isinstance(cache, dict):
    cache.update(Lz_Kzx=Lz_Kzx, Lz=Lz)
  return mu, cov


def linear_marginal_diag(m, S, Kzz, Kzx, Kxx_diag, cache=None):
  """var_gp/gp_utils.py:150-191."""
  Lz = chol_jitter(Kzz)
  Lz_m = _lsolve(Lz, m)
  Lz_Kzx = _lsolve(Lz, Kzx)
  mu = _atb(Lz_Kzx, Lz_m).squeeze(-1)
  d1 = Lz_Kzx.pow(2).sum(dim=-2)
  Lz_LS = _lsolve(Lz, chol_jitter(S))
  d2 = _atb(Lz_LS, Lz_Kzx).pow(2).sum(dim=-2)
  var = Kxx_diag - d1 + d2
  if isinstance(cache, dict):
    cache.update(Lz=Lz, Lz_Kzx=Lz_Kzx)
  return mu, var


# ----------------------------------------------------------------------------------------------
# var_gp/likelihoods.py (MulticlassSoftmax)
# ----------------------------------------------------------------------------------------------
def softmax_samples(mu, var, eps_f):
  """var_gp/likelihoods.py:13-31 with explicit eps_f (H, F, C, B)."""
  f = mu.unsqueeze(1) + var.sqrt().unsqueeze(1) * eps_f
  return torch.log_softmax(f, dim=-2)


def softmax_nll(mu, var, y, eps_f):
  """var_gp/likelihoods.py:33-47: sum_b mean_h mean_f -log p(y_b)."""
  logp = softmax_samples(mu, var, eps_f)                       # (H, F, C, B)
  idx = y.view(1, 1, 1, -1).expand(logp.size(0), logp.size(1), 1, -1)
  picked = logp.gather(2, idx).squeeze(2)                      # (H, F, B)
  return -(picked.mean(1).mean(0)).sum(0)


def softmax_predict(mu, var, eps_f):
  """var_gp/likelihoods.py:49-63 -> (B, C)."""
  logp = softmax_samples(mu, var, eps_f)
  flat = logp.reshape(-1, *mu.shape[-2:])
  return (flat.logsumexp(dim=0).exp() / flat.size(0)).T


# ----------------------------------------------------------------------------------------------
# torch.distributions pieces the reference leans on (vargp.py:137-138, 182-190)
# ----------------------------------------------------------------------------------------------
def mvn_kl_tril(mu_q, L_q, mu_p, L_p):
  """KL(N(mu_q, L_q L_q^T) || N(mu_p, L_p L_p^T)); restates
  torch.distributions.kl._kl_multivariatenormal_multivariatenormal (torch 2.11)."""
  n = mu_q.size(-1)
  half_logdet = L_p.diagonal(dim1=-2, dim2=-1).log().sum(-1) - L_q.diagonal(dim1=-2, dim2=-1).log().sum(-1)
  tr = _lsolve(L_p, L_q).pow(2).sum((-2, -1))
  diff = (mu_p - mu_q).unsqueeze(-1)
  maha = _lsolve(L_p, diff).pow(2).sum((-2, -1))
  return half_logdet + 0.5 * (tr + maha - n)


# ----------------------------------------------------------------------------------------------
# var_gp/vargp.py
# ----------------------------------------------------------------------------------------------
def compute_q(theta, prev, z, u_mean, u_tril_vec, cache=None):
  """var_gp/vargp.py:35-88.  `prev` = list of dict(z, u_mean, u_tril) with u_tril already unpacked."""
  H = theta.size(0)
  z_lt = prev[0]['z']
  mu_lt = prev[0]['u_mean'].unsqueeze(0).expand(H, -1, -1, -1)
  S_lt = llt(prev[0]['u_tril']).unsqueeze(0).expand(H, -1, -1, -1)
  for p in prev[1:]:
    Kzx = rbf_compute(theta, z_lt, p['z'])
    Kzz = rbf_compute(theta, z_lt)
    V = llt(p['u_tril']).unsqueeze(0).expand(H, -1, -1, -1)
    b = p['u_mean'].unsqueeze(0).expand(H, -1, -1, -1)
    mu_lt, S_lt = linear_joint(mu_lt, S_lt, Kzx, Kzz, V, b)
    z_lt = torch.cat([z_lt, p['z']], dim=-2)
  Kzx = rbf_compute(theta, z_lt, z)
  Kzz = rbf_compute(theta, z_lt)
  V = llt(vec2tril(u_tril_vec)).unsqueeze(0).expand(H, -1, -1, -1)
  b = u_mean.unsqueeze(0).expand(H, -1, -1, -1)
  c = dict()
  mu_leq, S_leq = linear_joint(mu_lt, S_lt, Kzx, Kzz, V, b, cache=c)
  z_leq = torch.cat([z_lt, z], dim=-2)
  if isinstance(cache, dict):
    cache['Lz_lt'] = c['Lz']
    cache['Lz_lt_Kz_lt_z_t'] = c['Lz_Kzx']
  return mu_lt, S_lt, mu_leq, S_leq, z_leq


def compute_pf_diag(theta, x, mu_leq, S_leq, z_leq, cache=None):
  """var_gp/vargp.py:90-113."""
  xf = x.unsqueeze(0).expand(z_leq.size(0), -1, -1)
  Kzz = rbf_compute(theta, z_leq)
  Kzx = rbf_compute(theta, z_leq, xf)
  return linear_marginal_diag(mu_leq, S_leq, Kzz, Kzx, rbf_diag(theta), cache=cache)


def forward(params, prev, x, noise, n_v, ep_var_mean=True, want_loss_cache=False, map_est=False):
  """var_gp/vargp.py:115-175.

  params: dict(z (C,M,D), u_mean (C,M,1), u_tril_vec (C,M(M+1)/2), log_mean, log_logvar)
  prev:   list of dict(z, u_mean, u_tril_vec) for tasks < t (constants)
  noise:  dict(eps_theta (H,D+1), eps_u (n_v,H,C,Q) if t>0 and want_loss_cache, eps_f (H,F,C,B))
  Returns f_mean, f_var (H,C,B) and the loss cache (or None).
  """
  z, u_mean, u_tril_vec = params['z'], params['u_mean'], params['u_tril_vec']
  M = z.size(-2)
  if map_est:
    theta = params['log_mean'].unsqueeze(0)
  else:
    theta = sample_hypers(params['log_mean'], params['log_logvar'], noise['eps_theta'])
  prevp = [dict(z=p['z'], u_mean=p['u_mean'], u_tril=vec2tril(p['u_tril_vec'])) for p in prev]
  lc = None
  if prevp:
    cq = dict()
    mu_lt, S_lt, mu_leq, S_leq, z_leq = compute_q(theta, prevp, z, u_mean, u_tril_vec, cache=cq)
    f_mean, f_var = compute_pf_diag(theta, x, mu_leq, S_leq, z_leq)
    if want_loss_cache:
      # MultivariateNormal(mu, covariance_matrix=S).rsample([n_v]): no-jitter Cholesky (vargp.py:137-138)
      L_S = torch.linalg.cholesky(S_lt)
      u_lt = mu_lt.squeeze(-1).unsqueeze(0) + (L_S.unsqueeze(0) @ noise['eps_u'].unsqueeze(-1)).squeeze(-1)
      u_lt = u_lt.unsqueeze(-1)                                     # (n_v, H, C, Q, 1)
      Lz = cq['Lz_lt'].unsqueeze(0)
      Lz_Kzx = cq['Lz_lt_Kz_lt_z_t'].unsqueeze(0).expand(n_v, *([-1] * (Lz.dim() - 1)))
      Kzz_t = rbf_compute(theta, z).unsqueeze(0)
      prior_mu, prior_cov = gp_cond(u_lt, None, None, Kzz_t, Lz=Lz, Lz_Kzx=Lz_Kzx)
      var_mu = prior_mu * float(ep_var_mean) + u_mean.unsqueeze(0).unsqueeze(0)
      var_L = vec2tril(u_tril_vec, M).unsqueeze(0).unsqueeze(0)
      lc = dict(var_mu_t=var_mu.squeeze(-1), var_L_cov_t=var_L,
                prior_mu_t=prior_mu.squeeze(-1), prior_L_cov_t=chol_jitter(prior_cov))
  else:
    cpf = dict()
    L_u = vec2tril(u_tril_vec, M)
    f_mean, f_var = compute_pf_diag(theta, x, u_mean, llt(L_u), z, cache=cpf)
    if want_loss_cache:
      mu_t = u_mean.squeeze(-1).unsqueeze(0).unsqueeze(0)
      lc = dict(var_mu_t=mu_t, var_L_cov_t=L_u.unsqueeze(0).unsqueeze(0),
                prior_mu_t=torch.zeros_like(mu_t), prior_L_cov_t=cpf['Lz'].unsqueeze(0))
  return f_mean, f_var, lc


def elbo_terms(params, prev, x, y, noise, n_v, ep_var_mean=True, map_est=False):
  """var_gp/vargp.py:177-194 -> (kl_hypers, kl_u, nll)."""
  f_mean, f_var, lc = forward(params, prev, x, noise, n_v, ep_var_mean, want_loss_cache=True, map_est=map_est)
  nll = softmax_nll(f_mean, f_var, y, noise['eps_f'])
  kl = mvn_kl_tril(lc['var_mu_t'], lc['var_L_cov_t'], lc['prior_mu_t'], lc['prior_L_cov_t'])
  kl_u = kl.sum(-1).mean(0).mean(0)
  if map_est:
    kl_h = torch.zeros((), dtype=f_mean.dtype)
  else:
    kl_h = kl_hypers(params['log_mean'], params['log_logvar'],
                     params['prior_log_mean'], params['prior_log_logvar'])
  return kl_h, kl_u, nll


def predict(params, prev, x, noise, n_v, map_est=False):
  """var_gp/vargp.py:196-198 -> probs (B, C)."""
  f_mean, f_var, _ = forward(params, prev, x, noise, n_v, want_loss_cache=False, map_est=map_est)
  return softmax_predict(f_mean, f_var, noise['eps_f'])


# seeded synthetic cases live with the product's data utilities
from vargp_b200.synthetic import make_case  # noqa: E402,F401


# ----------------------------------------------------------------------------------------------
# var_gp/vargp_retrain.py  (ablation: every previous task's variational parameters are re-trained)
# ----------------------------------------------------------------------------------------------
LOG_2PI = math.log(2. * math.pi)


def mvn_log_prob_tril(value, mu, L):
  """torch.distributions.MultivariateNormal(mu, scale_tril=L).log_prob(value) (torch 2.11):
  -(n log 2pi + |L^-1 (value - mu)|^2) / 2 - sum log diag L."""
  diff = (value - mu).unsqueeze(-1)
  maha = _lsolve(L, diff).pow(2).sum((-2, -1))
  return -0.5 * (mu.size(-1) * LOG_2PI + maha) - L.diagonal(dim1=-2, dim2=-1).log().sum(-1)


def retrain_compute_q(theta, chain, z, u_mean, u_tril_vec):
  """var_gp/vargp_retrain.py:40-95: the autoregressive joint over `chain` (tasks < t, dicts with the PACKED
  u_tril_vec) followed by the current task.  Returns mu_lt, S_lt, mu_leq, S_leq, z_lt, z_leq."""
  H = theta.size(0)
  z_lt = chain[0]['z']
  mu_lt = chain[0]['u_mean'].unsqueeze(0).expand(H, -1, -1, -1)
  S_lt = llt(vec2tril(chain[0]['u_tril_vec'])).unsqueeze(0).expand(H, -1, -1, -1)
  for p in chain[1:]:
    Kzx = rbf_compute(theta, z_lt, p['z'])
    Kzz = rbf_compute(theta, z_lt)
    V = llt(vec2tril(p['u_tril_vec'])).unsqueeze(0).expand(H, -1, -1, -1)
    b = p['u_mean'].unsqueeze(0).expand(H, -1, -1, -1)
    mu_lt, S_lt = linear_joint(mu_lt, S_lt, Kzx, Kzz, V, b)
    z_lt = torch.cat([z_lt, p['z']], dim=-2)
  Kzx = rbf_compute(theta, z_lt, z)
  Kzz = rbf_compute(theta, z_lt)
  V = llt(vec2tril(u_tril_vec)).unsqueeze(0).expand(H, -1, -1, -1)
  b = u_mean.unsqueeze(0).expand(H, -1, -1, -1)
  mu_leq, S_leq = linear_joint(mu_lt, S_lt, Kzx, Kzz, V, b)
  return mu_lt, S_lt, mu_leq, S_leq, z_lt, torch.cat([z_lt, z], dim=-2)


def retrain_elbo_terms(params, retrain, prev, x, y, noise, n_v):
  """var_gp/vargp_retrain.py:126-237 -> (kl_hypers, kl_u, nll).

  params:  current task (z, u_mean, u_tril_vec, log_mean, log_logvar, prior_*)
  retrain: trainable copies of the previous tasks' (z, u_mean, u_tril_vec)   [self.retrain_params]
  prev:    the frozen previous posteriors                                    [self.prev_params]
  noise:   eps_theta (H, D+1); if prev: eps_q (n_v, H, C, P), eps_p (n_v, n_v, H, C, Q); eps_f (H, F, C, B)
           -- the reference's draw order: hypers, q_leq_t.sample, p_lt_tilde.sample, likelihood.
  """
  theta = sample_hypers(params['log_mean'], params['log_logvar'], noise['eps_theta'])
  z, u_mean, u_tril_vec = params['z'], params['u_mean'], params['u_tril_vec']
  M = z.size(-2)
  kl_h = kl_hypers(params['log_mean'], params['log_logvar'], params['prior_log_mean'], params['prior_log_logvar'])
  if not prev:
    cpf = dict()
    L_u = vec2tril(u_tril_vec, M)
    f_mean, f_var = compute_pf_diag(theta, x, u_mean, llt(L_u), z, cache=cpf)
    nll = softmax_nll(f_mean, f_var, y, noise['eps_f'])
    mu_t = u_mean.squeeze(-1).unsqueeze(0).unsqueeze(0)
    kl = mvn_kl_tril(mu_t, L_u.unsqueeze(0).unsqueeze(0), torch.zeros_like(mu_t), cpf['Lz'].unsqueeze(0))
    return kl_h, kl.sum(-1).mean(0).mean(0), nll

  _, _, mu_leq, S_leq, _, z_leq = retrain_compute_q(theta, retrain, z, u_mean, u_tril_vec)
  f_mean, f_var = compute_pf_diag(theta, x, mu_leq, S_leq, z_leq)
  # p(u_<=t | theta), q(u~_<t | theta) from the FROZEN posteriors, p(u~_<t | theta)       [:147-157]
  prior_S_leq = rbf_compute(theta, z_leq)
  mu_tl, S_tl, _, _, z_tl, _ = retrain_compute_q(theta, prev, z, u_mean, u_tril_vec)
  prior_S_tl = rbf_compute(theta, z_tl)
  # u_<=t ~ q (non-reparametrised .sample), u~_<t ~ p(u~_<t | u_<=t, theta)                [:159-169]
  L_q = chol_jitter(S_leq)
  with torch.no_grad():
    u_leq = (mu_leq.squeeze(-1).unsqueeze(0) + (L_q.unsqueeze(0) @ noise['eps_q'].unsqueeze(-1)).squeeze(-1)).unsqueeze(-1)
    Kzz = rbf_compute(theta, z_leq).unsqueeze(0).expand(n_v, -1, -1, -1, -1)
    Kzx = rbf_compute(theta, z_leq, z_tl).unsqueeze(0).expand(n_v, -1, -1, -1, -1)
    Kxx = rbf_compute(theta, z_tl).unsqueeze(0).expand(n_v, -1, -1, -1, -1)
    p_mu, p_S = gp_cond(u_leq, Kzz, Kzx, Kxx)
    L_p = chol_jitter(p_S)
    u_tl = p_mu.squeeze(-1).unsqueeze(0) + (L_p.unsqueeze(0) @ noise['eps_p'].unsqueeze(-1)).squeeze(-1)
  nll = softmax_nll(f_mean, f_var, y, noise['eps_f'])
  # loss                                                                                   [:197-222]
  kl = mvn_kl_tril(mu_leq.squeeze(-1), L_q, torch.zeros_like(mu_leq.squeeze(-1)), chol_jitter(prior_S_leq))
  kl_u = kl.sum(-1).mean(0)
  mu_tl = mu_tl.squeeze(-1)
  ratio = mvn_log_prob_tril(u_tl, torch.zeros_like(mu_tl), chol_jitter(prior_S_tl)) \
      - mvn_log_prob_tril(u_tl, mu_tl, chol_jitter(S_tl))
  return kl_h, kl_u + ratio.sum(-1).mean(-1).mean(-1).mean(-1), nll


def retrain_predict(params, retrain, x, noise):
  """var_gp/vargp_retrain.py:239-241 -> probs (B, C)."""
  theta = sample_hypers(params['log_mean'], params['log_logvar'], noise['eps_theta'])
  z, u_mean, u_tril_vec = params['z'], params['u_mean'], params['u_tril_vec']
  if retrain:
    _, _, mu_leq, S_leq, _, z_leq = retrain_compute_q(theta, retrain, z, u_mean, u_tril_vec)
    f_mean, f_var = compute_pf_diag(theta, x, mu_leq, S_leq, z_leq)
  else:
    f_mean, f_var = compute_pf_diag(theta, x, u_mean, llt(vec2tril(u_tril_vec, z.size(-2))), z)
  return softmax_predict(f_mean, f_var, noise['eps_f'])
