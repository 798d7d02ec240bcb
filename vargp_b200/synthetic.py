"""Seed-pinned synthetic VAR-GP problems at the shapes of BASELINE.json: inputs
x ~ U[0,1]^D (optionally 19 %-sparse like MNIST), previous-task variational parameters, kernel hypers in the
learned-lengthscale regime.  Pure data generation (CPU torch.Generator streams, stable across machines);
used by bench.py, the tests and the golden-fixture generator."""
import math

import torch


def make_case(C, D, M, t, B, H=3, F=10, seed=0, sigma=10., dtype=torch.float32, with_eps_u=True,
              sparse=False, n_v=None):
  """Seed-pinned synthetic problem.  Everything is drawn in fp64 from one CPU generator and cast, so
  fp32 and fp64 cases see the same numbers.  H = number of hyper samples actually drawn (1 under
  map_est), n_v = n_var_samples of the model (defaults to H).  Returns (params, prev, x, y, noise)."""
  n_v = H if n_v is None else n_v
  g = torch.Generator().manual_seed(seed)
  U = lambda *s: torch.rand(*s, generator=g, dtype=torch.float64)
  Nrm = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
  T = M * (M + 1) // 2
  prev = [dict(z=U(C, M, D).to(dtype), u_mean=(0.5 * Nrm(C, M, 1)).to(dtype),
               u_tril_vec=(0.1 * Nrm(C, T)).to(dtype)) for _ in range(t)]
  log_mean = torch.cat([math.log(sigma) + 0.05 * Nrm(D), torch.full((1,), -0.7, dtype=torch.float64)])
  params = dict(z=U(C, M, D).to(dtype), u_mean=(0.5 * Nrm(C, M, 1)).to(dtype),
                u_tril_vec=(0.1 * Nrm(C, T)).to(dtype),
                log_mean=log_mean.to(dtype), log_logvar=(-2. + 0.1 * Nrm(D + 1)).to(dtype),
                prior_log_mean=(log_mean + 0.1 * Nrm(D + 1)).to(dtype),
                prior_log_logvar=(-1.5 + 0.1 * Nrm(D + 1)).to(dtype))
  x = U(B, D)
  if sparse:
    x = x * (U(B, D) < 0.19)
  x = x.to(dtype)
  y = torch.randint(0, C, (B,), generator=g)
  noise = dict(eps_theta=Nrm(H, D + 1).to(dtype), eps_f=Nrm(H, F, C, B).to(dtype))
  if t > 0 and with_eps_u:
    noise['eps_u'] = Nrm(n_v, H, C, t * M).to(dtype)
  return params, prev, x, y, noise


def make_retrain_case(C, D, M, t, B, H=3, F=10, seed=0, sigma=10., dtype=torch.float32, sparse=False):
  """`make_case` plus what the VARGPRetrain ablation needs (var_gp/vargp_retrain.py): trainable copies of the
  previous tasks' parameters that have moved away from the frozen posteriors, and the two extra draws of its loss
  (eps_q for u_<=t ~ q, eps_p for u~_<t ~ p(. | u_<=t)).  Returns (params, retrain, prev, x, y, noise)."""
  params, prev, x, y, noise = make_case(C, D, M, t, B, H=H, F=F, seed=seed, sigma=sigma, dtype=dtype,
                                        with_eps_u=False, sparse=sparse)
  g = torch.Generator().manual_seed(seed + 1000)
  Nrm = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
  retrain = [{k: (v.double() + 0.05 * Nrm(*v.shape)).to(dtype) for k, v in p.items()} for p in prev]
  if t > 0:
    noise['eps_q'] = Nrm(H, H, C, (t + 1) * M).to(dtype)
    noise['eps_p'] = Nrm(H, H, H, C, t * M).to(dtype)
  return params, retrain, prev, x, y, noise


def toy_data(N_K=50, K=4):
  """The reference's 4-class 2-D toy problem (var_gp/datasets.py:21-51, BASELINE configs[0]): issues the same global-RNG
  draws in the same order (six randn columns, then a 2-D Gaussian for class 3), so under the same torch.manual_seed it
  returns the same (X (4 N_K, 2), Y (4 N_K,)) as ToyDataset()._init_data.  Feed it to train.TensorTask."""
  col = lambda m, s: m + s * torch.randn(N_K, 1)
  X1 = torch.cat([col(0.8, 0.4), col(1.5, 0.4)], dim=-1)
  X2 = torch.cat([col(0.5, 0.6), col(-0.2, -0.1)], dim=-1)
  X3 = torch.cat([col(2.5, -0.1), col(1.0, 0.6)], dim=-1)
  # MultivariateNormal(mean, covariance_matrix=S).sample([N_K]) = mean + eps L^T, eps ~ N(0, I) of shape (N_K, 2)
  L = torch.linalg.cholesky(torch.tensor([[0.2, 0.1], [0.1, 0.1]]))
  # (the draw goes through the same helper as MultivariateNormal.rsample: on CPU it is not the randn stream)
  from torch.distributions.utils import _standard_normal
  eps = _standard_normal(torch.Size([N_K, 2]), dtype=L.dtype, device=L.device)
  X4 = torch.tensor([-0.5, 1.5]) + (L @ eps.unsqueeze(-1)).squeeze(-1)
  X = torch.cat([X1, X2, X3, X4], dim=0)
  X[:, 1] -= 1
  X[:, 0] -= 0.5
  Y = torch.arange(4).repeat_interleave(N_K)
  return X, Y
