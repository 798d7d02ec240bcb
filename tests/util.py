"""Shared test helpers: build the product model from an oracle case, load golden fixtures, compare."""
import os

import torch

from oracle import vargp_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
MODEL_FLAGS = ('ep_var_mean', 'map_est')
GRAD_KEYS = ('z', 'u_mean', 'u_tril_vec', 'log_mean', 'log_logvar')


def golden_names():
  """VARGP fixtures (make_golden.CASES)."""
  return sorted(f[:-3] for f in os.listdir(GOLDEN_DIR) if f.endswith('.pt') and not f.startswith(('retrain_', 'dkl_', 'data_', 'large_')))


def retrain_names():
  """VARGPRetrain fixtures (make_golden.RETRAIN_CASES)."""
  return sorted(f[:-3] for f in os.listdir(GOLDEN_DIR) if f.endswith('.pt') and f.startswith('retrain_'))


def load_golden(name):
  return torch.load(os.path.join(GOLDEN_DIR, name + '.pt'), weights_only=False)


def split_case(kw):
  kw = dict(kw)
  flags = {k: kw.pop(k) for k in MODEL_FLAGS if k in kw}
  return kw, flags


def case_tensors(kw, dtype):
  kw, flags = split_case(kw)
  params, prev, x, y, noise = orc.make_case(dtype=dtype, **kw)
  H, F = kw.get('H', 3), kw.get('F', 10)
  return params, prev, x, y, noise, kw.get('n_v', H), F, flags


def build_model(params, prev, n_v, F, flags, device, dtype):
  """vargp_b200.VARGP carrying exactly the oracle case's parameters."""
  from vargp_b200.vargp import VARGP
  from vargp_b200.kernels import RBFKernel
  from vargp_b200.likelihoods import MulticlassSoftmax
  D = params['z'].size(-1)
  kern = RBFKernel(D, prior_log_mean=params['prior_log_mean'].clone(),
                   prior_log_logvar=params['prior_log_logvar'].clone(), map_est=flags.get('map_est', False))
  gp = VARGP(params['z'].clone(), kern, MulticlassSoftmax(n_f=F), n_var_samples=n_v,
             ep_var_mean=flags.get('ep_var_mean', True),
             prev_params=[{k: v.clone() for k, v in p.items()} for p in prev])
  gp = gp.to(dtype)
  with torch.no_grad():
    gp.u_mean.copy_(params['u_mean'])
    gp.u_tril_vec.copy_(params['u_tril_vec'])
    gp.kernel.log_mean.copy_(params['log_mean'])
    gp.kernel.log_logvar.copy_(params['log_logvar'])
  return gp.to(device)


def model_grads(gp):
  z = lambda p: torch.zeros_like(p) if p.grad is None else p.grad
  return dict(z=z(gp.z), u_mean=z(gp.u_mean), u_tril_vec=z(gp.u_tril_vec),
              log_mean=z(gp.kernel.log_mean), log_logvar=z(gp.kernel.log_logvar))


def relerr(a, b):
  """norm-relative error with an absolute floor (element-wise rtol is meaningless on
  near-zero gradient entries)."""
  a, b = a.detach().double().cpu(), b.detach().double().cpu()
  return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def run_model(gp, x, y, noise, beta, Ntot):
  """One ELBO evaluation + backward with pinned noise; returns terms, grads."""
  dev = gp.z.device
  nz = {k: v.to(dev) for k, v in noise.items()}
  kl_h, kl_u, nll = gp.loss(x.to(dev), y.to(dev), noise=nz)
  total = beta * kl_h + kl_u + (Ntot / x.size(0)) * nll
  gp.zero_grad()
  total.backward()
  return dict(kl_hypers=kl_h, kl_u=kl_u, nll=nll, total=total), model_grads(gp)


def build_retrain_model(params, retrain, prev, n_v, F, device, dtype):
  """vargp_b200.VARGPRetrain carrying a make_retrain_case: trainable copies = `retrain`, frozen posteriors = `prev`."""
  from vargp_b200.vargp_retrain import VARGPRetrain
  from vargp_b200.kernels import RBFKernel
  from vargp_b200.likelihoods import MulticlassSoftmax
  D = params['z'].size(-1)
  kern = RBFKernel(D, prior_log_mean=params['prior_log_mean'].clone(), prior_log_logvar=params['prior_log_logvar'].clone())
  gp = VARGPRetrain(params['z'].clone(), kern, MulticlassSoftmax(n_f=F), n_var_samples=n_v,
                    prev_params=[{k: v.clone() for k, v in p.items()} for p in prev]).to(dtype)
  with torch.no_grad():
    gp.u_mean.copy_(params['u_mean'])
    gp.u_tril_vec.copy_(params['u_tril_vec'])
    gp.kernel.log_mean.copy_(params['log_mean'])
    gp.kernel.log_logvar.copy_(params['log_logvar'])
    for s, p in enumerate(retrain):
      for k, v in p.items():
        gp.retrain_params[s][k].copy_(v)
  return gp.to(device)


def run_retrain_model(gp, x, y, noise, beta, Ntot):
  dev = gp.z.device
  nz = {k: v.to(dev) for k, v in noise.items()}
  kl_h, kl_u, nll = gp.loss(x.to(dev), y.to(dev), noise=nz)
  total = beta * kl_h + kl_u + (Ntot / x.size(0)) * nll
  gp.zero_grad()
  total.backward()
  grads = model_grads(gp)
  for s in range(gp.n_prev):
    for k in ('z', 'u_mean', 'u_tril_vec'):
      grads[f'retrain.{s}.{k}'] = gp.retrain_params[s][k].grad
  return dict(kl_hypers=kl_h, kl_u=kl_u, nll=nll, total=total), grads


# ------------------------------------------------------------------------------------------------------
# full-size fixtures (tests/golden/make_large.py): big tensors are stored as norm + strided subsample + projections
# ------------------------------------------------------------------------------------------------------
def large_names():
  return sorted(f[:-3] for f in os.listdir(GOLDEN_DIR) if f.endswith('.pt') and f.startswith('large_'))


def compressed_err(actual, ref):
  """Error of `actual` against a reference stored by make_large.compress: the norm-relative error for a plain
  tensor; for a compressed one the max of (a) the norm-relative error over the stored strided subsample, (b) the
  relative difference of the norms and (c) the projection residuals |s.(a - b)| / (3 |b|) over the stored random-sign
  vectors s (for an error vector e, s.e ~ N(0, |e|^2): an error of norm tol*|b| ANYWHERE in the tensor shows up as
  residuals of that size; the factor 3 keeps a correct tensor from failing on the tail of that distribution)."""
  if not isinstance(ref, dict):
    return relerr(actual, ref)
  from tests.golden.make_large import project
  flat = actual.detach().double().cpu().reshape(-1)
  assert tuple(actual.shape) == tuple(ref['shape']), (tuple(actual.shape), ref['shape'])
  sub = flat[::ref['stride']]
  e_sub = ((sub - ref['sub']).norm() / ref['sub'].norm().clamp_min(1e-300)).item()
  e_norm = abs(flat.norm().item() - ref['norm']) / max(ref['norm'], 1e-300)
  e_proj = ((project(flat, ref['seed']) - ref['proj']).abs().max() / (3.0 * max(ref['norm'], 1e-300))).item()
  return max(e_sub, e_norm, e_proj)
