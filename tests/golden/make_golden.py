"""Generate golden fixtures from the LIVE reference (build container only).

Usage (in the build container, where /root/reference exists):
    python tests/golden/make_golden.py

Imports ``/root/reference/var_gp`` (unmodified), pins every random draw of ``VARGP.loss`` /
``VARGP.predict`` to explicit tensors (draw order theta -> u_<t -> likelihood),
applies the value-preserving ``nll_loss(input.contiguous())`` shim torch 2.11 needs for backward, runs
the seeded cases of ``oracle.vargp_oracle.make_case`` in fp32 and fp64, and stores the reference's
outputs under ``tests/golden/*.pt``.  Inputs are NOT stored: they are re-derived from the seed by
``make_case`` (CPU torch.Generator streams are stable), which keeps the fixtures small.

It also asserts, on the spot, that the oracle restatement reproduces the reference.
The reference cannot travel to the GPU box; these fixtures can.
"""
import contextlib
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get('VARGP_REFERENCE', '/root/reference')

# name -> make_case kwargs (+ model flags)
CASES = {
  'toy_t0':        dict(C=4, D=2, M=20, t=0, B=100, sigma=0.25, seed=11),
  'toy_t1':        dict(C=4, D=2, M=20, t=1, B=100, sigma=0.25, seed=12),
  'mnist_t0':      dict(C=10, D=784, M=12, t=0, B=48, sigma=10., seed=21),
  'mnist_t1':      dict(C=10, D=784, M=12, t=1, B=48, sigma=10., seed=22),
  'mnist_t3':      dict(C=10, D=784, M=12, t=3, B=48, sigma=10., seed=23),
  'mnist_t2_blockdiag': dict(C=10, D=784, M=12, t=2, B=48, sigma=10., seed=24, ep_var_mean=False),
  'mnist_t1_sparse':    dict(C=10, D=784, M=12, t=1, B=48, sigma=10., seed=25, sparse=True),
  'mnist_t1_default_init': dict(C=10, D=784, M=12, t=1, B=48, sigma=0.5, seed=26),   # kernel underflows to 0
  'odd_t2':        dict(C=3, D=37, M=7, t=2, B=33, sigma=3., seed=31, H=2, F=5),      # ragged sizes
  'map_t1':        dict(C=4, D=16, M=9, t=1, B=21, sigma=2., seed=32, H=1, n_v=3, map_est=True),
}
MODEL_FLAGS = ('ep_var_mean', 'map_est')

# VARGPRetrain ablation (var_gp/vargp_retrain.py): name -> make_retrain_case kwargs; fixtures retrain_*.pt
RETRAIN_CASES = {
  'retrain_toy_t0':   dict(C=4, D=2, M=20, t=0, B=100, sigma=0.25, seed=41),
  'retrain_toy_t1':   dict(C=4, D=2, M=20, t=1, B=100, sigma=0.25, seed=42),
  'retrain_mnist_t1': dict(C=10, D=784, M=12, t=1, B=48, sigma=10., seed=43),
  'retrain_odd_t2':   dict(C=3, D=37, M=7, t=2, B=33, sigma=3., seed=44, H=2, F=5),
}


def load_reference():
  sys.path.insert(0, REF)
  warnings.filterwarnings('ignore')
  import torch.nn.functional as F
  import var_gp.likelihoods as lk

  class _F:
    def __getattr__(self, k):
      return getattr(F, k)

    @staticmethod
    def nll_loss(inp, tgt, **kw):
      return F.nll_loss(inp.contiguous(), tgt, **kw)

  lk.F = _F()
  from var_gp.vargp import VARGP
  from var_gp.kernels import RBFKernel
  from var_gp.likelihoods import MulticlassSoftmax
  return VARGP, RBFKernel, MulticlassSoftmax


@contextlib.contextmanager
def pinned_noise(draws):
  """Feed `draws` (list of tensors) to, in order, Normal.rsample / MVN.rsample / torch.randn."""
  import torch.distributions.normal as dn
  import torch.distributions.multivariate_normal as dm
  queue = list(draws)

  def pop(shape):
    t = queue.pop(0)
    assert tuple(t.shape) == tuple(shape), (tuple(t.shape), tuple(shape))
    return t.clone()

  o1, o2, o3 = dn._standard_normal, dm._standard_normal, torch.randn
  dn._standard_normal = lambda shape, dtype=None, device=None: pop(shape)
  dm._standard_normal = lambda shape, dtype=None, device=None: pop(shape)
  torch.randn = lambda *shape, **kw: pop(shape)
  try:
    yield
  finally:
    dn._standard_normal, dm._standard_normal, torch.randn = o1, o2, o3
  assert not queue, 'unused noise draws'


def build_reference_model(refmods, params, prev, H, F, ep_var_mean=True, map_est=False):
  VARGP, RBFKernel, MulticlassSoftmax = refmods
  D = params['z'].size(-1)
  kern = RBFKernel(D, prior_log_mean=params['prior_log_mean'].clone(),
                   prior_log_logvar=params['prior_log_logvar'].clone(), map_est=map_est)
  gp = VARGP(params['z'].clone(), kern, MulticlassSoftmax(n_f=F), n_var_samples=H,
             ep_var_mean=ep_var_mean, prev_params=[{k: v.clone() for k, v in p.items()} for p in prev])
  with torch.no_grad():
    gp.u_mean.copy_(params['u_mean'])
    gp.u_tril_vec.copy_(params['u_tril_vec'])
    gp.kernel.log_mean.copy_(params['log_mean'])
    gp.kernel.log_logvar.copy_(params['log_logvar'])
  return gp


def run_reference(refmods, kw, dtype):
  from oracle import vargp_oracle as orc
  kw = dict(kw)
  flags = {k: kw.pop(k) for k in MODEL_FLAGS if k in kw}
  H, F = kw.get('H', 3), kw.get('F', 10)
  n_v = kw.get('n_v', H)
  old = torch.get_default_dtype()
  torch.set_default_dtype(dtype)
  try:
    params, prev, x, y, noise = orc.make_case(dtype=dtype, **kw)
    gp = build_reference_model(refmods, params, prev, n_v, F, **flags)
    draws = [] if flags.get('map_est') else [noise['eps_theta']]
    if prev:
      draws.append(noise['eps_u'])
    draws.append(noise['eps_f'])
    with pinned_noise(draws):
      kl_h, kl_u, nll = gp.loss(x, y)
    Ntot = 10 * x.size(0)
    beta = 1.7
    total = beta * kl_h + kl_u + (Ntot / x.size(0)) * nll
    gp.zero_grad()
    total.backward()
    grads = dict(z=gp.z.grad, u_mean=gp.u_mean.grad, u_tril_vec=gp.u_tril_vec.grad,
                 log_mean=gp.kernel.log_mean.grad,
                 log_logvar=(gp.kernel.log_logvar.grad if gp.kernel.log_logvar.grad is not None
                             else torch.zeros_like(gp.kernel.log_logvar)))
    pdraws = ([] if flags.get('map_est') else [noise['eps_theta']]) + [noise['eps_f']]
    with torch.no_grad():
      with pinned_noise(pdraws):
        probs = gp.predict(x)
      with pinned_noise(pdraws):
        f_mean, f_var = gp(x)
        torch.randn(*noise['eps_f'].shape)   # consume the likelihood draw left in the queue
    out = dict(kl_hypers=kl_h.detach(), kl_u=kl_u.detach(), nll=nll.detach(), total=total.detach(),
               probs=probs, f_mean=f_mean, f_var=f_var, beta=beta, Ntot=Ntot,
               grads={k: v.detach().clone() for k, v in grads.items()})

    # --- the oracle restatement must reproduce the reference ---
    op = {k: (v.clone().requires_grad_(True) if k in grads else v) for k, v in params.items()}
    okl_h, okl_u, onll = orc.elbo_terms(op, prev, x, y, noise, n_v=n_v, **flags)
    ototal = beta * okl_h + okl_u + (Ntot / x.size(0)) * onll
    ototal.backward()
    tol = 5e-5 if dtype == torch.float32 else 1e-10
    def close(a, b, name):
      err = (a.double() - b.double()).abs().max().item()
      ref = b.double().abs().max().item()
      assert err <= tol * max(ref, 1e-30) + (1e-7 if dtype == torch.float32 else 1e-14), (name, err, ref)
    close(okl_h, kl_h, 'kl_h'); close(okl_u, kl_u, 'kl_u'); close(onll, nll, 'nll')
    for k in grads:
      og = op[k].grad if op[k].grad is not None else torch.zeros_like(op[k])
      close(og, grads[k], 'grad ' + k)
    oprobs = orc.predict(params, prev, x, noise, n_v=n_v, map_est=flags.get('map_est', False))
    close(oprobs, probs, 'probs')
    return out
  finally:
    torch.set_default_dtype(old)


def _close(a, b, name, dtype, tol32=5e-5):
  tol = tol32 if dtype == torch.float32 else 1e-10
  err = (a.double() - b.double()).abs().max().item()
  ref = b.double().abs().max().item()
  assert err <= tol * max(ref, 1e-30) + (1e-7 if dtype == torch.float32 else 1e-14), (name, err, ref)


def run_reference_retrain(kw, dtype):
  """The live VARGPRetrain (var_gp/vargp_retrain.py) on a seeded case; its four draws pinned in the reference's
  order (hypers, q_leq_t.sample, p_lt_tilde.sample, likelihood)."""
  from oracle import vargp_oracle as orc
  from vargp_b200.synthetic import make_retrain_case
  from var_gp.vargp_retrain import VARGPRetrain
  from var_gp.kernels import RBFKernel
  from var_gp.likelihoods import MulticlassSoftmax
  H, F = kw.get('H', 3), kw.get('F', 10)
  old = torch.get_default_dtype()
  torch.set_default_dtype(dtype)
  try:
    params, retrain, prev, x, y, noise = make_retrain_case(dtype=dtype, **kw)
    D = params['z'].size(-1)
    kern = RBFKernel(D, prior_log_mean=params['prior_log_mean'].clone(), prior_log_logvar=params['prior_log_logvar'].clone())
    # the constructor wraps the tensors it is given into the trainable retrain_params (sharing storage); the frozen
    # posteriors self.prev_params are then pointed at separate tensors so that the two sets differ
    gp = VARGPRetrain(params['z'].clone(), kern, MulticlassSoftmax(n_f=F), n_var_samples=H,
                      prev_params=[{k: v.clone() for k, v in p.items()} for p in retrain] or None)
    if prev:
      gp.prev_params = [{k: v.clone() for k, v in p.items()} for p in prev]
    with torch.no_grad():
      gp.u_mean.copy_(params['u_mean'])
      gp.u_tril_vec.copy_(params['u_tril_vec'])
      gp.kernel.log_mean.copy_(params['log_mean'])
      gp.kernel.log_logvar.copy_(params['log_logvar'])
    draws = [noise['eps_theta']] + ([noise['eps_q'], noise['eps_p']] if prev else []) + [noise['eps_f']]
    with pinned_noise(draws):
      kl_h, kl_u, nll = gp.loss(x, y)
    Ntot, beta = 10 * x.size(0), 1.7
    total = beta * kl_h + kl_u + (Ntot / x.size(0)) * nll
    gp.zero_grad()
    total.backward()
    grads = dict(z=gp.z.grad, u_mean=gp.u_mean.grad, u_tril_vec=gp.u_tril_vec.grad,
                 log_mean=gp.kernel.log_mean.grad, log_logvar=gp.kernel.log_logvar.grad)
    for s in range(len(prev)):
      for k in ('z', 'u_mean', 'u_tril_vec'):
        grads[f'retrain.{s}.{k}'] = gp.retrain_params[s][k].grad
    with torch.no_grad(), pinned_noise([noise['eps_theta'], noise['eps_f']]):
      probs = gp.predict(x)
    out = dict(kl_hypers=kl_h.detach(), kl_u=kl_u.detach(), nll=nll.detach(), total=total.detach(), probs=probs,
               beta=beta, Ntot=Ntot, grads={k: v.detach().clone() for k, v in grads.items()})

    # --- the oracle restatement must reproduce the reference ---
    leaf = lambda d: {k: v.clone().requires_grad_(True) if v.is_floating_point() and not k.startswith('prior') else v
                      for k, v in d.items()}
    op, ort = leaf(params), [leaf(p) for p in retrain]
    okl_h, okl_u, onll = orc.retrain_elbo_terms(op, ort, prev, x, y, noise, n_v=H)
    (beta * okl_h + okl_u + (Ntot / x.size(0)) * onll).backward()
    _close(okl_h, kl_h, 'kl_h', dtype); _close(okl_u, kl_u, 'kl_u', dtype, 2e-4); _close(onll, nll, 'nll', dtype)
    for k, g in grads.items():
      src = ort[int(k.split('.')[1])][k.split('.')[2]] if k.startswith('retrain') else op[k]
      _close(src.grad, g, 'grad ' + k, dtype, 2e-4)
    _close(orc.retrain_predict(params, retrain, x, noise), probs, 'probs', dtype)
    return out
  finally:
    torch.set_default_dtype(old)


# DeepRBFKernel (var_gp/kernels.py:80-96): name -> shapes; fixtures dkl_*.pt also carry the MLP weights
DKL_CASES = {
  'dkl_t1': dict(C=4, Din=20, Fd=16, M=9, t=1, B=48, H=2, F=5, seed=51),
  'dkl_t0': dict(C=3, Din=12, Fd=8, M=5, t=0, B=17, H=3, F=2, seed=52),
}


def dkl_case(kw, dtype):
  """Seeded DKL problem: hypers / variational parameters / noise from make_case at the FEATURE dimension, inputs and
  inducing inputs at the input dimension, MLP weights from torch.manual_seed(seed)."""
  import torch.nn as nn
  from vargp_b200.synthetic import make_case
  C, Din, Fd, M, t, B, H, F, seed = (kw[k] for k in ('C', 'Din', 'Fd', 'M', 't', 'B', 'H', 'F', 'seed'))
  params, prev, _, y, noise = make_case(C=C, D=Fd, M=M, t=t, B=B, H=H, F=F, sigma=1.5, seed=seed, dtype=dtype)
  g = torch.Generator().manual_seed(seed + 7)
  U = lambda *sh: torch.rand(*sh, generator=g, dtype=torch.float64).to(dtype)
  params = dict(params, z=U(C, M, Din))
  x = U(B, Din)
  prev = [dict(p, z=U(C, M, Din)) for p in prev]
  old = torch.get_default_dtype()
  torch.set_default_dtype(torch.float32)       # nn.Linear's init consumes the generator in a dtype-dependent way:
  try:                                         # always draw the weights in fp32 and cast
    torch.manual_seed(seed)
    phi = nn.Sequential(nn.Linear(Din, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, Fd))
  finally:
    torch.set_default_dtype(old)
  return params, prev, x, y, noise, {k: v.detach().to(dtype) for k, v in phi.state_dict().items()}


def build_dkl_model(VARGP, DeepRBFKernel, MulticlassSoftmax, kw, params, prev, phi_sd, dtype):
  kern = DeepRBFKernel(kw['Din'], feature_size=kw['Fd'], prior_log_mean=params['prior_log_mean'].clone(),
                       prior_log_logvar=params['prior_log_logvar'].clone())
  gp = VARGP(params['z'].clone(), kern, MulticlassSoftmax(n_f=kw['F']), n_var_samples=kw['H'],
             prev_params=[{k: v.clone() for k, v in p.items()} for p in prev]).to(dtype)
  gp.kernel.phi.load_state_dict(phi_sd)
  with torch.no_grad():
    gp.u_mean.copy_(params['u_mean']); gp.u_tril_vec.copy_(params['u_tril_vec'])
    gp.kernel.log_mean.copy_(params['log_mean']); gp.kernel.log_logvar.copy_(params['log_logvar'])
  return gp


def run_reference_dkl(kw, dtype):
  from var_gp.vargp import VARGP
  from var_gp.kernels import DeepRBFKernel
  from var_gp.likelihoods import MulticlassSoftmax
  old = torch.get_default_dtype()
  torch.set_default_dtype(dtype)
  try:
    params, prev, x, y, noise, phi_sd = dkl_case(kw, dtype)
    gp = build_dkl_model(VARGP, DeepRBFKernel, MulticlassSoftmax, kw, params, prev, phi_sd, dtype)
    draws = [noise['eps_theta']] + ([noise['eps_u']] if prev else []) + [noise['eps_f']]
    with pinned_noise(draws):
      kl_h, kl_u, nll = gp.loss(x, y)
    beta, Ntot = 1.7, 10 * x.size(0)
    total = beta * kl_h + kl_u + (Ntot / x.size(0)) * nll
    gp.zero_grad()
    total.backward()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)) for k, p in gp.named_parameters()}
    with torch.no_grad(), pinned_noise([noise['eps_theta'], noise['eps_f']]):
      probs = gp.predict(x)
    return dict(kl_hypers=kl_h.detach(), kl_u=kl_u.detach(), nll=nll.detach(), total=total.detach(), probs=probs,
                beta=beta, Ntot=Ntot, grads=grads)
  finally:
    torch.set_default_dtype(old)


def save_toy_data():
  """ToyDataset of the live reference (var_gp/datasets.py:11-51) under a fixed seed -> data_toy_seed3.pt."""
  from var_gp.datasets import ToyDataset
  torch.manual_seed(3)
  ds = ToyDataset(N_K=50)
  torch.save(dict(seed=3, N_K=50, data=ds.data.clone(), targets=ds.targets.clone()), os.path.join(HERE, 'data_toy_seed3.pt'))
  print('data_toy_seed3.pt', tuple(ds.data.shape))


def main():
  refmods = load_reference()
  save_toy_data()
  if os.environ.get('VARGP_GOLDEN_ONLY') == 'toydata':
    return
  for name, kw in DKL_CASES.items():
    rec = dict(case=kw, torch=torch.__version__)
    for dtype, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
      rec[tag] = run_reference_dkl(kw, dtype)
    rec['phi'] = dkl_case(kw, torch.float32)[5]
    path = os.path.join(HERE, name + '.pt')
    torch.save(rec, path)
    print(f'{name:24s} kl_u={rec["f64"]["kl_u"].item():.6f} nll={rec["f64"]["nll"].item():.6f} '
          f'-> {os.path.getsize(path) / 1024:.0f} KiB')
  if os.environ.get('VARGP_GOLDEN_ONLY') == 'dkl':
    return
  for name, kw in RETRAIN_CASES.items():
    rec = dict(case=kw, torch=torch.__version__)
    for dtype, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
      rec[tag] = run_reference_retrain(kw, dtype)
    path = os.path.join(HERE, name + '.pt')
    torch.save(rec, path)
    print(f'{name:24s} kl_u={rec["f64"]["kl_u"].item():.6f} nll={rec["f64"]["nll"].item():.6f} '
          f'-> {os.path.getsize(path) / 1024:.0f} KiB')
  if os.environ.get('VARGP_GOLDEN_ONLY') == 'retrain':
    return
  for name, kw in CASES.items():
    rec = dict(case=kw, torch=torch.__version__)
    for dtype, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
      rec[tag] = run_reference(refmods, kw, dtype)
    path = os.path.join(HERE, name + '.pt')
    torch.save(rec, path)
    print(f'{name:24s} kl_u={rec["f64"]["kl_u"].item():.6f} nll={rec["f64"]["nll"].item():.6f} '
          f'-> {os.path.getsize(path) / 1024:.0f} KiB')


if __name__ == '__main__':
  main()
