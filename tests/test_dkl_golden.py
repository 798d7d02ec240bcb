"""DeepRBFKernel path (var_gp/kernels.py:80-96) against fixtures recorded from the LIVE
reference (tests/golden/dkl_*.pt, made by tests/golden/make_golden.py: reference outputs in fp32 / fp64 plus the MLP
weights): the product's fused host schedule through the `features` hook, on the torch emulation of the kernel
interface in fp64, must reproduce the ELBO terms, every gradient (variational parameters, hypers, MLP) and predict()."""
import os
import sys

import pytest
import torch

from tests import util

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
import make_golden as mg   # noqa: E402  (case construction only; the reference itself is not imported)

NAMES = sorted(f[:-3] for f in os.listdir(util.GOLDEN_DIR) if f.startswith('dkl_') and f.endswith('.pt'))


def _run(name, device, dtype):
  from vargp_b200.vargp import VARGP
  from vargp_b200.kernels import DeepRBFKernel
  from vargp_b200.likelihoods import MulticlassSoftmax
  rec = util.load_golden(name)
  kw = rec['case']
  params, prev, x, y, noise, phi_sd = mg.dkl_case(kw, dtype)
  for k, v in rec['phi'].items():                  # the recorded weights are the ones the reference ran with
    assert torch.equal(v, phi_sd[k].float())
  gp = mg.build_dkl_model(VARGP, DeepRBFKernel, MulticlassSoftmax, kw, params, prev, phi_sd, dtype).to(device)
  ref = rec['f64']
  nz = {k: v.to(device) for k, v in noise.items()}
  kl_h, kl_u, nll = gp.loss(x.to(device), y.to(device), noise=nz)
  total = ref['beta'] * kl_h + kl_u + (ref['Ntot'] / x.size(0)) * nll
  gp.zero_grad()
  total.backward()
  grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in gp.named_parameters()}
  with torch.no_grad():
    probs = gp.predict(x.to(device), noise=nz)
  return rec, dict(kl_hypers=kl_h, kl_u=kl_u, nll=nll, total=total), grads, probs


@pytest.mark.parametrize('name', NAMES)
def test_dkl_host_schedule_matches_reference_fp64(name, emu_ops):
  rec, terms, grads, probs = _run(name, 'cpu', torch.float64)
  ref = rec['f64']
  for k, v in terms.items():
    assert util.relerr(v, ref[k]) < 1e-9, k
  assert set(grads) == set(ref['grads'])
  scale = max(v.norm().item() for v in ref['grads'].values())
  for k, v in ref['grads'].items():
    # the bias of the last MLP layer has an exactly zero gradient (the RBF kernel is shift invariant): judged on scale
    err = ((grads[k] - v).norm() / max(v.norm().item(), 1e-6 * scale)).item()
    assert err < 1e-6, (k, err)
  assert (probs - ref['probs']).abs().max().item() < 1e-9
