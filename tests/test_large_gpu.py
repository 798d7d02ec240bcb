"""GPU parity at the BASELINE configs' full sizes -- the large-shape dispatch paths of the product -- against fp64
outputs of the LIVE reference stored by tests/golden/make_large.py:

  large_split_t4      the benched step (Split-MNIST shape, t=4: P=300, B=512)
  large_permuted_t9   Permuted-MNIST shape at t=9 (M=100, P=1000, B=512): the 2-CTA tcgen05 kernel (gemm_tc2) and an
                      8-block factorisation
  large_scaled_slice  one slice of the scaled synthetic config (P=M=2048, t=0, a 2048-point minibatch shard)

Tolerances are the north star's, un-relaxed: ELBO terms and the five gradients 1e-4 (norm-relative), predictive
probabilities 1e-5 absolute, predictive mean / variance 1e-4.  Reference semantics: var_gp/vargp.py:177-198.
"""
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def _run(name):
  rec = util.load_golden(name)
  params, prev, x, y, noise, n_v, F, flags = util.case_tensors(rec['case'], torch.float32)
  gp = util.build_model(params, prev, n_v, F, flags, 'cuda', torch.float32)
  return rec, gp, x, y, noise


@pytest.mark.parametrize('name', util.large_names())
def test_large_loss_and_grads_match_reference(name, cuda_ops):
  rec, gp, x, y, noise = _run(name)
  r64 = rec['f64']
  tc2_0 = cuda_ops.tc2_launch_count()
  terms, grads = util.run_model(gp, x, y, noise, rec['beta'], rec['Ntot'])
  gp.check_errors()
  if rec['case']['M'] * (rec['case']['t'] + 1) >= 1000:
    assert cuda_ops.tc2_launch_count() > tc2_0, 'the 2-CTA kernel was expected on this dispatch path'
  for k in ('kl_hypers', 'kl_u', 'nll', 'total'):
    err = util.relerr(terms[k], r64[k])
    assert err < 1e-4, f'{name} {k}: {err:.3e}'
  for k in util.GRAD_KEYS:
    err = util.compressed_err(grads[k], r64['grads'][k])
    assert err < 1e-4, f'{name} grad {k}: {err:.3e}'


@pytest.mark.parametrize('name', util.large_names())
def test_large_predict_matches_reference(name, cuda_ops):
  rec, gp, x, y, noise = _run(name)
  r64 = rec['f64']
  nz = {k: v.cuda() for k, v in noise.items()}
  with torch.no_grad():
    probs = gp.predict(x.cuda(), noise=nz)
    mu, var = gp(x.cuda(), noise=nz)
  gp.check_errors()
  err = (probs.double().cpu() - r64['probs']).abs().max().item()
  assert err < 1e-5, f'{name} probs: {err:.3e}'
  assert util.relerr(mu, r64['f_mean']) < 1e-4
  assert util.relerr(var, r64['f_var']) < 1e-4
