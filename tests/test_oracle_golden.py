"""The CPU oracle must reproduce the reference's outputs recorded in tests/golden (pins the oracle)."""
import pytest
import torch

from oracle import vargp_oracle as orc
from tests import util


@pytest.mark.parametrize('name', util.golden_names())
@pytest.mark.parametrize('tag,dtype,tol', [('f64', torch.float64, 1e-9), ('f32', torch.float32, 1e-4)])
def test_oracle_matches_reference_fixture(name, tag, dtype, tol):
  rec = util.load_golden(name)
  params, prev, x, y, noise, n_v, F, flags = util.case_tensors(rec['case'], dtype)
  ref = rec[tag]
  leaf = {k: (v.clone().requires_grad_(True) if k in util.GRAD_KEYS else v) for k, v in params.items()}
  kl_h, kl_u, nll = orc.elbo_terms(leaf, prev, x, y, noise, n_v=n_v, **flags)
  total = ref['beta'] * kl_h + kl_u + (ref['Ntot'] / x.size(0)) * nll
  total.backward()
  assert util.relerr(kl_u, ref['kl_u']) < tol
  assert util.relerr(nll, ref['nll']) < tol
  assert util.relerr(kl_h, ref['kl_hypers']) < tol or flags.get('map_est')
  for k in util.GRAD_KEYS:
    g = leaf[k].grad if leaf[k].grad is not None else torch.zeros_like(leaf[k])
    if ref['grads'][k].abs().max() == 0:
      assert g.abs().max() == 0
    else:
      assert util.relerr(g, ref['grads'][k]) < tol, k
  probs = orc.predict(params, prev, x, noise, n_v=n_v, map_est=flags.get('map_est', False))
  assert (probs.double() - ref['probs'].double()).abs().max().item() < (1e-10 if dtype == torch.float64 else 1e-5)
