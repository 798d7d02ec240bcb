"""Host-schedule tests on CPU: the product model (vargp_b200.VARGP -> functional -> elbo.py) driven through
the torch emulation of the kernel interface must reproduce the reference fixtures.  This checks the
whitened single-Cholesky algebra, the hand-derived backward and the module plumbing; the CUDA kernels
themselves are checked against the same emulation contracts in the -m gpu tests."""
import pytest
import torch

from tests import util


@pytest.mark.parametrize('name', util.golden_names())
def test_fused_schedule_matches_reference_fp64(name, emu_ops):
  rec = util.load_golden(name)
  params, prev, x, y, noise, n_v, F, flags = util.case_tensors(rec['case'], torch.float64)
  ref = rec['f64']
  gp = util.build_model(params, prev, n_v, F, flags, 'cpu', torch.float64)
  terms, grads = util.run_model(gp, x, y, noise, ref['beta'], ref['Ntot'])
  for k in ('kl_hypers', 'kl_u', 'nll', 'total'):
    if ref[k].abs() > 0:
      assert util.relerr(terms[k], ref[k]) < 1e-8, k
  for k in util.GRAD_KEYS:
    if ref['grads'][k].abs().max() > 0:
      assert util.relerr(grads[k], ref['grads'][k]) < 1e-7, k
  with torch.no_grad():
    probs = gp.predict(x, noise=noise)
  assert (probs - ref['probs']).abs().max().item() < 1e-9


@pytest.mark.parametrize('name', ['mnist_t0', 'mnist_t3', 'odd_t2'])
def test_fused_schedule_fp32_within_tolerance(name, emu_ops):
  """fp32 arithmetic in the new op order stays within the north-star tolerance of the fp64 reference."""
  rec = util.load_golden(name)
  params, prev, x, y, noise, n_v, F, flags = util.case_tensors(rec['case'], torch.float32)
  ref = rec['f64']
  gp = util.build_model(params, prev, n_v, F, flags, 'cpu', torch.float32)
  terms, grads = util.run_model(gp, x, y, noise, ref['beta'], ref['Ntot'])
  for k in ('kl_u', 'nll', 'total'):
    assert util.relerr(terms[k], ref[k]) < 1e-4, k
  for k in util.GRAD_KEYS:
    assert util.relerr(grads[k], ref['grads'][k]) < 1e-4, k


def test_forward_loss_cache_protocol(emu_ops):
  """forward(x, loss_cache=dict) fills the reference's four keys with the reference's shapes."""
  rec = util.load_golden('odd_t2')
  params, prev, x, y, noise, n_v, F, flags = util.case_tensors(rec['case'], torch.float64)
  gp = util.build_model(params, prev, n_v, F, flags, 'cpu', torch.float64)
  lc = dict()
  mu, var = gp(x, loss_cache=lc, noise=noise)
  H, C, M, B = 2, 3, 7, 33
  assert tuple(mu.shape) == (H, C, B) and tuple(var.shape) == (H, C, B)
  assert tuple(lc['var_mu_t'].shape) == (n_v, H, C, M)
  assert tuple(lc['prior_L_cov_t'].shape) == (n_v, H, C, M, M)
  assert (mu - rec['f64']['f_mean']).abs().max() < 1e-9
  assert (var - rec['f64']['f_var']).abs().max() < 1e-9


def test_state_dict_keys_match_reference(emu_ops):
  rec = util.load_golden('odd_t2')
  params, prev, x, y, noise, n_v, F, flags = util.case_tensors(rec['case'], torch.float32)
  gp = util.build_model(params, prev, n_v, F, flags, 'cpu', torch.float32)
  assert list(gp.state_dict().keys()) == ['z', 'u_mean', 'u_tril_vec', 'kernel.log_mean', 'kernel.log_logvar',
                                          'kernel.prior_log_mean', 'kernel.prior_log_logvar']
