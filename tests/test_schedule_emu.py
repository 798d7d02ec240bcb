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


def test_dkl_fused_schedule_agrees_with_composed_fp64(emu_ops):
  """DeepRBFKernel (var_gp/kernels.py:80-96): through the `features` hook the fused schedule (with the
  x-side RBF adjoint that carries gradients back into phi) and the reference-order composed path are the same function
  of every parameter, incl. the MLP weights: fp64 agreement to 1e-10.  The bias of the last MLP layer has an exactly
  zero gradient (the RBF kernel is shift invariant)."""
  from vargp_b200.vargp import VARGP
  from vargp_b200.kernels import DeepRBFKernel
  from vargp_b200.likelihoods import MulticlassSoftmax
  from vargp_b200.composed import loss_composed
  torch.manual_seed(0)
  C, Din, M, B, Fd = 4, 20, 9, 48, 16
  g = torch.Generator().manual_seed(5)
  prev = [dict(z=torch.rand(C, M, Din, generator=g), u_mean=0.5 * torch.randn(C, M, 1, generator=g),
               u_tril_vec=0.1 * torch.randn(C, M * (M + 1) // 2, generator=g))]
  gp = VARGP(torch.rand(C, M, Din, generator=g), DeepRBFKernel(Din, feature_size=Fd), MulticlassSoftmax(n_f=5),
             n_var_samples=2, prev_params=prev)
  with torch.no_grad():
    gp.kernel.log_mean[:Fd] = 0.3
  gp = gp.double()
  x, y = torch.rand(B, Din, generator=g).double(), torch.randint(0, C, (B,), generator=g)
  nz = dict(eps_theta=torch.randn(2, Fd + 1, generator=g).double(), eps_f=torch.randn(2, 5, C, B, generator=g).double(),
            eps_u=torch.randn(2, 2, C, M, generator=g).double())
  grads = lambda: {n: p.grad.detach().clone() for n, p in gp.named_parameters()}
  kl_h, kl_u, nll = loss_composed(gp, x, y, nz)
  gp.zero_grad(); (kl_h + kl_u + 7. * nll).backward()
  g_c = grads()
  kl_h2, kl_u2, nll2 = gp.loss(x, y, noise=nz)
  gp.zero_grad(); (kl_h2 + kl_u2 + 7. * nll2).backward()
  g_f = grads()
  assert util.relerr(kl_u2, kl_u) < 1e-10 and util.relerr(nll2, nll) < 1e-10
  for k in g_c:
    if k == 'kernel.phi.4.bias':
      assert g_f[k].norm() < 1e-10 * g_f['kernel.phi.4.weight'].norm()
    else:
      assert util.relerr(g_f[k], g_c[k]) < 1e-10, k


def test_compute_q_and_pf_diag_methods_match_oracle(emu_ops):
  """The reference's two intermediate methods (var_gp/vargp.py:35-113) keep their signatures and values: moments of
  q(u_<t), q(u_<=t), the stacked inducing inputs and the predictive marginal, against the oracle in fp64."""
  from oracle import vargp_oracle as orc
  rec = util.load_golden('odd_t2')
  params, prev, x, y, noise, n_v, F, flags = util.case_tensors(rec['case'], torch.float64)
  gp = util.build_model(params, prev, n_v, F, flags, 'cpu', torch.float64)
  theta = orc.sample_hypers(params['log_mean'], params['log_logvar'], noise['eps_theta'])
  cache = dict()
  with torch.no_grad():
    out = gp.compute_q(theta, cache=cache)
    f_mean, f_var = gp.compute_pf_diag(theta, x, *out[2:])
  prevp = [dict(z=p['z'], u_mean=p['u_mean'], u_tril=orc.vec2tril(p['u_tril_vec'])) for p in prev]
  ocache = dict()
  ref = orc.compute_q(theta, prevp, params['z'], params['u_mean'], params['u_tril_vec'], cache=ocache)
  for a, b in zip(out, ref):
    assert a.shape == b.shape and (a - b).abs().max() < 1e-10
  for k in ('Lz_lt', 'Lz_lt_Kz_lt_z_t'):
    assert (cache[k] - ocache[k]).abs().max() < 1e-10
  assert (f_mean - rec['f64']['f_mean']).abs().max() < 1e-9 and (f_var - rec['f64']['f_var']).abs().max() < 1e-9


@pytest.mark.parametrize('flag', ['STACK_CLASSES', 'V_SIDE', 'both'])
@pytest.mark.parametrize('name', ['mnist_t3', 'odd_t2', 'toy_t0'])
def test_optional_schedules_are_the_same_function(name, flag, emu_ops, monkeypatch):
  """The opt-in schedule variants of elbo.py (class-stacked Kzx / Gz1 products, V on the side branch) compute the same
  ELBO terms and gradients as the reference fixtures."""
  from vargp_b200 import elbo
  for f in (('STACK_CLASSES', 'V_SIDE') if flag == 'both' else (flag,)):
    monkeypatch.setattr(elbo, f, True)
  rec = util.load_golden(name)
  params, prev, x, y, noise, n_v, F, flags = util.case_tensors(rec['case'], torch.float64)
  ref = rec['f64']
  gp = util.build_model(params, prev, n_v, F, flags, 'cpu', torch.float64)
  terms, grads = util.run_model(gp, x, y, noise, ref['beta'], ref['Ntot'])
  for k in ('kl_u', 'nll', 'total'):
    assert util.relerr(terms[k], ref[k]) < 1e-8, k
  for k in util.GRAD_KEYS:
    assert util.relerr(grads[k], ref['grads'][k]) < 1e-7, k


def test_fused_schedule_at_the_benched_size_fp64(emu_ops):
  """The benched Split-MNIST t=4 step at FULL size (P=300, B=512) through the host schedule in fp64 vs the live
  reference's stored fp64 outputs (tests/golden/make_large.py; big gradients compared through subsample + projections)."""
  rec = util.load_golden('large_split_t4')
  params, prev, x, y, noise, n_v, F, flags = util.case_tensors(rec['case'], torch.float64)
  ref = rec['f64']
  gp = util.build_model(params, prev, n_v, F, flags, 'cpu', torch.float64)
  terms, grads = util.run_model(gp, x, y, noise, rec['beta'], rec['Ntot'])
  for k in ('kl_hypers', 'kl_u', 'nll', 'total'):
    assert util.relerr(terms[k], ref[k]) < 1e-8, k
  for k in util.GRAD_KEYS:
    assert util.compressed_err(grads[k], ref['grads'][k]) < 1e-7, k
