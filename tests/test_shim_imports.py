"""The drop-in boundary: every public name of the reference's hot-path modules resolves through
the `var_gp.*` module paths, with the reference's positional parameters in the reference's order (extra trailing
keyword arguments such as `noise=` / `eps=` are additive).  CPU only: nothing is computed."""
import importlib
import inspect

import pytest

# module -> {dotted name: parameter names of the reference (file:line of its definition under /root/reference/var_gp)}
REFERENCE_API = {
  'var_gp.vargp': {                                                                        # vargp.py
    'VARGP.__init__': ['self', 'z_init', 'kernel', 'likelihood', 'n_var_samples', 'ep_var_mean', 'prev_params'],  # :12
    'VARGP.compute_q': ['self', 'theta', 'cache'],                                         # :35
    'VARGP.compute_pf_diag': ['self', 'theta', 'x', 'mu_leq_t', 'S_leq_t', 'z_leq_t', 'cache'],   # :90
    'VARGP.forward': ['self', 'x', 'loss_cache'],                                          # :115
    'VARGP.loss': ['self', 'x', 'y'],                                                      # :177
    'VARGP.predict': ['self', 'x'],                                                        # :196
    'VARGP.create_clf': ['dataset', 'M', 'n_f', 'n_var_samples', 'prev_params', 'ep_var_mean', 'map_est_hypers', 'dkl'],  # :201
  },
  'var_gp.kernels': {                                                                      # kernels.py
    'RBFKernel.__init__': ['self', 'in_size', 'prior_log_mean', 'prior_log_logvar', 'map_est'],   # :8
    'RBFKernel.compute': ['self', 'kern_samples', 'x', 'y'],                               # :24
    'RBFKernel.compute_diag': ['self', 'kern_samples'],                                    # :58
    'RBFKernel.sample_hypers': ['self', 'n_hypers'],                                       # :62
    'RBFKernel.kl_hypers': ['self'],                                                       # :70
    'DeepRBFKernel.__init__': ['self', 'in_size', 'feature_size'],                         # :81
    'DeepRBFKernel.compute': ['self', 'kern_samples', 'x', 'y'],                           # :92
  },
  'var_gp.gp_utils': {                                                                     # gp_utils.py
    'cholesky': ['M', 'eps'],                                                              # :5
    'rev_cholesky': ['L'],                                                                 # :14
    'vec2tril': ['vec', 'm'],                                                              # :22
    'mat2trilvec': ['mat'],                                                                # :52
    'gp_cond': ['u', 'Kzz', 'Kzx', 'Kxx', 'Lz', 'Lz_Kzx'],                                 # :68
    'linear_joint': ['m', 'S', 'Kzx', 'Kzz', 'V', 'b', 'cache'],                           # :101
    'linear_marginal_diag': ['m', 'S', 'Kzz', 'Kzx', 'Kxx_diag', 'cache'],                 # :150
  },
  'var_gp.likelihoods': {                                                                  # likelihoods.py
    'MulticlassSoftmax.__init__': ['self', 'n_f'],                                         # :8
    'MulticlassSoftmax.forward': ['self', 'mu', 'var'],                                    # :13
    'MulticlassSoftmax.loss': ['self', 'pred_mu', 'pred_var', 'y'],                        # :33
    'MulticlassSoftmax.predict': ['self', 'mu', 'var'],                                    # :49
  },
  'var_gp.train_utils': {                                                                  # train_utils.py
    'set_seeds': ['seed'],                                                                 # :13
    'compute_accuracy': ['dataset', 'gp', 'batch_size', 'device'],                         # :21
    'compute_acc_ent': ['dataset', 'gp', 'batch_size', 'device'],                          # :38
    'compute_bwt': ['acc_mat'],                                                            # :59
    'EarlyStopper.__init__': ['self', 'patience', 'delta'],                                # :70
    'EarlyStopper.is_done': ['self'],                                                      # :79
    'EarlyStopper.info': ['self'],                                                         # :84
    'VARGP.loss': ['self', 'x', 'y'],                      # `from .vargp import VARGP` (:10) is part of the module's namespace
  },
  'var_gp.vargp_retrain': {                                                                # vargp_retrain.py
    'VARGPRetrain.__init__': ['self', 'z_init', 'kernel', 'likelihood', 'n_var_samples', 'prev_params'],   # :12
    'VARGPRetrain.forward': ['self', 'x', 'loss_cache'],                                   # :119
    'VARGPRetrain.loss': ['self', 'x', 'y'],                                               # :191
    'VARGPRetrain.predict': ['self', 'x'],                                                 # :235
    'VARGPRetrain.create_clf': ['dataset', 'M', 'n_f', 'n_var_samples', 'prev_params'],    # :240
  },
}


@pytest.mark.parametrize('module', sorted(REFERENCE_API))
def test_reference_names_and_signatures_resolve_through_var_gp(module):
  mod = importlib.import_module(module)
  for dotted, ref_params in REFERENCE_API[module].items():
    obj = mod
    for part in dotted.split('.'):
      assert hasattr(obj, part), f'{module}.{dotted} is missing'
      obj = getattr(obj, part)
    got = list(inspect.signature(obj).parameters)
    if ref_params and ref_params[0] == 'self' and (not got or got[0] != 'self'):
      got = ['self'] + got                        # bound / static access drops `self`
    assert got[:len(ref_params)] == ref_params, f'{module}.{dotted}: {got} vs reference {ref_params}'
    for extra in got[len(ref_params):]:           # additions must be optional
      p = inspect.signature(obj).parameters[extra]
      assert p.default is not inspect.Parameter.empty or p.kind in (p.VAR_KEYWORD, p.VAR_POSITIONAL), (dotted, extra)


def test_shim_classes_are_the_product_classes():
  import var_gp.vargp, var_gp.kernels, var_gp.likelihoods, var_gp.train_utils
  import vargp_b200.vargp, vargp_b200.kernels, vargp_b200.likelihoods, vargp_b200.train
  assert var_gp.vargp.VARGP is vargp_b200.vargp.VARGP is var_gp.train_utils.VARGP
  assert var_gp.kernels.RBFKernel is vargp_b200.kernels.RBFKernel
  assert var_gp.likelihoods.MulticlassSoftmax is vargp_b200.likelihoods.MulticlassSoftmax
  assert var_gp.train_utils.EarlyStopper is vargp_b200.train.EarlyStopper


def test_reference_driver_import_lines_work():
  """experiments/vargp.py:9-11 of the reference, verbatim (datasets are out of scope: SURVEY.md section 2.1)."""
  from var_gp.train_utils import set_seeds, EarlyStopper, compute_accuracy  # noqa: F401
  from var_gp.vargp import VARGP  # noqa: F401
  set_seeds(1)
