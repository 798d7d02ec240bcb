"""Loader / driver of the UNMODIFIED reference hot path under ``oracle/_ref/refvargp`` (see oracle/make_ref.py).

TEST / BENCH INFRASTRUCTURE ONLY (same rule as oracle/vargp_oracle.py): used by ``bench.py --impl reference`` and
its ``cpu_baseline`` leg, and by tests that cross-check the oracle port against the live reference.

One value-preserving shim is applied, outside the reference's files: torch 2.11's CPU ``F.nll_loss`` backward
rejects the permuted (non-contiguous) input of var_gp/likelihoods.py:43-46, so ``likelihoods.F`` is replaced by a
proxy whose ``nll_loss`` calls the real one on ``input.contiguous()`` (bit-identical forward).
"""
import importlib
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
_MODS = None


def available():
  return os.path.exists(os.path.join(HERE, '_ref', 'refvargp', 'vargp.py'))


def load():
  """-> (VARGP, RBFKernel, MulticlassSoftmax) of the reference, stock code."""
  global _MODS
  if _MODS is None:
    if not available():
      raise RuntimeError('oracle/_ref/refvargp is missing: run `python oracle/make_ref.py` in the build container')
    # torch.cholesky / torch.triangular_solve deprecation notices of the reference's stock calls
    warnings.filterwarnings('ignore', category=UserWarning, module=r'refvargp\..*')
    p = os.path.join(HERE, '_ref')
    if p not in sys.path:
      sys.path.insert(0, p)
    with warnings.catch_warnings():
      warnings.simplefilter('ignore')          # SyntaxWarnings for `\s`, `\i` in the reference's docstrings
      lk = importlib.import_module('refvargp.likelihoods')
      import torch.nn.functional as F

      class _F:
        def __getattr__(self, k):
          return getattr(F, k)

        @staticmethod
        def nll_loss(inp, tgt, **kw):
          return F.nll_loss(inp.contiguous(), tgt, **kw)

      lk.F = _F()
      vg = importlib.import_module('refvargp.vargp')
      kn = importlib.import_module('refvargp.kernels')
    _MODS = (vg.VARGP, kn.RBFKernel, lk.MulticlassSoftmax)
  return _MODS


def build_model(params, prev, n_v, F):
  """The reference VARGP carrying a `make_case` problem (same construction as tests/golden/make_golden.py)."""
  VARGP, RBFKernel, MulticlassSoftmax = load()
  D = params['z'].size(-1)
  with warnings.catch_warnings():
    warnings.simplefilter('ignore')
    kern = RBFKernel(D, prior_log_mean=params['prior_log_mean'].clone(), prior_log_logvar=params['prior_log_logvar'].clone())
    gp = VARGP(params['z'].clone(), kern, MulticlassSoftmax(n_f=F), n_var_samples=n_v,
               prev_params=[{k: v.clone() for k, v in p.items()} for p in prev])
  with torch.no_grad():
    gp.u_mean.copy_(params['u_mean'])
    gp.u_tril_vec.copy_(params['u_tril_vec'])
    gp.kernel.log_mean.copy_(params['log_mean'])
    gp.kernel.log_logvar.copy_(params['log_logvar'])
  return gp
