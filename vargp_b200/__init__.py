"""vargp_b200: B200-native (sm_100a) implementation of VAR-GP's sparse-variational ELBO hot path.

Public API mirrors the reference package ``var_gp``:
    vargp_b200.vargp.VARGP, vargp_b200.kernels.{RBFKernel, DeepRBFKernel},
    vargp_b200.gp_utils.{cholesky, rev_cholesky, vec2tril, mat2trilvec, gp_cond, linear_joint, linear_marginal_diag},
    vargp_b200.likelihoods.MulticlassSoftmax
(the repo-root ``var_gp`` package re-exports them under the reference's module paths).
All arithmetic runs in ``libvargp_sm100.so`` (hand-written CUDA, C ABI in include/vargp_sm100.h).
"""
from .vargp import VARGP                       # noqa: F401
from .kernels import RBFKernel, DeepRBFKernel  # noqa: F401
from .likelihoods import MulticlassSoftmax     # noqa: F401

__version__ = '0.1.0'
