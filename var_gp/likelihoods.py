"""Drop-in module path of the reference (`var_gp/likelihoods.py`); implementation: `vargp_b200/likelihoods.py`."""
from vargp_b200.likelihoods import *          # noqa: F401,F403
from vargp_b200 import likelihoods as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
