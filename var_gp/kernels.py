"""Drop-in module path of the reference (`var_gp/kernels.py`); implementation: `vargp_b200/kernels.py`."""
from vargp_b200.kernels import *          # noqa: F401,F403
from vargp_b200 import kernels as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
