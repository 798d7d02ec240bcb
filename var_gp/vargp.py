"""Drop-in module path of the reference (`var_gp/vargp.py`); implementation: `vargp_b200/vargp.py`."""
from vargp_b200.vargp import *          # noqa: F401,F403
from vargp_b200 import vargp as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
