"""Drop-in module path of the reference (`var_gp/gp_utils.py`); implementation: `vargp_b200/gp_utils.py`."""
from vargp_b200.gp_utils import *          # noqa: F401,F403
from vargp_b200 import gp_utils as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
