"""Drop-in module path of the reference (`var_gp/vargp_retrain.py`); implementation: `vargp_b200/vargp_retrain.py`."""
from vargp_b200.vargp_retrain import *          # noqa: F401,F403
from vargp_b200 import vargp_retrain as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
