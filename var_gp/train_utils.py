"""Drop-in module path of the reference (`var_gp/train_utils.py`): `set_seeds`, `compute_accuracy`, `compute_acc_ent`,
`compute_bwt`, `EarlyStopper` with the reference's signatures; implementation: `vargp_b200/train.py`."""
from vargp_b200.train import set_seeds, compute_accuracy, compute_acc_ent, compute_bwt, EarlyStopper  # noqa: F401
from .vargp import VARGP  # noqa: F401
