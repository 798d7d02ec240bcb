"""Drop-in for ``var_gp/likelihoods.py`` (MulticlassSoftmax): Monte-Carlo softmax likelihood on
libvargp_sm100.so.  ``GaussianLikelihood`` of the reference is never instantiated by any experiment and is
out of scope."""
import torch
import torch.nn as nn

from .functional import SoftmaxNllFn, softmax_predict


class MulticlassSoftmax(nn.Module):
  """var_gp/likelihoods.py:7-63."""

  def __init__(self, n_f=1):
    super().__init__()
    self.n_f = n_f

  def _noise(self, mu, eps):
    n_hypers, out_size, B = mu.shape
    if eps is None:
      # same call (shape, device) as var_gp/likelihoods.py:26 -> same stream of the global generator
      eps = torch.randn(n_hypers, self.n_f, out_size, B, device=mu.device)
    return eps

  def forward(self, mu, var, eps=None):
    """mu, var (n_hypers, out_size, B) -> log-softmax samples (n_hypers, n_f, out_size, B).
    Kept for API parity (plain torch: the fused kernels never materialise this tensor)."""
    eps = self._noise(mu, eps)
    f_samples = mu.unsqueeze(1) + var.sqrt().unsqueeze(1) * eps
    return torch.log_softmax(f_samples, dim=-2)

  def loss(self, pred_mu, pred_var, y, eps=None):
    """sum_b mean_h mean_f -log p(y_b | f)  -> scalar."""
    eps = self._noise(pred_mu, eps)
    return SoftmaxNllFn.apply(pred_mu, pred_var, y, eps)

  def predict(self, mu, var, eps=None):
    """-> class probabilities (B, out_size)."""
    eps = self._noise(mu, eps)
    return softmax_predict(mu, var, eps)
