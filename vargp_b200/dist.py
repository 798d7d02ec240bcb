"""Data-parallel ELBO over the GPUs of one box.

Only the minibatch data term shards: rank r gets its own slice of the minibatch; Kzz / Cholesky / KL are replicated
(identical theta on every rank: same seed, same draw).  The drivers here (bench.py, tests) seed every rank identically,
so the likelihood noise eps_f / eps_u is the SAME tensor on every rank, applied to different data slices: the estimator
stays unbiased, its noise is correlated across ranks (a per-rank generator for eps_f would decorrelate it; theta must
stay shared).  Each rank forms
    loss_r = (beta * kl_hypers + kl_u) / R + (N / B_global) * nll_r
so that the SUM over ranks of the gradients is the full-batch gradient, and one all-reduce(SUM) of a flat
gradient bucket per step is the only collective (NCCL over NVLink; gloo in the CPU tests)."""
import torch
import torch.distributed as dist


class GradBucket:
  """Flat fp32 bucket holding the gradients of `params`; `allreduce()` sums it across ranks in one call."""

  def __init__(self, params):
    self.params = [p for p in params if p.requires_grad]
    n = sum(p.numel() for p in self.params)
    p0 = self.params[0]
    self.flat = torch.zeros(n, device=p0.device, dtype=p0.dtype)
    self.views, o = [], 0
    for p in self.params:
      self.views.append(self.flat[o:o + p.numel()].view_as(p))
      o += p.numel()

  def allreduce(self):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
      return
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
    torch._foreach_copy_(self.views, grads)
    dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
    for p, v in zip(self.params, self.views):
      if p.grad is None:
        p.grad = v.clone()
      else:
        p.grad.copy_(v)


def shard_coef(beta, n_data, global_batch, world_size, factor_sharded=False):
  """Coefficients of (kl_hypers, kl_u, nll_local) in the per-rank loss.  With the factor stage sharded
  (elbo.FactorShard) every rank holds only its share of kl_u, so that term enters with weight 1."""
  return (beta / world_size, 1.0 if factor_sharded else 1.0 / world_size, n_data / global_batch)


def shard_loss(kl_hypers, kl_u, nll_local, beta, n_data, global_batch, world_size, coef=None):
  """Per-rank loss whose gradients SUM (over ranks) to the gradient of the full-batch ELBO.
  With `coef` (a device tensor holding `shard_coef(...)`) the combination is one fused autograd node."""
  if coef is not None:
    from .functional import Combine3Fn
    return Combine3Fn.apply(kl_hypers, kl_u, nll_local, coef)
  a, b, c = shard_coef(beta, n_data, global_batch, world_size)
  return a * kl_hypers + b * kl_u + c * nll_local
