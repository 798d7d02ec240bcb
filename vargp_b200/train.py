"""Training-step driver for the ELBO hot path (first slice of SURVEY.md section 8f, N1).

`ElboStepper` owns one optimisation step of experiments/vargp.py:30-37,

    zero_grad -> kl_h, kl_u, lik = gp.loss(x, y) -> loss = beta kl_h + kl_u + (N/B) lik -> backward -> Yogi step,

with the data-parallel reduction of `vargp_b200.dist` folded in, and can capture the whole step (about 45
library launches plus torch's autograd glue) into ONE CUDA graph: the Split-MNIST-shape step is a few hundred
microseconds of kernels, so per-launch host overhead would otherwise dominate.  Inputs are copied into static
device buffers before each replay; the three loss terms come back in static 0-d tensors.
"""
import torch
import torch.distributed as dist

from .dist import shard_coef, shard_loss
from .optim import FlatYogi


class ElboStepper:
  def __init__(self, gp, n_data, batch_size, beta=1.0, lr=1e-2, world_size=1, use_graph=True, optimizer=None):
    self.gp, self.n_data, self.beta, self.world = gp, n_data, beta, world_size
    self.global_batch = batch_size * world_size
    self.opt = optimizer or FlatYogi(gp.parameters(), lr=lr)
    self.use_graph = use_graph
    self.graph = None
    dev = next(gp.parameters()).device
    D = gp.z.size(-1) if not hasattr(gp.kernel, 'phi') else gp.kernel.phi[0].in_features
    self.x = torch.empty(batch_size, D, device=dev)
    self.y = torch.empty(batch_size, dtype=torch.int64, device=dev)
    self.coef = torch.tensor(shard_coef(beta, n_data, self.global_batch, world_size), device=dev)
    self.terms = None
    self.launches_per_step = None
    gp.sync_errors = False

  def _body(self):
    self.opt.zero_grad()
    kl_h, kl_u, nll = self.gp.loss(self.x, self.y)
    loss = shard_loss(kl_h, kl_u, nll, self.beta, self.n_data, self.global_batch, self.world, coef=self.coef)
    loss.backward()
    if self.world > 1:
      dist.all_reduce(self.opt.flat_g, op=dist.ReduceOp.SUM)
    self.opt.step()
    return kl_h.detach(), kl_u.detach(), nll.detach()

  def _capture(self):
    from . import ops as _ops_mod
    ops = _ops_mod.get_ops()
    # warm up on a side stream (allocator pools, lazy inits, cuBLAS-free so nothing else to prime)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
      for _ in range(3):
        self._body()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    self.graph = torch.cuda.CUDAGraph()
    n0 = ops.launch_count()
    with torch.cuda.graph(self.graph):
      self.terms = self._body()
    self.launches_per_step = ops.launch_count() - n0

  def step(self, x, y):
    """One optimisation step on the minibatch (x, y) (device or pinned-host tensors).  Returns the three
    loss terms as 0-d device tensors (valid until the next call)."""
    self.x.copy_(x, non_blocking=True)
    self.y.copy_(y, non_blocking=True)
    if not self.use_graph:
      return self._body()
    if self.graph is None:
      self._capture()      # note: the capture itself does not advance the parameters
    self.graph.replay()
    return self.terms
