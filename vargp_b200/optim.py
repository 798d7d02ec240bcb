"""Yogi optimizer (Zaheer et al. 2018), the optimizer the reference trains with
(``torch_optimizer.Yogi`` at experiments/vargp.py:23; that package is not vendored, so the update rule is
restated here with its defaults: betas (0.9, 0.999), eps 1e-3, initial accumulator 1e-6).
Implemented with torch._foreach ops (a handful of launches per step, CUDA-graph capturable); the same class
drives both the B200 path and the CPU reference arm in bench.py so that step counts compare like for like."""
import math

import torch


class Yogi(torch.optim.Optimizer):
  def __init__(self, params, lr=1e-2, betas=(0.9, 0.999), eps=1e-3, initial_accumulator=1e-6):
    super().__init__(params, dict(lr=lr, betas=betas, eps=eps, initial_accumulator=initial_accumulator))
    self._t = 0

  @torch.no_grad()
  def step(self):
    self._t += 1
    for group in self.param_groups:
      ps = [p for p in group['params'] if p.grad is not None]
      if not ps:
        continue
      gs = [p.grad for p in ps]
      b1, b2 = group['betas']
      for p in ps:
        st = self.state[p]
        if not st:
          st['m'] = torch.full_like(p, group['initial_accumulator'])
          st['v'] = torch.full_like(p, group['initial_accumulator'])
      ms = [self.state[p]['m'] for p in ps]
      vs = [self.state[p]['v'] for p in ps]
      torch._foreach_mul_(ms, b1)
      torch._foreach_add_(ms, gs, alpha=1 - b1)
      g2 = torch._foreach_mul(gs, gs)
      # v <- v - (1 - b2) * sign(v - g^2) * g^2
      diff = torch._foreach_sub(vs, g2)
      sg = [d.sign_() for d in diff]
      torch._foreach_mul_(sg, g2)
      torch._foreach_add_(vs, sg, alpha=-(1 - b2))
      bc1 = 1 - b1 ** self._t
      bc2 = 1 - b2 ** self._t
      den = torch._foreach_sqrt(vs)
      torch._foreach_div_(den, math.sqrt(bc2))
      torch._foreach_add_(den, group['eps'])
      torch._foreach_addcdiv_(ps, ms, den, value=-group['lr'] / bc1)


class FlatYogi:
  """Yogi over ONE flat fp32 buffer, updated by a single fused kernel of libvargp_sm100.so.

  The parameters are re-pointed to views of `flat_p` and their `.grad` to views of `flat_g`, so
    * `zero_grad()` is one memset and `step()` one launch (CUDA-graph replayable: the bias-correction
      powers live on the device),
    * `flat_g` IS the data-parallel gradient bucket: `all_reduce(flat_g)` needs no pack/unpack copies
      ("backward kernels write straight into the flat bucket").
  Same update rule and defaults as `Yogi` above (tested against it)."""

  def __init__(self, params, lr=1e-2, betas=(0.9, 0.999), eps=1e-3, initial_accumulator=1e-6):
    from . import ops as _ops_mod
    self._ops = _ops_mod.get_ops
    self.params = [p for p in params if p.requires_grad]
    self.lr, self.betas, self.eps = lr, betas, eps
    self.n_params = sum(p.numel() for p in self.params)
    n = (self.n_params + 3) // 4 * 4             # padded to whole float4s (the pad holds zeros and stays zero)
    p0 = self.params[0]
    self.flat_p = torch.zeros(n, device=p0.device, dtype=p0.dtype)
    self.flat_g = torch.zeros(n, device=p0.device, dtype=p0.dtype)
    self.m = torch.full((n,), initial_accumulator, device=p0.device, dtype=p0.dtype)
    self.v = torch.full((n,), initial_accumulator, device=p0.device, dtype=p0.dtype)
    self.pows = torch.ones(2, device=p0.device, dtype=p0.dtype)
    self.peer = None
    o = 0
    with torch.no_grad():
      for p in self.params:
        k = p.numel()
        self.flat_p[o:o + k].copy_(p.reshape(-1))
        p.data = self.flat_p[o:o + k].view_as(p)
        p.grad = self.flat_g[o:o + k].view_as(p)
        o += k

  def zero_grad(self, set_to_none=False):
    self.flat_g.zero_()

  @torch.no_grad()
  def step(self):
    self._ops().yogi_step(self.flat_p, self.flat_g, self.m, self.v, self.lr, self.betas[0], self.betas[1],
                          self.eps, self.pows)

  def enable_peer_allreduce(self, group=None):
    """Data parallel: allocate this rank's staging buffer in symmetric memory (torch.distributed._symmetric_memory) and
    exchange the peers' pointers, so that `step_allreduce` can run the gradient all-reduce and the Yogi update as one
    kernel pair over NVLink peer memory (csrc/peer.cu) instead of an NCCL all-reduce followed by `step`.  Collective:
    every rank of `group` must call it.  Returns False (and leaves the NCCL route in place) if symmetric memory is not
    available."""
    import torch.distributed as dist
    try:
      import torch.distributed._symmetric_memory as symm
      grp = group or dist.group.WORLD
      nf = self._ops().peer_buffer_floats(self.flat_g.numel())
      buf = symm.empty(nf, dtype=torch.float32, device=self.flat_g.device)
      hdl = symm.rendezvous(buf, grp)
      buf.zero_()
      torch.cuda.synchronize(self.flat_g.device)
      dist.barrier(group=group)                     # nobody signals before every rank has cleared its flags
      ptrs = [int(q) for q in hdl.buffer_ptrs]
      ok = len(ptrs) == dist.get_world_size(group) and all(ptrs)
    except Exception as e:                          # no P2P / no symmetric-memory backend on this box
      import warnings
      warnings.warn(f'peer all-reduce unavailable ({type(e).__name__}: {e}); using NCCL all_reduce + yogi_step')
      ok = False
    # the decision must be the same on every rank
    flag = torch.tensor([1 if ok else 0], device=self.flat_g.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if not int(flag.item()):
      return False
    # NVLS: with a multicast mapping of the symmetric buffers the reduce kernel lets the NVSwitch add the ranks' gradients
    # (one multimem.ld_reduce per 16 bytes instead of a load from every peer); VARGP_PEER_NVLS=0 keeps the peer loads
    import os
    mc = 0
    if os.environ.get('VARGP_PEER_NVLS', '1') != '0':
      try:
        mc = int(hdl.multicast_ptr or 0)
      except Exception:
        mc = 0
    mflag = torch.tensor([1 if mc else 0], device=self.flat_g.device)
    dist.all_reduce(mflag, op=dist.ReduceOp.MIN, group=group)
    if not int(mflag.item()):
      mc = 0
    self.peer = dict(buf=buf, hdl=hdl, ptrs=ptrs, rank=dist.get_rank(group), multicast=mc,
                     ctr=torch.zeros(4, dtype=torch.int32, device=self.flat_g.device))
    return True

  @torch.no_grad()
  def step_allreduce(self):
    """flat_g <- sum over ranks; Yogi update with the summed gradient (fused, peer memory)."""
    pr = self.peer
    self._ops().peer_allreduce_yogi(pr['ptrs'], pr['rank'], self.flat_g, self.flat_p, self.m, self.v, self.lr,
                                    self.betas[0], self.betas[1], self.eps, self.pows, pr['ctr'],
                                    multicast=pr.get('multicast', 0))
