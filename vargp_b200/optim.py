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
