"""Fused ELBO value-and-gradient for the training step (experiments/vargp.py:30-36 of the reference):

    kl_h, kl_u, lik = gp.loss(x, y);  loss = beta kl_h + kl_u + (N / B) lik;  loss.backward()

as ONE straight-line launch sequence on libvargp_sm100.so, without the autograd tape.  The numerics are the ones of
`VARGP.loss` (same kernels, same schedule: `elbo.marginal_forward` / `elbo.marginal_backward`); what disappears is the
plumbing around them, ~25 small torch launches per step at the Split-MNIST shape:

  * `torch.cat` of the previous and current inducing inputs / means / factors and the tril unpack  ->  the previous
    tasks' parts are written ONCE into static stacked buffers, `ops.step_assemble` drops the current task's in (1 launch)
  * zero-filled accumulators (Cholesky status, KL, NLL workspace, the three loss terms)  ->  one per-step arena, 1 fill
  * loss combination, the scaling of the likelihood adjoints by N / B, the two sums over hyper samples, the slices back
    to parameter shapes, the tril / hyper adjoints and the five AccumulateGrad adds  ->  the coefficients enter the
    kernels directly and `ops.step_grad_finish` writes every parameter gradient into its view of the flat Yogi gradient
    buffer (1 launch); gradients of the previous tasks' (constant) variational parameters are not formed at all

RNG parity with the reference is kept: the three draws (hypers -> u_<t -> likelihood noise) are issued with the same
calls, shapes and order as `VARGP.loss`.  Eligible models: plain `RBFKernel` (no DKL feature map), sampled hypers,
`ep_var_mean=True`; `ElboStepper` falls back to the autograd path otherwise.
"""
import torch

from . import elbo
from . import ops as _ops_mod


def eligible(gp):
  from .kernels import RBFKernel
  from .likelihoods import MulticlassSoftmax
  return (type(gp.kernel) is RBFKernel and not gp.kernel.map_est and gp.var_mean_mask == 1.0 and
          type(gp.likelihood) is MulticlassSoftmax and gp.z.dtype == torch.float32 and gp.z.is_cuda and
          type(gp).__name__ == 'VARGP')


class FusedElbo:
  def __init__(self, gp, coef, shard=None):
    """coef: (3,) device tensor, the weights of (kl_hypers, kl_u, nll) in the loss (`dist.shard_coef`)."""
    self.gp, self.coef, self.shard = gp, coef, shard
    self.nll_scale = float(coef[2].item())            # enters the likelihood kernel as a launch constant
    z = gp.z
    C, M, D = z.shape
    S = gp.n_prev + 1
    P = S * M
    dev, dt = z.device, z.dtype
    self.dims = (C, M, D, S, P)
    # stacked operands of the step; the previous tasks' parts never change
    self.Zcat = torch.empty(C, P, D, device=dev, dtype=dt)
    self.m_all = torch.empty(S, C, M, device=dev, dtype=dt)
    self.Lu_all = torch.empty(S, C, M, M, device=dev, dtype=dt)
    if gp.n_prev:
      with torch.no_grad():
        self.Zcat[:, :P - M].copy_(gp.prev_z)
        self.m_all[:S - 1].copy_(gp.prev_u_mean)
        self.Lu_all[:S - 1].copy_(gp._prev_factors())
    self.terms = None

  def value_and_grad(self, x, y):
    """Runs forward + backward on the minibatch (x, y) and leaves d loss / d parameter in every parameter's `.grad`
    (overwritten, not accumulated).  Returns the (3,) device tensor (kl_hypers, kl_u, nll)."""
    ops = _ops_mod.get_ops()
    gp, kern = self.gp, self.gp.kernel
    C, M, D, S, P = self.dims
    H, F = gp.n_v, gp.likelihood.n_f
    B = x.shape[0]
    G = H * C
    dev, dt = x.device, x.dtype
    for p in (gp.z, gp.u_mean, gp.u_tril_vec, kern.log_mean, kern.log_logvar):
      if p.grad is None or not p.grad.is_contiguous():
        p.grad = torch.empty_like(p)
    # ---- draws, in the reference's order ----
    eps_theta = torch.empty(H, D + 1, device=dev, dtype=dt).normal_()
    # the other two draws depend on nothing: issued here, in the reference's order, on the side stream, so that they run
    # beside the prologue instead of between the marginal reduction and the likelihood kernel (marginal_forward joins the
    # side stream before it returns)
    eps_u = torch.empty((gp.n_v, H, C, gp.n_prev * M), dtype=dt, device=dev) if gp.n_prev else None
    eps_f = torch.empty(H, F, C, B, device=dev, dtype=dt)
    with elbo._Fork(dev):
      if eps_u is not None:
        # the reference draws u_<t here (vargp.py:138); with ep_var_mean=True nothing depends on it -- the draw is issued
        # so that the generator stays in step with the reference
        eps_u.normal_()
      eps_f.normal_()
    # ---- per-step arena: [terms(3) | pad | info(G) | whiten workspace(1 + G) | nll workspace], one fill ----
    nw, ww = ops.nll_work(H, B), ops.whiten_work(H, C)
    arena = torch.zeros(4 + G + ww + nw, device=dev, dtype=dt)
    terms, info = arena[:3], arena[4:4 + G].view(torch.int32)
    wwork, work = arena[4 + G:4 + G + ww], arena[4 + G + ww:]
    theta = torch.empty(H, D + 1, device=dev, dtype=dt)
    ops.hyper_fwd(kern.log_mean.detach(), kern.log_logvar.detach(), kern.prior_log_mean, kern.prior_log_logvar, eps_theta,
                  theta, terms[0])
    ops.step_assemble(gp.z.detach(), gp.u_mean.detach(), gp.u_tril_vec.detach(), self.Zcat, self.m_all[S - 1],
                      self.Lu_all[S - 1])
    ctx = elbo._Ctx()
    f_mean, f_var, _, _, _ = elbo.marginal_forward(theta, self.Zcat, x, self.m_all, self.Lu_all, M, True, ctx,
                                                   shard=self.shard, zeroed=(info, terms[1], wwork))
    gp._last_info = info
    gmv = torch.empty(2, H, C, B, device=dev, dtype=dt)
    ops.nll_fwd_bwd(f_mean, f_var, eps_f, y, terms[2], gmv[0], gmv[1], work=work, gscale=self.nll_scale)
    # ---- backward ----
    theta_bar, Z_bar, _, mbar, Lubar = elbo.marginal_backward(ctx, gmv[0], gmv[1], self.coef[1:2], last_raw=True)
    kl_lu = self.coef[1:2] if (self.shard is None or self.shard.rank == 0) else None
    ops.step_grad_finish(Z_bar, mbar, Lubar, self.Lu_all[S - 1], gp.u_tril_vec.detach(), kl_lu,
                         kern.log_mean.detach(), kern.log_logvar.detach(), kern.prior_log_mean, kern.prior_log_logvar,
                         eps_theta, theta_bar, self.coef[0:1], gp.z.grad, gp.u_mean.grad, gp.u_tril_vec.grad,
                         kern.log_mean.grad, kern.log_logvar.grad)
    self.terms = terms
    return terms
