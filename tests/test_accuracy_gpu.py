"""End-of-run accuracy parity (north star: per-task test accuracy of the toy and the synthetic Split-MNIST-shaped
continual runs within 0.5 pt of the reference path).

Both arms run the SAME continual protocol -- same data, same initial parameters, same minibatches, same pinned
noise, same Yogi rule, hyper-prior := previous posterior, previous (z, u_mean, u_tril_vec) frozen -- once through
the CPU oracle (reference op order, fp32) and once through the product model on the B200; the final models are
then evaluated on every task's held-out set with identical predictive noise."""
import math

import pytest
import torch

from oracle import vargp_oracle as orc
from tests import util

pytestmark = pytest.mark.gpu
LEAF = ('z', 'u_mean', 'u_tril_vec', 'log_mean', 'log_logvar')


def _toy_tasks(g):
  centers = torch.tensor([[-1., -1.], [1., 1.], [-1., 1.], [1., -1.]])
  def draw(n):
    y = torch.arange(4).repeat_interleave(n)
    x = centers[y] + 0.55 * torch.randn(4 * n, 2, generator=g)
    return x, y
  tr, te = draw(50), draw(200)
  pick = lambda d, cls: tuple(t[(d[1] == cls[0]) | (d[1] == cls[1])] for t in d)
  return [(pick(tr, c), pick(te, c)) for c in ((0, 1), (2, 3))], 4, 2


def _mnist_like_tasks(g, n_tasks=3, n_train=192, n_test=200, D=784, C=10):
  means = torch.rand(C, D, generator=g) * (torch.rand(C, D, generator=g) < 0.3)
  def draw(cls, n):
    y = torch.tensor(cls).repeat_interleave(n)
    x = (means[y] + 0.55 * torch.randn(len(y), D, generator=g)).clamp_(0., 1.)
    return x, y
  return [(draw((2 * t, 2 * t + 1), n_train), draw((2 * t, 2 * t + 1), n_test)) for t in range(n_tasks)], C, D


def _permuted_tasks(g, n_tasks=2, n_train=40, n_test=40, D=784, C=10):
  """Permuted-MNIST shape: every task holds all 10 classes, inputs pass through a task-specific pixel permutation."""
  means = torch.rand(C, D, generator=g) * (torch.rand(C, D, generator=g) < 0.3)
  def draw(n, perm):
    y = torch.arange(C).repeat_interleave(n)
    x = (means[y] + 0.45 * torch.randn(len(y), D, generator=g)).clamp_(0., 1.)
    return x[:, perm], y
  out = []
  for t in range(n_tasks):
    perm = torch.arange(D) if t == 0 else torch.randperm(D, generator=g)
    out.append((draw(n_train, perm), draw(n_test, perm)))
  return out, C, D


def _init_params(g, x_train, C, D, M, log_sigma, prior):
  idx = torch.stack([torch.randperm(x_train.size(0), generator=g)[:M] for _ in range(C)])
  T = M * (M + 1) // 2
  eye_vec = orc.mat2trilvec(torch.eye(M).expand(C, M, M))                       # var_gp/vargp.py:32-33
  if prior is None:
    log_mean = torch.full((D + 1,), log_sigma) + 0.05 * torch.randn(D + 1, generator=g)
    log_mean[D] = math.log(0.5)
    prior = (torch.zeros(D + 1), torch.zeros(D + 1))
    log_logvar = torch.full((D + 1,), -2.)
  else:
    log_mean, log_logvar = prior[0].clone(), prior[1].clone()
  assert eye_vec.shape == (C, T)
  return dict(z=x_train[idx].clone(), u_mean=0.5 * torch.randn(C, M, 1, generator=g), u_tril_vec=eye_vec.clone(),
              log_mean=log_mean, log_logvar=log_logvar, prior_log_mean=prior[0].clone(), prior_log_logvar=prior[1].clone())


def _step_noise(g, H, F, C, D, B, Q):
  nz = dict(eps_theta=torch.randn(H, D + 1, generator=g), eps_f=torch.randn(H, F, C, B, generator=g))
  if Q:
    nz['eps_u'] = torch.randn(H, H, C, Q, generator=g)
  return nz


def _run(arm, tasks, C, D, M, steps, B, lr, beta, log_sigma, H=3, F=10, F_eval=40):
  from vargp_b200.optim import Yogi, FlatYogi
  g = torch.Generator().manual_seed(11)                     # identical streams in both arms
  prev, prior = [], None
  final = None
  for t, ((xtr, ytr), _) in enumerate(tasks):
    p0 = _init_params(g, xtr, C, D, M, log_sigma, prior)
    N = xtr.size(0)
    if arm == 'oracle':
      p = {k: (v.clone().requires_grad_(True) if k in LEAF else v) for k, v in p0.items()}
      opt = Yogi([p[k] for k in LEAF], lr=lr)
    else:
      gp = util.build_model(p0, prev, H, F, {}, 'cuda', torch.float32)
      opt = FlatYogi(gp.parameters(), lr=lr)
    for s in range(steps):
      idx = torch.randperm(N, generator=g)[:B]
      x, y = xtr[idx], ytr[idx]
      nz = _step_noise(g, H, F, C, D, x.size(0), t * M)
      if arm == 'oracle':
        opt.zero_grad(set_to_none=True)
        kl_h, kl_u, nll = orc.elbo_terms(p, prev, x, y, nz, n_v=H)
      else:
        opt.zero_grad()
        kl_h, kl_u, nll = gp.loss(x.cuda(), y.cuda(), noise={k: v.cuda() for k, v in nz.items()})
      (beta * kl_h + kl_u + (N / x.size(0)) * nll).backward()
      opt.step()
    if arm == 'oracle':
      cur = {k: v.detach().clone() for k, v in p.items()}
    else:
      gp.check_errors()
      cur = dict(p0, z=gp.z.detach().cpu(), u_mean=gp.u_mean.detach().cpu(), u_tril_vec=gp.u_tril_vec.detach().cpu(),
                 log_mean=gp.kernel.log_mean.detach().cpu(), log_logvar=gp.kernel.log_logvar.detach().cpu())
    final = (cur, list(prev))
    prev = prev + [dict(z=cur['z'], u_mean=cur['u_mean'], u_tril_vec=cur['u_tril_vec'])]
    prior = (cur['log_mean'], cur['log_logvar'])            # hyper-prior := previous posterior (vargp.py:216-217)
  # evaluation of the final model on every task's test set
  cur, prv = final
  ge = torch.Generator().manual_seed(5)
  accs, probs = [], []
  for _, (xte, yte) in tasks:
    nz = dict(eps_theta=torch.randn(H, D + 1, generator=ge), eps_f=torch.randn(H, F_eval, C, xte.size(0), generator=ge))
    with torch.no_grad():
      if arm == 'oracle':
        pr = orc.predict(cur, prv, xte, nz, n_v=H)
      else:
        gp = util.build_model(cur, prv, H, F_eval, {}, 'cuda', torch.float32)
        pr = gp.predict(xte.cuda(), noise={k: v.cuda() for k, v in nz.items()}).cpu()
    accs.append((pr.argmax(-1) == yte).float().mean().item())
    probs.append(pr)
  return accs, probs


@pytest.mark.parametrize('name', ['toy', 'split_mnist_shape', 'permuted_mnist_shape'])
def test_end_of_run_accuracy_matches_reference_path(name, cuda_ops):
  g = torch.Generator().manual_seed(3)
  if name == 'toy':                  # experiments/vargp.py toy: C=4, D=2, M=20, lr=1e-2, full batch
    tasks, C, D = _toy_tasks(g)
    kw = dict(M=20, steps=300, B=100, lr=3e-2, beta=1.0, log_sigma=math.log(0.5))
  elif name == 'permuted_mnist_shape':   # experiments/vargp.py permuted_mnist: beta=1.64, all classes in every task
    tasks, C, D = _permuted_tasks(g)
    kw = dict(M=24, steps=80, B=128, lr=1e-2, beta=1.64, log_sigma=math.log(10.))
  else:                              # Split-MNIST shape (D=784, 10 output GPs, 2 classes per task), learned-lengthscale regime
    tasks, C, D = _mnist_like_tasks(g)
    kw = dict(M=20, steps=80, B=128, lr=1e-2, beta=10.0, log_sigma=math.log(10.))
  acc_o, pr_o = _run('oracle', tasks, C, D, **kw)
  acc_g, pr_g = _run('b200', tasks, C, D, **kw)
  print(name, 'oracle acc', [f'{a:.4f}' for a in acc_o], 'b200 acc', [f'{a:.4f}' for a in acc_g],
        'max |dp|', [f'{(a - b).abs().max().item():.2e}' for a, b in zip(pr_o, pr_g)])
  chance = 1.0 / C
  for t, (a, b) in enumerate(zip(acc_o, acc_g)):
    assert a > chance + 0.2, f'task {t}: the oracle run did not learn ({a:.3f})'
    assert abs(a - b) <= 0.005 + 1e-9, f'task {t}: accuracy {b:.4f} vs reference path {a:.4f} (> 0.5 pt)'


def test_fused_stepper_reaches_the_same_accuracy_at_the_benched_size(cuda_ops):
  """Full Split-MNIST shape (C=10, D=784, M=60 per task, 5 tasks, B=512): the graph-replayed tape-free training step
  (train.ElboStepper -> fused_step.FusedElbo, what bench.py times) against loss() + autograd + the same Yogi rule on the
  same kernels (the path the fixtures and the oracle runs above validate).  Same seeds -> same minibatches and noise
  streams; final per-task test accuracies must agree within 0.5 pt and the final parameters closely."""
  from vargp_b200.train import ElboStepper
  from vargp_b200.optim import FlatYogi
  g = torch.Generator().manual_seed(4)
  tasks, C, D = _mnist_like_tasks(g, n_tasks=5, n_train=1024, n_test=400)
  M, steps, B, lr, beta, H, F = 60, 60, 512, 1e-2, 10.0, 3, 10
  out = {}
  for arm in ('autograd', 'fused'):
    gi = torch.Generator().manual_seed(21)
    prev, prior, final = [], None, None
    torch.manual_seed(99)
    for t, ((xtr, ytr), _) in enumerate(tasks):
      p0 = _init_params(gi, xtr, C, D, M, math.log(10.), prior)
      gp = util.build_model(p0, prev, H, F, {}, 'cuda', torch.float32)
      N = xtr.size(0)
      xd, yd = xtr.cuda(), ytr.cuda()
      st = ElboStepper(gp, n_data=N, batch_size=B, beta=beta, lr=lr, use_graph=arm == 'fused', fused=arm == 'fused')
      assert (st.fused is not None) == (arm == 'fused')
      for s in range(steps):
        idx = torch.randperm(N, generator=gi)[:B].cuda()
        st.step(xd[idx], yd[idx])
      st.check_errors()
      cur = dict(p0, z=gp.z.detach().cpu(), u_mean=gp.u_mean.detach().cpu(), u_tril_vec=gp.u_tril_vec.detach().cpu(),
                 log_mean=gp.kernel.log_mean.detach().cpu(), log_logvar=gp.kernel.log_logvar.detach().cpu())
      final = (cur, list(prev))
      prev = prev + [dict(z=cur['z'], u_mean=cur['u_mean'], u_tril_vec=cur['u_tril_vec'])]
      prior = (cur['log_mean'], cur['log_logvar'])
    cur, prv = final
    ge = torch.Generator().manual_seed(5)
    accs = []
    for _, (xte, yte) in tasks:
      nz = dict(eps_theta=torch.randn(H, D + 1, generator=ge), eps_f=torch.randn(H, 40, C, xte.size(0), generator=ge))
      with torch.no_grad():
        gpe = util.build_model(cur, prv, H, 40, {}, 'cuda', torch.float32)
        pr = gpe.predict(xte.cuda(), noise={k: v.cuda() for k, v in nz.items()}).cpu()
      accs.append((pr.argmax(-1) == yte).float().mean().item())
    out[arm] = (accs, cur)
  print('autograd acc', [f'{a:.4f}' for a in out['autograd'][0]], 'fused acc', [f'{a:.4f}' for a in out['fused'][0]])
  for t, (a, b) in enumerate(zip(*[out[k][0] for k in ('autograd', 'fused')])):
    assert a > 0.3, f'task {t}: the run did not learn ({a:.3f})'
    assert abs(a - b) <= 0.005 + 1e-9, (t, a, b)
  errs = {k: util.relerr(out['fused'][1][k], out['autograd'][1][k]) for k in ('z', 'u_mean', 'u_tril_vec', 'log_mean')}
  print('final-parameter differences', {k: f'{v:.2e}' for k, v in errs.items()})
  for k, v in errs.items():            # 300 Yogi steps apart only by summation order of a few float atomics
    assert v < 5e-2, (k, v)
