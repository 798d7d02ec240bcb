"""Randomised cross-check against the LIVE reference (build container only; needs /root/reference):

    python tests/golden/sweep_reference.py [n_cases] [seed]

For n random problem shapes (C, D, M, t, B, H, F, flags; including the degenerate ends M = 1, B = 1, C = 1, H = 1) it
runs the unmodified reference VARGP in fp64 with pinned noise, and checks
  (1) the oracle restatement (asserted inside make_golden.run_reference: <= 1e-10), and
  (2) the product's host schedule -- vargp_b200.VARGP -> functional -> elbo.py (or composed.py for the variants) on the
      torch emulation of the kernel interface -- for the ELBO terms, all five gradients and predict().
The ten committed fixtures pin fixed shapes; this sweep is how the shape-generic claims were checked.  Prints one summary
line; the result of the last run is recorded in DESIGN.md section 5."""
import os
import random
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_golden as mg          # noqa: E402
from tests import util            # noqa: E402
from tests.emu_ops import EmuOps  # noqa: E402
from vargp_b200 import ops        # noqa: E402


def random_case(rng):
  C = rng.choice([1, 2, 3, 5])
  D = rng.choice([1, 2, 3, 7, 16, 33])
  M = rng.choice([1, 2, 3, 5, 8, 12])
  t = rng.choice([0, 0, 1, 1, 2, 3, 4])
  B = rng.choice([1, 2, 5, 17, 40])
  H = rng.choice([1, 2, 3])
  F = rng.choice([1, 2, 4])
  kw = dict(C=C, D=D, M=M, t=t, B=B, H=H, F=F, sigma=(D ** 0.5) * rng.uniform(0.6, 2.0), seed=rng.randrange(10 ** 6))
  r = rng.random()
  if r < 0.2 and t > 0:
    kw['ep_var_mean'] = False
  elif r < 0.35:
    kw.update(map_est=True, n_v=H, H=1)
  return kw


def sweep_retrain(n, rng):
  """Same for VARGPRetrain (var_gp/vargp_retrain.py): oracle asserted inside make_golden.run_reference_retrain, product
  host schedule checked here (terms, the gradients of the current AND the re-trained parameters, predict)."""
  from vargp_b200.synthetic import make_retrain_case
  worst = dict(terms=0.0, grads=0.0, probs=0.0)
  done = skipped = 0
  for i in range(n):
    kw = random_case(rng)
    kw = {k: v for k, v in kw.items() if k not in ('ep_var_mean', 'map_est', 'n_v')}
    kw['H'] = max(kw['H'], 1)
    kw['t'] = min(kw['t'], 3)
    try:
      ref = mg.run_reference_retrain(kw, torch.float64)
    except Exception as e:
      skipped += 1
      print(f'retrain case {i} {kw}: reference raised {type(e).__name__}: {str(e)[:80]}')
      continue
    params, retrain, prev, x, y, noise = make_retrain_case(dtype=torch.float64, **kw)
    gp = util.build_retrain_model(params, retrain, prev, kw['H'], kw['F'], 'cpu', torch.float64)
    terms, grads = util.run_retrain_model(gp, x, y, noise, ref['beta'], ref['Ntot'])
    for k in ('kl_hypers', 'kl_u', 'nll', 'total'):
      e = util.relerr(terms[k], ref[k])
      worst['terms'] = max(worst['terms'], e)
      assert e < 1e-7, (i, kw, k, e)
    assert set(grads) == set(ref['grads'])
    for k, g in ref['grads'].items():
      if g.abs().max() > 0:
        e = util.relerr(grads[k], g)
        worst['grads'] = max(worst['grads'], e)
        assert e < 1e-6, (i, kw, k, e)
    with torch.no_grad():
      probs = gp.predict(x, noise=noise)
    e = (probs - ref['probs']).abs().max().item()
    worst['probs'] = max(worst['probs'], e)
    assert e < 1e-8, (i, kw, e)
    done += 1
  print(f'retrain sweep: {done} cases checked, {skipped} skipped (reference raised); worst rel. error terms '
        f'{worst["terms"]:.1e}, grads {worst["grads"]:.1e}, worst abs. error probs {worst["probs"]:.1e}')


def sweep_dkl(n, rng):
  """DeepRBFKernel (var_gp/kernels.py:80-96) in the live reference vs the product's fused path through the `features`
  hook: same MLP weights, fp64, terms + every gradient (incl. kernel.phi.*) + predict."""
  import torch.nn as nn
  from var_gp.vargp import VARGP as RefVARGP
  from var_gp.kernels import DeepRBFKernel as RefDeep
  from var_gp.likelihoods import MulticlassSoftmax as RefLik
  from vargp_b200.vargp import VARGP
  from vargp_b200.kernels import DeepRBFKernel
  from vargp_b200.likelihoods import MulticlassSoftmax
  from vargp_b200.synthetic import make_case
  torch.set_default_dtype(torch.float64)
  worst = dict(terms=0.0, grads=0.0, probs=0.0)
  for i in range(n):
    C, Din, Fd = rng.choice([2, 3, 4]), rng.choice([5, 12, 20]), rng.choice([4, 8, 16])
    M, t, B, H, F = rng.choice([2, 5, 9]), rng.choice([0, 1, 2]), rng.choice([3, 17, 40]), rng.choice([1, 2]), rng.choice([1, 3])
    seed = rng.randrange(10 ** 6)
    params, prev, _, y, noise = make_case(C=C, D=Fd, M=M, t=t, B=B, H=H, F=F, sigma=1.5, seed=seed, dtype=torch.float64)
    g = torch.Generator().manual_seed(seed + 7)
    U = lambda *sh: torch.rand(*sh, generator=g, dtype=torch.float64)
    z, x = U(C, M, Din), U(B, Din)
    prev = [dict(p, z=U(C, M, Din)) for p in prev]
    torch.manual_seed(seed)
    phi = nn.Sequential(nn.Linear(Din, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, Fd)).double()
    out = {}
    for arm, (V, K, L) in dict(ref=(RefVARGP, RefDeep, RefLik), new=(VARGP, DeepRBFKernel, MulticlassSoftmax)).items():
      kern = K(Din, feature_size=Fd, prior_log_mean=params['prior_log_mean'].clone(),
               prior_log_logvar=params['prior_log_logvar'].clone())
      gp = V(z.clone(), kern, L(n_f=F), n_var_samples=H, prev_params=[{k: v.clone() for k, v in p.items()} for p in prev])
      gp = gp.double()
      gp.kernel.phi.load_state_dict(phi.state_dict())
      with torch.no_grad():
        gp.u_mean.copy_(params['u_mean']); gp.u_tril_vec.copy_(params['u_tril_vec'])
        gp.kernel.log_mean.copy_(params['log_mean']); gp.kernel.log_logvar.copy_(params['log_logvar'])
      draws = [noise['eps_theta']] + ([noise['eps_u']] if prev else []) + [noise['eps_f']]
      if arm == 'ref':
        with mg.pinned_noise(draws):
          kl_h, kl_u, nll = gp.loss(x, y)
      else:
        kl_h, kl_u, nll = gp.loss(x, y, noise=noise)
      total = 1.7 * kl_h + kl_u + 10. * nll
      gp.zero_grad()
      total.backward()
      grads = {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in gp.named_parameters()}
      with torch.no_grad():
        if arm == 'ref':
          with mg.pinned_noise([noise['eps_theta'], noise['eps_f']]):
            probs = gp.predict(x)
        else:
          probs = gp.predict(x, noise=noise)
      out[arm] = (dict(kl_u=kl_u.detach(), nll=nll.detach(), total=total.detach()), grads, probs)
    for k, v in out['ref'][0].items():
      e = util.relerr(out['new'][0][k], v)
      worst['terms'] = max(worst['terms'], e)
      assert e < 1e-7, (i, k, e)
    scale = max(v.norm().item() for v in out['ref'][1].values())
    for k, v in out['ref'][1].items():
      e = ((out['new'][1][k] - v).norm() / max(v.norm().item(), 1e-6 * scale)).item()   # (the last bias has an exactly zero gradient: judged on scale)
      worst['grads'] = max(worst['grads'], e)
      assert e < 1e-5, (i, k, e)
    e = (out['new'][2] - out['ref'][2]).abs().max().item()
    worst['probs'] = max(worst['probs'], e)
    assert e < 1e-8, (i, e)
  print(f'dkl sweep: {n} cases checked; worst rel. error terms {worst["terms"]:.1e}, grads {worst["grads"]:.1e}, '
        f'worst abs. error probs {worst["probs"]:.1e}')


def sweep_utils(n, rng):
  """var_gp/train_utils.py (EarlyStopper, compute_bwt, compute_accuracy, compute_acc_ent) of the live reference --
  imported with inert stand-ins for its absent logging / optimizer dependencies (torch_optimizer, wandb) -- against
  vargp_b200.train on random score sequences, accuracy matrices and a fixed classifier."""
  import types
  for mod in ('torch_optimizer', 'wandb'):
    sys.modules.setdefault(mod, types.ModuleType(mod))
  import var_gp.train_utils as ref
  from torch.utils.data import TensorDataset
  from vargp_b200 import train as new

  class FixedClf:                                    # predict() = softmax of a fixed linear map: deterministic on both sides
    def __init__(self, D, C, g):
      self.Wm = torch.randn(D, C, generator=g)
      self.z = torch.zeros(1)
    def predict(self, x, noise=None):
      return torch.softmax(x @ self.Wm, dim=-1)

  for i in range(n):
    patience, delta = rng.choice([-1, 0, 1, 3, 10]), rng.choice([1e-4, 1e-2, 0.0])
    a, b = ref.EarlyStopper(patience=patience, delta=delta), new.EarlyStopper(patience=patience, delta=delta)
    for step in range(rng.randrange(1, 30)):
      if a.is_done():
        break
      score = round(rng.random(), rng.choice([1, 2, 6]))
      a(score, step); b(score, step)
      assert a.is_done() == b.is_done() and a.info() == b.info(), (i, step)
    T = rng.randrange(1, 7)
    acc = torch.rand(T, T, generator=torch.Generator().manual_seed(i))
    ra, rb = ref.compute_bwt(acc), new.compute_bwt(acc)
    assert (torch.isnan(ra) and torch.isnan(rb)) or torch.equal(ra, rb), (i, ra, rb)
    g = torch.Generator().manual_seed(1000 + i)
    N, D, C, bs = rng.randrange(1, 300), rng.randrange(1, 9), rng.randrange(2, 6), rng.choice([1, 7, 64, 512])
    x, y = torch.randn(N, D, generator=g), torch.randint(0, C, (N,), generator=g)
    clf = FixedClf(D, C, g)
    assert ref.compute_accuracy(TensorDataset(x, y), clf, batch_size=bs) == new.compute_accuracy(new.TensorTask(x, y), clf, batch_size=bs)
    (a1, e1), (a2, e2) = ref.compute_acc_ent(TensorDataset(x, y), clf, batch_size=bs), new.compute_acc_ent(new.TensorTask(x, y), clf, batch_size=bs)
    assert a1 == a2 and abs(e1 - e2) <= 1e-5 * max(1.0, abs(e1)), (i, e1, e2)
  print(f'utils sweep: {n} cases checked (EarlyStopper traces, compute_bwt, compute_accuracy, compute_acc_ent identical)')


def sweep_init(n, rng):
  """Construction parity under the same seed: toy data (var_gp/datasets.py:21-51 vs synthetic.toy_data), and
  create_clf of VARGP (plain / with prev_params / dkl) and VARGPRetrain -- inducing points picked by the same randperm
  draws, same initial parameters, same hyper-prior hand-off: state dicts must be equal bit for bit."""
  from var_gp.datasets import ToyDataset
  from var_gp.vargp import VARGP as RefVARGP
  from var_gp.vargp_retrain import VARGPRetrain as RefRetrain
  from vargp_b200.vargp import VARGP
  from vargp_b200.vargp_retrain import VARGPRetrain
  from vargp_b200.synthetic import toy_data
  from vargp_b200.train import TensorTask

  def same(a, b, what):
    assert list(a.keys()) == list(b.keys()), (what, list(a.keys()), list(b.keys()))
    for k in a:
      assert torch.equal(a[k], b[k]), (what, k)

  for i in range(n):
    seed = rng.randrange(10 ** 6)
    N_K = rng.choice([5, 20, 50])
    torch.manual_seed(seed); ds = ToyDataset(N_K=N_K)
    torch.manual_seed(seed); X, Y = toy_data(N_K=N_K)
    assert torch.equal(ds.data, X) and torch.equal(ds.targets, Y), i
    M, n_f, n_v = rng.choice([3, 8, 20]), rng.choice([2, 10]), rng.choice([1, 3])
    dkl, map_est, ep = rng.random() < 0.3, rng.random() < 0.3, rng.random() < 0.7
    ds.filter_by_class([0, 1])
    task = TensorTask.for_classes(X, Y, (0, 1))
    kw = dict(M=M, n_f=n_f, n_var_samples=n_v, ep_var_mean=ep, map_est_hypers=map_est, dkl=dkl)
    torch.manual_seed(seed + 1); a = RefVARGP.create_clf(ds, **kw)
    torch.manual_seed(seed + 1); b = VARGP.create_clf(task, **kw)
    same(a.state_dict(), b.state_dict(), 'task 0')
    # second task: the previous state dict is handed over (the reference pops its kernel.* keys: give it a copy)
    sd = {k: v.clone() + 0.01 for k, v in a.state_dict().items()}
    ds.filter_by_class([2, 3])
    task = TensorTask.for_classes(X, Y, (2, 3))
    torch.manual_seed(seed + 2); a = RefVARGP.create_clf(ds, prev_params=[dict(sd)], **kw)
    torch.manual_seed(seed + 2); b = VARGP.create_clf(task, prev_params=[dict(sd)], **kw)
    same(a.state_dict(), b.state_dict(), 'task 1')
    pa, pb = a.prev_params, b.prev_params
    assert len(pa) == len(pb) == 1 and torch.equal(pa[0]['z'], pb[0]['z']) and torch.equal(pa[0]['u_mean'], pb[0]['u_mean'])
    if not dkl:
      rkw = dict(M=M, n_f=n_f, n_var_samples=n_v)
      torch.manual_seed(seed + 3); a = RefRetrain.create_clf(ds, prev_params=[dict(sd)], **rkw)
      torch.manual_seed(seed + 3); b = VARGPRetrain.create_clf(task, prev_params=[dict(sd)], **rkw)
      same(a.state_dict(), b.state_dict(), 'retrain')
  print(f'init sweep: {n} cases checked (toy data, create_clf state dicts of VARGP / dkl / prev_params / VARGPRetrain bit-identical)')


def main():
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
  rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
  refmods = mg.load_reference()
  ops.set_ops(EmuOps())
  if os.environ.get('VARGP_SWEEP', 'vargp') == 'retrain':
    return sweep_retrain(n, rng)
  if os.environ.get('VARGP_SWEEP') == 'dkl':
    return sweep_dkl(n, rng)
  if os.environ.get('VARGP_SWEEP') == 'utils':
    return sweep_utils(n, rng)
  if os.environ.get('VARGP_SWEEP') == 'init':
    return sweep_init(n, rng)
  worst = dict(terms=0.0, grads=0.0, probs=0.0)
  done = skipped = 0
  for i in range(n):
    kw = random_case(rng)
    try:
      ref = mg.run_reference(refmods, kw, torch.float64)
    except Exception as e:          # e.g. the reference's own no-jitter Cholesky of S_<t failing on a degenerate draw
      skipped += 1
      print(f'case {i} {kw}: reference raised {type(e).__name__}: {str(e)[:80]}')
      continue
    params, prev, x, y, noise, n_v, F, flags = util.case_tensors(kw, torch.float64)
    gp = util.build_model(params, prev, n_v, F, flags, 'cpu', torch.float64)
    terms, grads = util.run_model(gp, x, y, noise, ref['beta'], ref['Ntot'])
    for k in ('kl_hypers', 'kl_u', 'nll', 'total'):
      if ref[k].abs() > 0:
        e = util.relerr(terms[k], ref[k])
        worst['terms'] = max(worst['terms'], e)
        assert e < 1e-7, (i, kw, k, e)
    for k in util.GRAD_KEYS:
      if ref['grads'][k].abs().max() > 0:
        e = util.relerr(grads[k], ref['grads'][k])
        worst['grads'] = max(worst['grads'], e)
        assert e < 1e-6, (i, kw, k, e)
    with torch.no_grad():
      probs = gp.predict(x, noise=noise)
    e = (probs - ref['probs']).abs().max().item()
    worst['probs'] = max(worst['probs'], e)
    assert e < 1e-8, (i, kw, e)
    done += 1
  print(f'sweep: {done} cases checked, {skipped} skipped (reference raised); worst rel. error terms {worst["terms"]:.1e}, '
        f'grads {worst["grads"]:.1e}, worst abs. error probs {worst["probs"]:.1e}')


if __name__ == '__main__':
  main()
