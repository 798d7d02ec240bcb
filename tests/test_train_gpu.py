"""Training-step driver on the GPU: fused flat Yogi vs the reference-rule Yogi, graphed vs eager step."""
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def test_flat_yogi_matches_foreach_yogi(cuda_ops):
  from vargp_b200.optim import Yogi, FlatYogi
  torch.manual_seed(0)
  shapes = [(10, 60, 784), (10, 60, 1), (10, 1830), (785,), (785,)]
  pa = [torch.nn.Parameter(torch.randn(*s, device='cuda')) for s in shapes]
  pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
  oa, ob = Yogi(pa, lr=3e-3), FlatYogi(pb, lr=3e-3)
  for it in range(25):
    gs = [torch.randn_like(p) * (10.0 ** (it % 5 - 2)) for p in pa]
    ob.zero_grad()
    for p, q, g in zip(pa, pb, gs):
      p.grad = g.clone()
      q.grad.add_(g)
    oa.step(); ob.step()
  for p, q in zip(pa, pb):
    assert util.relerr(q, p) < 1e-5


def _model(seed=3):
  from vargp_b200.synthetic import make_case
  params, prev, x, y, noise = make_case(C=10, D=784, M=20, t=2, B=256, sigma=10., seed=seed)
  return util.build_model(params, prev, 3, 10, {}, 'cuda', torch.float32), x.cuda(), y.cuda()


def test_graphed_step_trains_like_eager(cuda_ops):
  """Same data, same seeds: the graph-replayed step and the eager step follow statistically identical
  trajectories (the RNG streams differ, so compare the objective after 30 steps, not bits)."""
  from vargp_b200.train import ElboStepper
  res = {}
  for mode in (False, True):
    gp, x, y = _model()
    st = ElboStepper(gp, n_data=2560, batch_size=256, beta=1.0, lr=1e-2, use_graph=mode)
    torch.manual_seed(7)
    first = None
    for i in range(40):
      kl_h, kl_u, nll = st.step(x, y)
      tot = float(kl_h + kl_u + 10. * nll)
      first = tot if first is None else first
    gp.check_errors()
    res[mode] = (first, tot)
    assert tot < first, (mode, first, tot)            # the ELBO objective goes down
    assert all(torch.isfinite(p).all() for p in gp.parameters())
  assert abs(res[True][1] - res[False][1]) < 0.2 * abs(res[False][1])   # different RNG streams
  assert st.launches_per_step and st.launches_per_step < 80


def test_train_driver_continual_toy(cuda_ops):
  """vargp_b200.train.train (experiments/vargp.py:14-73 restated): two toy tasks through create_clf -> graphed
  steps -> on-device evaluation -> early stopper; the returned state dicts carry the reference's keys, feed the
  next task as prev_params and reload through create_clf + load_state_dict (the notebooks' protocol)."""
  from vargp_b200.train import TensorTask, train, compute_accuracy, compute_acc_ent, compute_bwt
  from vargp_b200.vargp import VARGP
  g = torch.Generator().manual_seed(0)
  centers = torch.tensor([[-1., -1.], [1., 1.], [-1., 1.], [1., -1.]])
  y = torch.arange(4).repeat_interleave(60)
  x = centers[y] + 0.4 * torch.randn(240, 2, generator=g)
  tasks = [TensorTask.for_classes(x, y, c) for c in ((0, 1), (2, 3))]
  torch.manual_seed(1)
  prev, accs = [], []
  for t, ds in enumerate(tasks):
    sd = train(t, ds, ds, ds, epochs=60, M=12, lr=3e-2, eval_interval=20, patience=5, prev_params=prev,
               batch_size=64, device='cuda')
    assert set(sd) >= {'z', 'u_mean', 'u_tril_vec', 'kernel.log_mean', 'kernel.log_logvar'}
    assert sd['z'].shape == (4, 12, 2)                        # all four output GPs exist from task 0 (vargp.py:204)
    gp = VARGP.create_clf(ds, M=12, prev_params=prev).cuda()
    gp.load_state_dict(sd)
    prev.append(sd)
    accs.append([compute_accuracy(d, gp) for d in tasks])
  acc, ent = compute_acc_ent(TensorTask(x, y), gp)
  assert accs[0][0] > 0.8 and accs[1][1] > 0.8 and acc > 0.6 and ent > 0
  assert compute_bwt(torch.tensor(accs)).abs() < 0.5


def test_prefetched_inputs_give_the_same_steps(cuda_ops):
  """step(x, y, prefetch=next) copies the next minibatch on the copy stream while the step runs; the trajectory must be
  the same as feeding the same pinned batches directly (same graph, same RNG stream)."""
  from vargp_b200.train import ElboStepper
  g = torch.Generator().manual_seed(11)
  xs = torch.rand(4, 256, 784, generator=g).pin_memory()
  ys = torch.randint(0, 10, (4, 256), generator=g).pin_memory()
  out = {}
  for mode in ('direct', 'prefetch'):
    gp, _, _ = _model()
    st = ElboStepper(gp, n_data=2560, batch_size=256, beta=1.0, lr=1e-2, use_graph=True)
    torch.manual_seed(7)
    vals = []
    for i in range(12):
      if mode == 'prefetch':
        st.step(xs[i % 4], ys[i % 4], prefetch=(xs[(i + 1) % 4], ys[(i + 1) % 4]))
      else:
        st.step(xs[i % 4], ys[i % 4])
      vals.append(st.terms_vec.clone())
    gp.check_errors()
    out[mode] = (torch.stack(vals).cpu(), gp.z.detach().clone().cpu())
    if mode == 'prefetch':
      assert st._copy_stream is not None
  # same graph, same noise; a few reductions use float atomics, so equal up to summation order, not bits
  assert util.relerr(out['prefetch'][0], out['direct'][0]) < 1e-4
  assert util.relerr(out['prefetch'][1], out['direct'][1]) < 1e-4
  assert util.relerr(out['direct'][0][1:], out['direct'][0][:-1]) > 1e-4      # and the steps really differ from each other
  assert out['direct'][0].shape == (12, 3) and torch.isfinite(out['direct'][0]).all()


def test_predict_at_notebook_sizes(cuda_ops):
  """The paper's evaluation recipe (notebooks: n_var_samples = 20 hyper samples, n_f = 50 likelihood samples) on a
  Split-MNIST-shaped model: predict() vs the fp64 CPU oracle, 1e-5 absolute."""
  from oracle import vargp_oracle as orc
  kw = dict(C=10, D=784, M=20, t=2, B=96, H=20, F=50, sigma=10., seed=17)
  params, prev, x, y, noise = orc.make_case(dtype=torch.float32, **kw)
  gp = util.build_model(params, prev, 20, 50, {}, 'cuda', torch.float32)
  with torch.no_grad():
    probs = gp.predict(x.cuda(), noise={k: v.cuda() for k, v in noise.items()})
  p64, prev64, x64, _, n64 = orc.make_case(dtype=torch.float64, **kw)
  ref = orc.predict(p64, prev64, x64, n64, n_v=20)
  assert tuple(probs.shape) == (96, 10)
  assert (probs.double().cpu() - ref).abs().max().item() < 1e-5
  assert (probs.sum(-1) - 1).abs().max().item() < 1e-5


def test_first_graphed_step_is_exactly_one_update(cuda_ops):
  """The warm-up steps behind the graph capture must leave no trace: after exactly one `step()` from the
  same seed the graphed and the eager stepper hold the same parameters, Yogi moments and bias-correction powers."""
  from vargp_b200.train import ElboStepper
  out = {}
  for mode in (False, True):
    gp, x, y = _model()
    st = ElboStepper(gp, n_data=2560, batch_size=256, beta=1.0, lr=1e-2, use_graph=mode)
    torch.manual_seed(7)
    st.step(x, y)
    torch.cuda.synchronize()
    out[mode] = (st.opt.flat_p.clone(), st.opt.m.clone(), st.opt.v.clone(), st.opt.pows.clone(), st.terms_vec.clone())
  for a, b in zip(out[True], out[False]):
    assert util.relerr(a, b) < 1e-5
  assert torch.equal(out[True][3], out[False][3])                 # b1^1, b2^1: exactly one Yogi step


def test_stepper_reports_non_pd_after_the_step(cuda_ops):
  """A non-PD Gram does not raise inside the graph-replayed step, but `stepper.check_errors()` (called by train() once
  per epoch) does; loss()/predict() outside the stepper keep raising immediately (var_gp/gp_utils.py:10)."""
  from vargp_b200.train import ElboStepper
  gp, x, y = _model()
  st = ElboStepper(gp, n_data=2560, batch_size=256, use_graph=True)
  st.step(x, y)
  st.check_errors()
  assert gp.sync_errors
  with torch.no_grad():
    gp.z[0, 0].fill_(float('nan'))
  st.step(x, y)
  with pytest.raises(torch.linalg.LinAlgError):
    st.check_errors()
  with pytest.raises(torch.linalg.LinAlgError):
    gp.predict(x)


@pytest.mark.parametrize('t', [0, 2])
def test_fused_step_matches_autograd_step(cuda_ops, t):
  """fused_step.FusedElbo (tape-free value-and-gradient, gradients written straight into the flat Yogi buffer) vs
  VARGP.loss + autograd through the same kernels: same seed -> same loss terms and gradients, and after a few Yogi steps
  the same parameters."""
  from vargp_b200.train import ElboStepper
  from vargp_b200.synthetic import make_case
  out = {}
  for fused in (False, True):
    params, prev, x, y, _ = make_case(C=10, D=784, M=20, t=t, B=256, sigma=10., seed=3)
    gp = util.build_model(params, prev, 3, 10, {}, 'cuda', torch.float32)
    st = ElboStepper(gp, n_data=2560, batch_size=256, beta=1.7, lr=1e-2, use_graph=False, fused=fused)
    assert (st.fused is not None) == fused
    torch.manual_seed(7)
    st._load_inputs(x.cuda(), y.cuda())
    st._grad_body()
    torch.cuda.synchronize()
    g1, t1 = st.opt.flat_g.clone(), st.terms_vec.clone()
    torch.manual_seed(7)
    for i in range(4):
      st.step(x.cuda(), y.cuda())
    st.check_errors()
    out[fused] = (g1, t1, st.opt.flat_p.clone(), st.terms_vec.clone())
  assert torch.equal(out[True][1], out[False][1])                 # forward: same kernels, deterministic reductions
  assert util.relerr(out[True][0], out[False][0]) < 1e-5
  # per parameter (the flat buffer is dominated by z)
  sizes = [10 * 20 * 784, 200, 10 * 210, 785, 785]          # (the flat buffer is padded to a multiple of 4 elements)
  for k, (a, b) in enumerate(zip(out[True][0][:sum(sizes)].split(sizes), out[False][0][:sum(sizes)].split(sizes))):
    assert util.relerr(a, b) < 1e-5, k
  assert util.relerr(out[True][2], out[False][2]) < 1e-5
  assert util.relerr(out[True][3], out[False][3]) < 1e-5
