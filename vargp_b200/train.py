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

import os

from .dist import shard_coef, shard_loss
from .optim import FlatYogi

USE_PRIORITY = os.environ.get('VARGP_PRIO', '1') != '0'
GRAPH_NCCL = os.environ.get('VARGP_GRAPH_NCCL', '1') != '0'      # capture the factor-shard collectives (P >= 512) into the step graph
GRAPH_TAIL = os.environ.get('VARGP_GRAPH_TAIL', '0') != '0'      # N > 1: also capture the gradient all-reduce + Yogi tail
FUSED_STEP = os.environ.get('VARGP_FUSED_STEP', '1') != '0'
PEER_ALLREDUCE = os.environ.get('VARGP_PEER_ALLREDUCE', '0') != '0'
NO_ALLREDUCE = os.environ.get('VARGP_NO_ALLREDUCE', '0') != '0'


class ElboStepper:
  def __init__(self, gp, n_data, batch_size, beta=1.0, lr=1e-2, world_size=1, use_graph=True, optimizer=None,
               shard_factor=None, fused=None, peer=None, graph_tail=None):
    """shard_factor: also shard the replicated O(P^3) factor stage over the ranks (elbo.FactorShard); default: on
    when world_size > 1 and the task has at least 512 inducing points per class (below that the three extra
    collectives cost more than the replicated work).
    peer: data parallel -- gradient all-reduce + Yogi as one kernel pair over NVLink peer memory (csrc/peer.cu) instead
    of NCCL all_reduce + yogi_step (opt-in: VARGP_PEER_ALLREDUCE=1; measured on par with NCCL at N = 2, behind it at N = 8).
    graph_tail: N > 1 -- capture the NCCL gradient all-reduce + Yogi step into the step graph as well (opt-in:
    VARGP_GRAPH_TAIL=1; measured slower than launching them eagerly behind the graph).
    fused: take the tape-free value-and-gradient path of `fused_step.FusedElbo` (default: whenever the model is
    eligible -- plain RBF kernel, sampled hypers, ep_var_mean=True); False keeps loss() + autograd."""
    self.gp, self.n_data, self.beta, self.world = gp, n_data, beta, world_size
    self.global_batch = batch_size * world_size
    self.opt = optimizer or FlatYogi(gp.parameters(), lr=lr)
    P = (gp.n_prev + 1) * gp.M
    if shard_factor is None:
      shard_factor = world_size > 1 and P >= 512
    self.shard = None
    if shard_factor and world_size > 1 and gp.var_mean_mask == 1.0:
      from .elbo import FactorShard
      self.shard = FactorShard()
    # N > 1, measured on 2 and 8 B200s at the Split-MNIST shape (profiles/r2_multi_gpu_ab.txt): the fastest tail is the
    # round-1 one -- the graph ends with the backward pass; NCCL all_reduce and the Yogi step follow eagerly (the host
    # enqueues them while the graph runs): 1.012 ms/step at N = 2, 1.027 at N = 8, against 0.998 without any exchange.
    # Capturing the NCCL all-reduce + Yogi into the graph (VARGP_GRAPH_TAIL=1) measured 1.031 (N = 2); the fused
    # peer-memory all-reduce + Yogi kernel pair of csrc/peer.cu (VARGP_PEER_ALLREDUCE=1, always inside the graph) 1.017
    # (N = 2) / 1.044 (N = 8: a one-shot all-reduce reads 7 x 1.9 MB over NVLink per rank).  Both stay available and
    # tested (tests/test_dist_gpu.py); neither is the default.  The factor-shard collectives (P >= 512) ARE captured by
    # default (VARGP_GRAPH_NCCL=0 restores the round-1 behaviour: no graph at all with the factor stage sharded).
    nccl = world_size > 1 and dist.is_initialized() and dist.get_backend() == 'nccl'
    self.graph_nccl = GRAPH_NCCL and nccl
    self.graph_tail = (GRAPH_TAIL if graph_tail is None else graph_tail) and nccl
    if self.shard is not None and not self.graph_nccl:
      use_graph = False                    # collectives inside forward / backward: launch eagerly
    # gradient all-reduce + Yogi fused over peer memory (csrc/peer.cu) when asked for, the optimizer is the flat one and
    # symmetric memory is available
    self.peer = False
    if ((PEER_ALLREDUCE if peer is None else peer) and world_size > 1 and dist.is_initialized() and dist.get_backend() == 'nccl' and
        hasattr(self.opt, 'enable_peer_allreduce')):
      self.peer = self.opt.peer is not None or self.opt.enable_peer_allreduce()
    self.use_graph = use_graph
    self.graph = None
    dev = next(gp.parameters()).device
    D = gp.z.size(-1) if not hasattr(gp.kernel, 'phi') else gp.kernel.phi[0].in_features
    self.x = torch.empty(batch_size, D, device=dev)
    self.y = torch.empty(batch_size, dtype=torch.int64, device=dev)
    self.coef = torch.tensor(shard_coef(beta, n_data, self.global_batch, world_size, self.shard is not None), device=dev)
    self.terms = None
    self.launches_per_step = None
    self.terms_vec = None        # (kl_hypers, kl_u, nll) of the last step as one (3,) device tensor
    self._pf = None              # prefetch state: (x_src, y_src, x_staging, y_staging, ready event)
    self._copy_stream = None
    from . import fused_step
    if fused is None:
      fused = FUSED_STEP
    self.fused = fused_step.FusedElbo(gp, self.coef, self.shard) if (fused and fused_step.eligible(gp)) else None
    self.info = None             # Cholesky status words of the last step (the graph's static buffer in graph mode)
    self._host_terms = None      # pinned (2, 3) ring of the loss terms fetched by `fetch_terms_async`
    self._host_ev = [None, None]
    self._host_n = 0

  def _grad_body(self):
    if self.fused is not None:           # tape-free path: every gradient is overwritten in place, nothing to zero
      self.terms_vec = tv = self.fused.value_and_grad(self.x, self.y)
      self.info = self.gp._last_info
      return (tv[0], tv[1], tv[2])
    self.opt.zero_grad()
    gp = self.gp
    sync, gp.sync_errors, gp.factor_shard = gp.sync_errors, False, self.shard   # no host sync inside the step ...
    try:
      kl_h, kl_u, nll = gp.loss(self.x, self.y)
    finally:
      gp.sync_errors, gp.factor_shard = sync, None       # ... but predict() / loss() outside it raise as before
    self.info = gp._last_info
    loss = shard_loss(kl_h, kl_u, nll, self.beta, self.n_data, self.global_batch, self.world, coef=self.coef)
    loss.backward()
    terms = (kl_h.detach(), kl_u.detach(), nll.detach())
    self.terms_vec = torch.stack(terms)          # the three terms as one (3,) tensor: a single D2H for loggers
    return terms

  def _finish(self):
    if NO_ALLREDUCE:                     # measurement only (VARGP_NO_ALLREDUCE=1): ranks step without exchanging gradients
      self.opt.step()
      return
    if self.world > 1 and self.peer:
      self.opt.step_allreduce()          # gradient all-reduce + Yogi as one kernel pair over NVLink peer memory
      return
    if self.world > 1:
      dist.all_reduce(self.opt.flat_g, op=dist.ReduceOp.SUM)
    self.opt.step()

  def _body(self):
    terms = self._grad_body()
    self._finish()
    return terms

  def _snapshot(self):
    """Everything a warm-up step mutates: parameters, optimizer state, the device RNG stream."""
    opt = self.opt
    if hasattr(opt, 'flat_p'):
      st = [t.clone() for t in (opt.flat_p, opt.flat_g, opt.m, opt.v, opt.pows)]
    else:
      import copy
      st = ([p.detach().clone() for p in self.gp.parameters()], copy.deepcopy(opt.state_dict()), getattr(opt, '_t', None))
    return st, torch.cuda.get_rng_state(self.x.device)

  def _restore(self, snap):
    st, rng = snap
    opt = self.opt
    with torch.no_grad():
      if hasattr(opt, 'flat_p'):
        for dst, src in zip((opt.flat_p, opt.flat_g, opt.m, opt.v, opt.pows), st):
          dst.copy_(src)
      else:
        for p, q in zip(self.gp.parameters(), st[0]):
          p.copy_(q)
        opt.load_state_dict(st[1])
        if st[2] is not None:
          opt._t = st[2]
    torch.cuda.set_rng_state(rng, self.x.device)

  def _capture(self):
    """The whole step -- zero_grad, loss, backward, (N > 1: the NCCL gradient all-reduce and the factor-shard
    collectives,) Yogi -- becomes ONE CUDA graph.  The three warm-up steps that prime the allocator pools, the lazy
    inits and the NCCL communicator run on real data, so everything they mutate (parameters, Yogi moments and
    bias-correction powers, the RNG stream) is snapshotted before and restored after: the first `step()` applies
    exactly one update, like the eager loop of experiments/vargp.py:30-37."""
    from . import ops as _ops_mod
    ops = self._ops = _ops_mod.get_ops()
    # The capture stream has HIGH priority: its kernel nodes inherit it, the nodes of the side branches
    # (elbo._Fork: the minibatch-sized Kzx / Gz1 GEMMs, a few hundred CTAs each) keep the default, lowest one.  The
    # block scheduler then hands SMs to the critical chain (Kzz -> Cholesky -> whitening ...: many short kernels of
    # <= 30..270 CTAs) first, and the side GEMMs fill what is left instead of making the chain queue behind them.
    snap = self._snapshot()
    s = torch.cuda.Stream(priority=-1 if USE_PRIORITY else 0)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
      for _ in range(3):
        self._body()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    self._restore(snap)
    torch.cuda.synchronize()
    self.graph = torch.cuda.CUDAGraph()
    self._tail_in_graph = self.world == 1 or self.graph_tail or bool(self.peer)
    n0 = ops.launch_count()
    with torch.cuda.graph(self.graph, stream=s):
      self.terms = self._body() if self._tail_in_graph else self._grad_body()
    self._graph_terms_vec = self.terms_vec      # static outputs of the graph
    self._graph_info = self.info
    self.launches_per_step = ops.launch_count() - n0 + (0 if self._tail_in_graph else 2)

  def _load_inputs(self, x, y):
    """Inputs of this step into the static buffers: from the staging buffers if (x, y) are the tensors announced
    by the previous call's `prefetch` (their H2D copy ran on the copy stream beside the previous step), else by a
    direct copy."""
    pf, self._pf = self._pf, None
    if pf is not None and pf[0].data_ptr() == x.data_ptr() and pf[0].shape == x.shape and pf[1].data_ptr() == y.data_ptr():
      cur = torch.cuda.current_stream()
      cur.wait_event(pf[4])
      self.x.copy_(pf[2], non_blocking=True)
      self.y.copy_(pf[3], non_blocking=True)
      self._staging_free.record(cur)
    else:
      self.x.copy_(x, non_blocking=True)
      self.y.copy_(y, non_blocking=True)

  def _prefetch(self, x, y):
    if self._copy_stream is None:
      self._copy_stream = torch.cuda.Stream()
      self._xs, self._ys = torch.empty_like(self.x), torch.empty_like(self.y)
      self._staging_free = torch.cuda.Event()
      self._staging_free.record(torch.cuda.current_stream())
    cs = self._copy_stream
    cs.wait_event(self._staging_free)              # the previous step has read the staging buffers
    with torch.cuda.stream(cs):
      self._xs.copy_(x, non_blocking=True)
      self._ys.copy_(y, non_blocking=True)
      ready = torch.cuda.Event()
      ready.record(cs)
    self._pf = (x, y, self._xs, self._ys, ready)

  def step(self, x, y, prefetch=None):
    """One optimisation step on the minibatch (x, y) (device or pinned-host tensors).  Returns the three
    loss terms as 0-d device tensors (valid until the next call); `terms_vec` holds them as one (3,) tensor.

    prefetch=(x_next, y_next): the NEXT call's inputs (pinned host tensors); their host-to-device copy is issued on
    a copy stream right after this step has been launched, so it overlaps the step instead of preceding the next
    (the caller must leave those host tensors untouched until then, as with any non_blocking copy)."""
    self._load_inputs(x, y)
    if not self.use_graph:
      out = self._body()
    else:
      if self.graph is None:
        self._capture()      # neither the warm-up steps nor the capture advance parameters / optimizer / RNG
      self.graph.replay()
      if not self._tail_in_graph:
        self._finish()
      out, self.terms_vec, self.info = self.terms, self._graph_terms_vec, self._graph_info
    if prefetch is not None:
      self._prefetch(*prefetch)
    return out

  def check_errors(self):
    """Raise torch.linalg.LinAlgError if the Cholesky of the LAST training step hit a non-positive pivot (one device
    sync).  The reference raises inside the step (var_gp/gp_utils.py:10); the graph-replayed step cannot, so
    `train()` calls this once per epoch."""
    if self.info is not None:
      bad = int(self.info.max().item())
      if bad:
        from .vargp import CholeskyError
        raise CholeskyError(f'linalg.cholesky: the input is not positive-definite '
                            f'(leading minor of order {bad} is not positive-definite)')

  def fetch_terms_async(self):
    """Queue the device-to-host copy of this step's (kl_hypers, kl_u, nll) into a pinned two-slot ring without
    waiting for it; `host_terms(lag)` hands out a completed slot.  Loggers that read the terms one step late never
    stall the launch of the next step (the synchronous read costs ~4 % of a Split-MNIST-shape step)."""
    if self._host_terms is None:
      self._host_terms = torch.empty(2, 3, dtype=self.terms_vec.dtype).pin_memory()
    k = self._host_n & 1
    self._host_terms[k].copy_(self.terms_vec, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    self._host_ev[k] = ev
    self._host_n += 1

  def host_terms(self, lag=1):
    """(3,) pinned host tensor with the loss terms of the step `lag` fetches ago (0 = the last one: waits for it);
    None while fewer steps have been fetched."""
    if lag not in (0, 1) or self._host_n <= lag:
      return None
    k = (self._host_n - 1 - lag) & 1
    self._host_ev[k].synchronize()
    return self._host_terms[k]


# ------------------------------------------------------------------------------------------------------
# Training / evaluation driver: the per-task loop of experiments/vargp.py:14-73 and the
# helpers of var_gp/train_utils.py:21-98, restated for device-resident data -- tensors instead of Dataset/DataLoader
# objects, the graph-replayed ElboStepper as the inner loop, accuracies accumulated on the device (one host sync
# per evaluation instead of one per batch).  Logging (wandb / tensorboard) and checkpoint files stay with the caller.
# ------------------------------------------------------------------------------------------------------
class TensorTask:
  """Device-friendly stand-in for the reference's datasets (var_gp/datasets.py:70-138): `data` (N, D) and `targets`
  (N,) hold the WHOLE benchmark, `task_ids` (optional index tensor) selects the items of the current task -- like
  SplitMNIST, whose `targets` stay unfiltered, which is why create_clf sizes the model for every class from task 0
  on (var_gp/vargp.py:204).  Indexing returns (x, y) of the selected items; `.x` / `.y` are the selected tensors."""

  def __init__(self, data, targets, task_ids=None):
    self.data, self.targets, self.task_ids = data, targets, task_ids

  @classmethod
  def for_classes(cls, data, targets, classes):
    mask = torch.zeros_like(targets, dtype=torch.bool)
    for c in classes:
      mask |= targets == c
    return cls(data, targets, mask.nonzero().squeeze(-1))

  @property
  def x(self):
    return self.data if self.task_ids is None else self.data[self.task_ids]

  @property
  def y(self):
    return self.targets if self.task_ids is None else self.targets[self.task_ids]

  def __len__(self):
    return self.data.size(0) if self.task_ids is None else self.task_ids.numel()

  def __getitem__(self, idx):
    if self.task_ids is not None:
      idx = self.task_ids[idx]
    return self.data[idx], self.targets[idx]


def set_seeds(seed=None):
  """var_gp/train_utils.py:13-18."""
  if seed:
    import random
    import numpy as np
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)


def _batches(dataset, batch_size, device):
  """(x, y) minibatches on `device`, in order: slices of the device tensors of a TensorTask, or -- for any other
  map-style dataset, as the reference's callers pass (var_gp/datasets.py) -- a DataLoader like train_utils.py:22."""
  if hasattr(dataset, 'x') and hasattr(dataset, 'y'):
    xs, ys = dataset.x, dataset.y
    for lo in range(0, len(dataset), batch_size):
      yield xs[lo:lo + batch_size].to(device), ys[lo:lo + batch_size].to(device)
  else:
    from torch.utils.data import DataLoader
    for x, y in DataLoader(dataset, batch_size=batch_size):
      yield x.to(device), y.to(device)


@torch.no_grad()
def compute_accuracy(dataset, gp, batch_size=512, device=None, noise_fn=None):
  """var_gp/train_utils.py:21-35.  `noise_fn(i, B)` may pin the predictive noise of batch i.  Counts stay on the
  device (one host sync per evaluation instead of one per batch); the reference's NaN assertion is kept."""
  device = device or gp.z.device
  correct = torch.zeros((), dtype=torch.int64, device=device)
  nans = torch.zeros((), dtype=torch.bool, device=device)
  for i, (x, y) in enumerate(_batches(dataset, batch_size, device)):
    preds = gp.predict(x, noise=None if noise_fn is None else noise_fn(i, x.size(0)))
    nans |= torch.isnan(preds).any()
    correct += (preds.argmax(dim=-1) == y).sum()
  if bool(nans.item()):
    raise AssertionError('Found NaNs')
  return correct.item() / len(dataset)


@torch.no_grad()
def compute_acc_ent(dataset, gp, batch_size=512, device=None):
  """var_gp/train_utils.py:38-56: mean accuracy and mean predictive entropy."""
  device = device or gp.z.device
  correct = torch.zeros((), dtype=torch.int64, device=device)
  ent = torch.zeros((), device=device)
  nans = torch.zeros((), dtype=torch.bool, device=device)
  for x, y in _batches(dataset, batch_size, device):
    preds = gp.predict(x)
    nans |= torch.isnan(preds).any()
    correct += (preds.argmax(dim=-1) == y).sum()
    ent += -(preds * preds.clamp_min(1e-38).log()).sum()
  if bool(nans.item()):
    raise AssertionError('Found NaNs')
  n = len(dataset)
  return correct.item() / n, ent.item() / n


def compute_bwt(acc_mat):
  """Backward transfer (var_gp/train_utils.py:59-66): mean over earlier tasks of final minus just-trained accuracy."""
  acc_mat = torch.as_tensor(acc_mat)
  if acc_mat.ndim != 2 or acc_mat.shape[0] != acc_mat.shape[1]:
    raise AssertionError('acc_mat must be square')
  return (acc_mat[-1][:-1] - acc_mat.diagonal()[:-1]).mean()


class EarlyStopper:
  """var_gp/train_utils.py:70-98 (same counter / delta semantics)."""

  def __init__(self, patience=10, delta=1e-4):
    self.patience, self.delta = patience, delta
    self._counter, self._best_info, self._best_score = 0, None, None

  def is_done(self):
    return self.patience >= 0 and self._counter >= self.patience

  def info(self):
    return self._best_info

  def __call__(self, score, info):
    if self.is_done():
      raise AssertionError('stopper is done')
    if self._best_score is None:
      self._best_score, self._best_info = score, info
    elif score < self._best_score + self.delta:
      self._counter += 1
    else:
      self._best_score, self._best_info, self._counter = score, info, 0


def snapshot_state(gp):
  """Detached copy of the state dict (the reference keeps aliases of the live parameters, so its "best" checkpoint
  is in fact the latest one -- SURVEY.md section 5; here the best really is the best)."""
  return {k: v.detach().clone() for k, v in gp.state_dict().items()}


def train(task_id, train_set, val_set, test_set, ep_var_mean=True, map_est_hypers=False, dkl=False, epochs=1, M=20,
          n_f=10, n_var_samples=3, batch_size=512, lr=1e-2, beta=1.0, eval_interval=10, patience=20,
          prev_params=None, logger=None, device=None, use_graph=True):
  """experiments/vargp.py:14-73 with the same keyword arguments.  Returns the best state dict (by validation
  accuracy), which the caller appends to `prev_params` for the next task."""
  from .vargp import VARGP
  device = torch.device(device or 'cuda')
  gp = VARGP.create_clf(train_set, M=M, n_f=n_f, n_var_samples=n_var_samples, prev_params=prev_params,
                        ep_var_mean=ep_var_mean, map_est_hypers=map_est_hypers, dkl=dkl).to(device)
  stopper = EarlyStopper(patience=patience)
  N = len(train_set)
  xs, ys = train_set.x.to(device), train_set.y.to(device)
  fused = ep_var_mean and not dkl
  steppers = {}                      # one captured graph per batch size (the last batch of an epoch is short)

  def stepper_for(B):
    if B not in steppers:
      opt = steppers[next(iter(steppers))].opt if steppers else None
      steppers[B] = ElboStepper(gp, n_data=N, batch_size=B, beta=beta, lr=lr, use_graph=use_graph and fused,
                                optimizer=opt)
    return steppers[B]

  terms = None
  for e in range(epochs):
    perm = torch.randperm(N, device=device)                 # DataLoader(shuffle=True)
    for lo in range(0, N, batch_size):
      idx = perm[lo:lo + batch_size]
      terms = stepper_for(idx.numel()).step(xs[idx], ys[idx])
    for st in steppers.values():                            # a non-PD Gram surfaces within the epoch (the reference
      st.check_errors()                                     # raises inside the step, gp_utils.py:10): one sync per epoch
    if (e + 1) % eval_interval == 0:
      acc = {k: compute_accuracy(d, gp, device=device) for k, d in (('train', train_set), ('val', val_set),
                                                                    ('test', test_set))}
      summary = {f'task{task_id}/{k}/acc': v for k, v in acc.items()}
      summary.update({f'task{task_id}/loss/{k}': float(v) for k, v in zip(('kl_hypers', 'kl_u', 'lik'), terms)})
      if logger is not None:
        for k, v in summary.items():
          logger.add_scalar(k, v, global_step=e + 1)
      stopper(acc['val'], dict(state_dict=snapshot_state(gp), acc_summary=summary, step=e + 1))
      if stopper.is_done():
        break
  for st in steppers.values():
    st.check_errors()
  info = stopper.info()
  return info['state_dict'] if info is not None else snapshot_state(gp)
