"""Data-parallel ELBO with world_size = 2 on the gloo backend (CPU, emulated kernels):
sharding the minibatch term, replicating Kzz/Cholesky/KL and summing the flat gradient bucket must reproduce
the single-process full-batch gradient and loss terms."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import util

CASE = dict(C=4, D=16, M=6, t=2, B=24, sigma=2., seed=41, H=2, F=3)


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _setup():
  from tests.emu_ops import EmuOps
  from vargp_b200 import ops
  ops.set_ops(EmuOps())


def _worker(rank, world, port, flat, out, shard=False):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  torch.set_num_threads(1)
  _setup()
  from vargp_b200.dist import GradBucket, shard_loss, shard_coef
  from vargp_b200.optim import FlatYogi
  params, prev, x, y, noise, n_v, F, flags = util.case_tensors(CASE, torch.float64)
  gp = util.build_model(params, prev, n_v, F, flags, 'cpu', torch.float64)
  B = x.size(0)
  sl = slice(rank * B // world, (rank + 1) * B // world)
  nz = dict(noise, eps_f=noise['eps_f'][..., sl].contiguous())
  beta, N = 1.7, 240.
  if flat:
    opt = FlatYogi(gp.parameters(), lr=1e-2)
    opt.zero_grad()
  if shard:                                     # the O(P^3) factor stage is dealt out over the ranks as well
    from vargp_b200.elbo import FactorShard
    gp.factor_shard = FactorShard()
  kl_h, kl_u, nll = gp.loss(x[sl], y[sl], noise=nz)
  if shard:
    a, b, c = shard_coef(beta, N, B, world, factor_sharded=True)
    loss = a * kl_h + b * kl_u + c * nll
    kl_u = kl_u.detach().clone()
    dist.all_reduce(kl_u)                       # each rank holds its share
  else:
    loss = shard_loss(kl_h, kl_u, nll, beta, N, B, world)
  loss.backward()
  if flat:
    dist.all_reduce(opt.flat_g, op=dist.ReduceOp.SUM)      # the flat gradient buffer IS the bucket
  else:
    GradBucket(gp.parameters()).allreduce()
  nll_sum = nll.detach().clone()
  dist.all_reduce(nll_sum)
  if rank == 0:
    out.put({k: v.clone() for k, v in util.model_grads(gp).items()} | dict(nll=nll_sum, kl_u=kl_u.detach().clone()))
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.parametrize('flat', [False, True])
def test_two_rank_gradient_equals_single_process(flat):
  _setup()
  params, prev, x, y, noise, n_v, F, flags = util.case_tensors(CASE, torch.float64)
  gp = util.build_model(params, prev, n_v, F, flags, 'cpu', torch.float64)
  terms, grads = util.run_model(gp, x, y, noise, 1.7, 240.)
  ctx = mp.get_context('spawn')
  out = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, flat, out)) for r in range(2)]
  for p in procs:
    p.start()
  got = out.get(timeout=180)
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert util.relerr(got['nll'], terms['nll']) < 1e-12
  assert util.relerr(got['kl_u'], terms['kl_u']) < 1e-12
  for k in util.GRAD_KEYS:
    assert util.relerr(got[k], grads[k]) < 1e-10, k
  from vargp_b200 import ops
  ops.set_ops(None)


@pytest.mark.parametrize('world', [2, 3])
def test_factor_sharding_reproduces_single_process(world):
  """elbo.FactorShard: every rank factors only its (h, c) pairs (H*C = 8 pairs: 4+4 at world 2, 3+3+2 at world 3,
  rectangles crossing the hyper-sample boundary), W / N / nu are all-gathered, the minibatch sums reduce-scattered;
  the summed gradients and the summed kl_u shares must equal the single-process result."""
  _setup()
  params, prev, x, y, noise, n_v, F, flags = util.case_tensors(CASE, torch.float64)
  gp = util.build_model(params, prev, n_v, F, flags, 'cpu', torch.float64)
  terms, grads = util.run_model(gp, x, y, noise, 1.7, 240.)
  ctx = mp.get_context('spawn')
  out = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, world, port, True, out, True)) for r in range(world)]
  for p in procs:
    p.start()
  got = out.get(timeout=180)
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert util.relerr(got['nll'], terms['nll']) < 1e-12
  assert util.relerr(got['kl_u'], terms['kl_u']) < 1e-12
  for k in util.GRAD_KEYS:
    assert util.relerr(got[k], grads[k]) < 1e-10, k
  from vargp_b200 import ops
  ops.set_ops(None)
