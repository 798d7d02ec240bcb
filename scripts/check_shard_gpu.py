#!/usr/bin/env python
"""Multi-GPU check of elbo.FactorShard under NCCL (torchrun --nproc-per-node N scripts/check_shard_gpu.py):
the same data-parallel ELBO step with and without sharding of the factor stage; prints the largest norm-relative
difference of the all-reduced gradients (fp32: expect ~1e-6) and the step times."""
import os, sys, time, json
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vargp_b200.synthetic import make_case
from vargp_b200.dist import shard_coef
from vargp_b200.elbo import FactorShard

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
# VARGP_CHECK_BACKEND=gloo: all ranks share cuda:0 and the collectives go through gloo (host staged) -- checks the
# kernels on the sharded (h, c) sub-rectangles on a single-GPU box; the default is one GPU per rank over NCCL.
backend = os.environ.get('VARGP_CHECK_BACKEND', 'nccl')
if backend == 'nccl':
  torch.cuda.set_device(local)
  dist.init_process_group('nccl', device_id=torch.device('cuda', local))
else:
  torch.cuda.set_device(0)
  dist.init_process_group(backend)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 640
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
params, prev, x, y, noise = make_case(C=10, D=784, M=P, t=0, B=B * world, sigma=10., seed=5)
sl = slice(rank * B, (rank + 1) * B)
nz = {k: v.cuda() for k, v in dict(noise, eps_f=noise['eps_f'][..., sl].contiguous()).items()}
res = {}
for mode in ('replicated', 'sharded'):
  gp = bench.build_gpu_model(params, prev, torch.device('cuda'))
  gp.sync_errors = False
  gp.factor_shard = FactorShard() if mode == 'sharded' else None
  a, b, c = shard_coef(1.0, 10. * B * world, B * world, world, factor_sharded=mode == 'sharded')
  for it in range(3):
    gp.zero_grad()
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    kl_h, kl_u, nll = gp.loss(x[sl].cuda(), y[sl].cuda(), noise=nz)
    (a * kl_h + b * kl_u + c * nll).backward()
    flat = torch.cat([p.grad.reshape(-1) for p in gp.parameters()])
    dist.all_reduce(flat)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
  gp.check_errors()
  klu = kl_u.detach().clone()
  if mode == 'sharded':
    dist.all_reduce(klu)
  res[mode] = (flat, klu, dt)
if rank == 0:
  f0, k0, t0 = res['replicated']; f1, k1, t1 = res['sharded']
  print(json.dumps(dict(backend=backend, world=world, P=P, B_per_rank=B, grad_relerr=((f0 - f1).norm() / f0.norm()).item(),
                        kl_u_relerr=abs((k0 - k1).item()) / abs(k0.item()), ms_replicated=round(1e3 * t0, 2),
                        ms_sharded=round(1e3 * t1, 2))))
dist.destroy_process_group()
