"""Golden fixtures at the BASELINE configs' FULL sizes, from the LIVE reference in fp64 (build container only).

    python tests/golden/make_large.py [name ...]

The cases are the large-shape dispatch paths of the product (the benched Split-MNIST t=4 step, the Permuted-MNIST
t=9 step that takes the 2-CTA GEMM and an 8-block factorisation, one slice of the scaled synthetic config with
P=2048) -- too slow to re-run through the fp64 reference / oracle on the GPU box at every test run (40 s ... 10 min
here), so the reference's fp64 outputs are stored.  Inputs are re-derived from the seed by `make_case`.

The big gradients (z: C*P*784 values, u_tril_vec: C*M(M+1)/2 values) would be 10..150 MB each, so tensors above
`FULL_MAX` elements are stored *compressed*: their norm, every `stride`-th element, and `NPROJ` projections onto
seeded random-sign vectors (an error anywhere in the tensor moves every projection by ~ its norm).
`tests/util.compare_compressed` is the matching comparator.  As in make_golden.py the oracle restatement is
checked against the reference on the spot (run_reference asserts it).
"""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

LARGE_CASES = {
  # the benched step (bench.py default): Split-MNIST shape at the last task
  'large_split_t4':    dict(C=10, D=784, M=60, t=4, B=512, sigma=10., seed=61),
  # Permuted-MNIST shape at the last task: P = 1000 (gemm_tc2 + multi-block Cholesky dispatch)
  'large_permuted_t9': dict(C=10, D=784, M=100, t=9, B=512, sigma=10., seed=62),
  # one slice of the scaled synthetic config: P = M = 2048 inducing points, a 2048-point shard of the minibatch
  'large_scaled_slice': dict(C=10, D=784, M=2048, t=0, B=2048, sigma=10., seed=63),
}
FULL_MAX = 300_000
SUB_TARGET = 150_000
NPROJ = 8


def signs(n, seed):
  g = torch.Generator().manual_seed(seed)
  return torch.randint(0, 2, (NPROJ, n), generator=g, dtype=torch.int8)


def project(flat, seed):
  """NPROJ random-sign projections of a flat fp64 tensor (chunked: the sign matrix of a 21 M-element tensor is 170 MB)."""
  n = flat.numel()
  s = signs(n, seed)
  out = torch.zeros(NPROJ, dtype=torch.float64)
  step = 1 << 22
  for lo in range(0, n, step):
    out += (s[:, lo:lo + step].double() * 2 - 1) @ flat[lo:lo + step].double()
  return out


def compress(t, seed):
  t = t.detach().double().cpu()
  if t.numel() <= FULL_MAX:
    return t
  flat = t.reshape(-1)
  stride = max(2, flat.numel() // SUB_TARGET)
  return dict(compressed=True, shape=tuple(t.shape), norm=flat.norm().item(), stride=stride,
              sub=flat[::stride].clone(), seed=seed, proj=project(flat, seed))


def main():
  from make_golden import load_reference, run_reference
  names = sys.argv[1:] or list(LARGE_CASES)
  refmods = load_reference()
  for name in names:
    kw = LARGE_CASES[name]
    t0 = time.time()
    out = run_reference(refmods, kw, torch.float64)
    rec = dict(case=kw, torch=torch.__version__, beta=out['beta'], Ntot=out['Ntot'])
    seed = 1000 + kw['seed']
    rec['f64'] = {k: (compress(v, seed) if torch.is_tensor(v) else v) for k, v in out.items() if k != 'grads'}
    rec['f64']['grads'] = {k: compress(v, seed + 1 + i) for i, (k, v) in enumerate(sorted(out['grads'].items()))}
    path = os.path.join(HERE, name + '.pt')
    torch.save(rec, path)
    print(f'{name:22s} kl_u={out["kl_u"].item():.6f} nll={out["nll"].item():.6f} '
          f'{time.time() - t0:.0f} s -> {os.path.getsize(path) / 1024:.0f} KiB', flush=True)


if __name__ == '__main__':
  main()
