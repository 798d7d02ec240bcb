import math, sys, os, torch
sys.path.insert(0, '/root/repo')
from vargp_b200 import ops as vops
from tests.emu_ops import EmuOps
EMU = EmuOps()
ops = vops.get_ops()
def relerr(a, b):
  a, b = a.double().cpu(), b.double().cpu()
  return ((a - b).norm() / b.norm()).item()
H, C, P, B, D = 3, 4, 512, 1024, 784
g = torch.Generator().manual_seed(1)
theta = 0.1 * torch.randn(H, D + 1, generator=g, dtype=torch.float64) + math.log(math.sqrt(D) / 3)
zs, xs = torch.rand(H, C, P, D, dtype=torch.float64, generator=g) / 8, torch.rand(H, 1, B, D, dtype=torch.float64, generator=g) / 8
zn, xn = (zs * zs).sum(-1), (xs * xs).sum(-1)
K64 = torch.empty(H, C, P, B, dtype=torch.float64)
EMU.rbf_gram(zs, zn, xs, xn, theta, K64, False)
f = lambda t: t.to('cuda', torch.float32)
for mt in (-1, 1):
  ops.tc2_config(mt)
  Kd = torch.empty(H, C, P, B, device='cuda')
  ops.rbf_gram(f(zs), f(zn), f(xs), f(xn), f(theta), Kd, False)
  # plain products
  A = torch.randn(2, 1024, 1024, generator=g, dtype=torch.float64); Bm = torch.randn(2, 1024, 1024, generator=g, dtype=torch.float64)
  Cd = torch.empty(2, 1024, 1024, device='cuda')
  ops.gemm(f(A), f(Bm), Cd)
  Ap = torch.rand(2, 1024, 784, generator=g, dtype=torch.float64); Bp = torch.rand(2, 784, 1024, generator=g, dtype=torch.float64)
  Cp = torch.empty(2, 1024, 1024, device='cuda')
  ops.gemm(f(Ap), f(Bp), Cp)
  # fp32 torch reference error for context
  print('tc2' if mt == 1 else 'tc1', 'rna' if os.environ.get('VARGP_TC2_RNA') == '1' else 'raw', 'rbf', relerr(Kd, K64), 'randn', relerr(Cd, A @ Bm), 'pos', relerr(Cp, Ap @ Bp),
        'pos bias', ((Cp.double().cpu() - Ap @ Bp) / (Ap @ Bp)).mean().item())
torch.backends.cuda.matmul.allow_tf32 = False
print('torch fp32 pos', relerr(f(Ap) @ f(Bp), Ap @ Bp), 'randn', relerr(f(A) @ f(Bm), A @ Bm))
