"""Diagnostic for the MN-major tcgen05 operand path: C = I * B with index-coded B shows which element the
tensor core fetched for every logical (k, n)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vargp_b200 import ops
o = ops.get_ops()
torch.set_printoptions(linewidth=200, sci_mode=False)
M = N = K = 128
A = torch.eye(M, K, device='cuda')
B = (torch.arange(K, device='cuda').float()[:, None] * 1000 + torch.arange(N, device='cuda').float()[None, :]).contiguous()
C = torch.full((M, N), float('nan'), device='cuda')
n0 = o.tc_calls
o.gemm(A, B, C)          # A K-major, B N-major (MN-major)
print('tc used', o.tc_calls - n0, 'NN: expect C == B')
print(C[:10, :10]); print(C[:4, 28:40]); print(C[30:36, :6]); print('max abs diff', (C - B).abs().max().item(), 'nonzero frac', (C != 0).float().mean().item())
# A M-major (A^T stored), B K-major
At = torch.eye(K, M, device='cuda')
Bk = B.t().contiguous()   # (N, K): K-major
C2 = torch.full((M, N), float('nan'), device='cuda')
o.gemm(At.t(), Bk.t(), C2)
print('TN(A mn, B k): expect C == B'); print(C2[:6, :8]); print('max abs diff', (C2 - B).abs().max().item())
# A coded, B identity (A M-major)
Ac = (torch.arange(M, device='cuda').float()[:, None] * 1000 + torch.arange(K, device='cuda').float()[None, :])
Act = Ac.t().contiguous()            # stored (K, M): m contiguous
C3 = torch.full((M, N), float('nan'), device='cuda')
o.gemm(Act.t(), torch.eye(K, N, device='cuda').t().contiguous().t(), C3)
print('A coded M-major x I(K-major): expect C == Ac'); print(C3[:6, :8]); print(C3[30:36, :6]); print('max abs diff', (C3 - Ac).abs().max().item())
