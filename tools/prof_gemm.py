"""One DMMA GEMM launch (NT, n^3) for ncu."""
import sys
import torch
sys.path.insert(0, '.')
from lsqfitgp_b200 import _lib, _ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device('cuda:0')
A = torch.randn(n, n, dtype=torch.float64, device=dev)
B = torch.randn(n, n, dtype=torch.float64, device=dev)
C = torch.zeros(n, n, dtype=torch.float64, device=dev)
for _ in range(3):
    _ops.dgemm(A, B, C, a_kmajor=True, b_kmajor=True, M=n, N=n, K=n, flags=_lib.GEMM_BETA0)
torch.cuda.synchronize()
print('done')
