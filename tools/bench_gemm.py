"""GEMM shapes that matter for the factorisation: big NT, rank-512 SYRK (trailing update), LAUUM-like TN."""
import sys
import torch
sys.path.insert(0, '.')
from lsqfitgp_b200 import _lib, _ops
dev = torch.device('cuda:0')
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)
n = 8192
A = torch.randn(n, n, dtype=torch.float64, device=dev)
B = torch.randn(n, n, dtype=torch.float64, device=dev)
C = torch.zeros(n, n, dtype=torch.float64, device=dev)
t = timeit(lambda: _ops.dgemm(A, B, C, a_kmajor=True, b_kmajor=True, M=n, N=n, K=n, flags=_lib.GEMM_BETA0))
print(f'NT {n}^3: {t:.2f} ms {2*n**3/t/1e9:.2f} TF')
t = timeit(lambda: _ops.dgemm(A, B, C, a_kmajor=False, b_kmajor=False, M=n, N=n, K=n, flags=_lib.GEMM_BETA0))
print(f'TN(mm) {n}^3: {t:.2f} ms {2*n**3/t/1e9:.2f} TF')
m = 16384
P = torch.randn(m, 1024, dtype=torch.float64, device=dev)
Cm = torch.zeros(m, m, dtype=torch.float64, device=dev)
for k in (256, 512, 1024):
    t = timeit(lambda: _ops.dgemm(P, P, Cm, a_kmajor=True, b_kmajor=True, M=m, N=m, K=k, alpha=-1.0, flags=_lib.GEMM_LOWER))
    print(f'SYRK lower m={m} k={k}: {t:.2f} ms {m*m*k/t/1e9:.2f} TF')
t = timeit(lambda: _ops.dgemm(A, A, C, a_kmajor=False, b_kmajor=False, M=n, N=n, K=n, flags=_lib.GEMM_BETA0 | _lib.GEMM_LOWER | _lib.GEMM_A_UPPER_K))
print(f'LAUUM-like n={n}: {t:.2f} ms {n**3/3/t/1e9:.2f} TF')
