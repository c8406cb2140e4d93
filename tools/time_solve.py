"""Time the vector triangular solves (forward + backward sweep, m = 1) behind lgp_chol_solve at the sizes given, and
check them against the residual of the factorised matrix."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from lsqfitgp_b200 import _lib, _ops
dev = torch.device('cuda:0')
sizes = [int(a) for a in sys.argv[1].split(',')] if len(sys.argv) > 1 else [20000]
for n in sizes:
    x = torch.rand(3, n, dtype=torch.float64, device=dev) * 10
    descs = [dict(kind=_lib.K_MATERNP, term=0, dimmask=7, ipar=2, par0=0.0, scale_x=1.5, scale_y=1.5, amp=1.0),
             dict(kind=_lib.K_WHITE, term=1, dimmask=7, amp=0.01)]
    K = _ops.gram_iso(descs, x, x, symmetric=True)
    st = _ops.chol_factor(K)
    y = torch.randn(n, 1, dtype=torch.float64, device=dev)
    def pair():
        a = _ops.chol_solve(st, y, False)
        return _ops.chol_solve(st, a, True, inplace=True)
    sol = pair(); torch.cuda.synchronize()
    eps = float(st.scalars()[1].item())
    res = K @ sol + eps * sol - y  # (equilibration scales are powers of two of a unit-diagonal-ish kernel: S = 1 here)
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pair(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f'n={n}: solve pair {np.mean(ts):.3f} ms (min {min(ts):.3f}); 8 n^2 B / t = {8*n*n/np.mean(ts)/1e9:.2f} TB/s; residual {float(res.norm()/y.norm()):.2e}')
    del K, st
