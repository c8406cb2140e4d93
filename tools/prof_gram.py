"""Symmetric Gram build at size n (Matern-5/2 + White, 3-D: BASELINE configs[1]) for timing / ncu."""
import sys
import torch
sys.path.insert(0, '.')
from lsqfitgp_b200 import _lib, _ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
kind = sys.argv[2] if len(sys.argv) > 2 else 'matern52'
dev = torch.device('cuda:0')
x = torch.rand(3, n, dtype=torch.float64, device=dev) * 10
if kind == 'matern52':
    descs = [dict(kind=_lib.K_MATERNP, term=0, dimmask=7, ipar=2, par0=0.0, scale_x=1.5, scale_y=1.5, amp=1.0),
             dict(kind=_lib.K_WHITE, term=1, dimmask=7, amp=0.01)]
elif kind == 'ratquad':
    descs = [dict(kind=_lib.K_CAUCHY, term=0, dimmask=7, par0=2.0, par1=3.0, scale_x=1.5, scale_y=1.5, amp=1.0),
             dict(kind=_lib.K_WHITE, term=1, dimmask=7, amp=0.01)]
else:
    descs = [dict(kind=_lib.K_EXPQUAD, term=0, dimmask=7, scale_x=1.5, scale_y=1.5, amp=1.0),
             dict(kind=_lib.K_WHITE, term=1, dimmask=7, amp=0.01)]
K = _ops.aligned_empty(n, n, dev)
for sym in (True, False):
    for _ in range(2):
        _ops.gram_iso(descs, x, x, out=K, symmetric=sym)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); _ops.gram_iso(descs, x, x, out=K, symmetric=sym); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts)
    print(f'{kind} n={n} symmetric={sym}: {t:.3f} ms  {8*n*n/t/1e6:.0f} GB/s written')
