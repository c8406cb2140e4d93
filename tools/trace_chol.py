"""Per-panel timeline of one factorisation (LGP_TRACE=1): warm up at a small size first, then the size of interest."""
import os
import sys
os.environ['LGP_TRACE'] = '1'
import torch
sys.path.insert(0, '.')
from lsqfitgp_b200 import _lib, _ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
dev = torch.device('cuda:0')
descs = [dict(kind=_lib.K_MATERNP, term=0, dimmask=7, ipar=2, par0=0.0, scale_x=1.5, scale_y=1.5, amp=1.0),
         dict(kind=_lib.K_WHITE, term=1, dimmask=7, amp=0.01)]
for nn in (4096, n, n):
    x = torch.rand(3, nn, dtype=torch.float64, device=dev) * 10
    K = _ops.gram_iso(descs, x, x)
    torch.cuda.synchronize()
    print(f'==== n = {nn}', file=sys.stderr, flush=True)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    st = _ops.chol_factor(K)
    e1.record()
    torch.cuda.synchronize()
    print(f'==== total {e0.elapsed_time(e1):.3f} ms (with trace synchronisation)', file=sys.stderr, flush=True)
    del st, K
