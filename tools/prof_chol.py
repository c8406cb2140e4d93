"""Run one factorisation (+solve, inverse) at size n for profiling."""
import sys
import torch
sys.path.insert(0, '.')
from lsqfitgp_b200 import _lib, _ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
what = sys.argv[2] if len(sys.argv) > 2 else 'factor'
dev = torch.device('cuda:0')
x = torch.rand(3, n, dtype=torch.float64, device=dev) * 10
descs = [dict(kind=_lib.K_MATERNP, term=0, dimmask=7, ipar=2, par0=0.0, scale_x=1.5, scale_y=1.5, amp=1.0),
         dict(kind=_lib.K_WHITE, term=1, dimmask=7, amp=0.01)]
K = _ops.gram_iso(descs, x, x)
torch.cuda.synchronize()
st = _ops.chol_factor(K)
torch.cuda.synchronize()
if what == 'all':
    b = torch.randn(n, 1, dtype=torch.float64, device=dev)
    _ops.chol_solve(st, b, False)
    _ops.chol_inverse(st)
    torch.cuda.synchronize()
print('info', int(st.info.item()))
