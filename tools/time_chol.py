"""Time lgp_chol_factor (mean of `reps` calls after a warm-up) at the sizes given; env switches are read once per process."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from lsqfitgp_b200 import _lib, _ops
dev = torch.device('cuda:0')
sizes = [int(a) for a in sys.argv[1].split(',')] if len(sys.argv) > 1 else [20000]
out = []
for n in sizes:
    x = torch.rand(3, n, dtype=torch.float64, device=dev) * 10
    descs = [dict(kind=_lib.K_MATERNP, term=0, dimmask=7, ipar=2, par0=0.0, scale_x=1.5, scale_y=1.5, amp=1.0),
             dict(kind=_lib.K_WHITE, term=1, dimmask=7, amp=0.01)]
    K = _ops.gram_iso(descs, x, x, symmetric=True)
    st = _ops.chol_factor(K)
    torch.cuda.synchronize()
    ld = float(st.scalars()[4].item())
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); st = _ops.chol_factor(K); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    out.append(f'n={n}: {np.mean(ts):.3f} ms ({n**3/3/np.mean(ts)/1e9:.2f} TF) logdet {ld:.10g} info {int(st.info.item())}')
    del st, K
print(' | '.join(out))
