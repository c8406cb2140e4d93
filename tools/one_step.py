"""One logML+gradient evaluation at n (default 20000) straight through the C ABI, as bench.py's device arm does it: the
command profiled for the per-round launch list (profiles/launches_step_rNN.csv)."""
import math
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from lsqfitgp_b200 import _lib, _ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device('cuda:0')
rng = np.random.default_rng(2002)
X = rng.uniform(0, 10, (n, 3))
y = np.sin(X[:, 0]) + np.cos(X[:, 1]) * X[:, 2] / 10 + 0.1 * rng.standard_normal(n)
xd = torch.tensor(np.ascontiguousarray(X.T)).to(dev)
yd = torch.tensor(y).to(dev)
descs = [dict(kind=_lib.K_MATERNP, term=0, dimmask=7, ipar=2, par0=0.0, scale_x=1.5, scale_y=1.5, amp=1.0),
         dict(kind=_lib.K_WHITE, term=1, dimmask=7, amp=0.01)]
K = _ops.aligned_empty(n, n, dev)
side = torch.cuda.Stream(dev)
for _ in range(reps):
    main = torch.cuda.current_stream()
    st, Kinv = _ops.gram_chol_factor(descs, xd, side=side)   # Gram fused with the equilibration pass (no K)
    a = _ops.chol_solve(st, yd[:, None], False)
    ldq = _ops.chol_logdet_quad(st, a[:, 0].contiguous())
    b = _ops.chol_solve(st, a, True, inplace=True)
    main.wait_stream(side)
    vjp = _ops.gram_iso_vjp(descs, xd, Kinv, b[:, 0].contiguous())
    ld, q = ldq.cpu().numpy()
    print('neg logML', 0.5 * (n * math.log(2 * math.pi) + 2 * ld + q), 'vjp', vjp.cpu().numpy().ravel()[:4], 'info',
          int(st.info.item()))
