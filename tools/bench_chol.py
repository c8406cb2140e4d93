"""Time chol_factor at size n for several panel widths; compare with cuSOLVER."""
import os, sys
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
sizes = [int(a) for a in sys.argv[1].split(',')] if len(sys.argv) > 1 else [20000]
pbs = [int(a) for a in sys.argv[2].split(',')] if len(sys.argv) > 2 else [2, 4, 8]
for nn in sizes:
    x = torch.rand(3, nn, dtype=torch.float64, device=dev) * 10
    descs = [dict(kind=_lib.K_MATERNP, term=0, dimmask=7, ipar=2, par0=0.0, scale_x=1.5, scale_y=1.5, amp=1.0),
             dict(kind=_lib.K_WHITE, term=1, dimmask=7, amp=0.01)]
    K = _ops.gram_iso(descs, x, x)
    ref = None
    for pb in pbs:
        os.environ['LGP_PANEL_BLOCKS'] = str(pb)
        hold = {}
        def fac():
            hold['st'] = _ops.chol_factor(K)
        t = timeit(fac)
        st = hold['st']
        ld = float(st.scalars()[4].item())
        if ref is None:
            Lt = torch.linalg.cholesky(K)
            ref = float(torch.log(torch.diagonal(Lt)).sum().item())
            del Lt
        print(f'n={nn} pb={pb}: {t:.2f} ms  {nn**3/3/t/1e9:.2f} TFLOP/s info={int(st.info.item())} logdet relerr={(ld-ref)/abs(ref):.2e}', flush=True)
        del st, hold
    t2 = timeit(lambda: torch.linalg.cholesky(K))
    print(f'n={nn} cusolver: {t2:.2f} ms  {nn**3/3/t2/1e9:.2f} TFLOP/s', flush=True)
    del K
