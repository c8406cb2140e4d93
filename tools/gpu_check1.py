"""First on-GPU check: GEMM variants, Cholesky pipeline, Gram kernels vs numpy/torch fp64; timings."""
import sys, time, json
import numpy as np
import torch
sys.path.insert(0, '.')
from lsqfitgp_b200 import _lib, _ops

dev = torch.device('cuda:0')
torch.manual_seed(0)
rng = np.random.default_rng(0)
ok = True

def report(name, err, tol):
    global ok
    good = bool(err <= tol)
    ok &= good
    print(f'{"PASS" if good else "FAIL"} {name}: err={err:.3e} tol={tol:.1e}', flush=True)

def relerr(a, b):
    return float((a - b).abs().max() / b.abs().max())

# ---- GEMM
for (M, N, K) in [(128, 128, 128), (300, 200, 150), (257, 129, 77), (1024, 512, 2048)]:
    for akm in (True, False):
        for bkm in (True, False):
            Aop = torch.randn(M, K, dtype=torch.float64, device=dev)
            Bop = torch.randn(N, K, dtype=torch.float64, device=dev)
            A = _ops.as_aligned(Aop if akm else Aop.T.contiguous())
            B = _ops.as_aligned(Bop if bkm else Bop.T.contiguous())
            C0 = torch.randn(M, N, dtype=torch.float64, device=dev)
            C = _ops.as_aligned(C0.clone())
            _ops.dgemm(A, B, C, a_kmajor=akm, b_kmajor=bkm, M=M, N=N, K=K, alpha=-1.0)
            ref = C0 - Aop @ Bop.T
            report(f'gemm {M}x{N}x{K} akm={akm} bkm={bkm}', relerr(C, ref), 1e-13)
# lower flag
M = 384; K = 200
Aop = torch.randn(M, K, dtype=torch.float64, device=dev)
A = _ops.as_aligned(Aop)
C0 = torch.randn(M, M, dtype=torch.float64, device=dev)
C = _ops.as_aligned(C0.clone())
_ops.dgemm(A, A, C, a_kmajor=True, b_kmajor=True, M=M, N=M, K=K, alpha=-1.0, flags=_lib.GEMM_LOWER)
ref = C0 - Aop @ Aop.T
report('syrk lower (lower part)', relerr(torch.tril(C), torch.tril(ref)), 1e-13)
report('syrk lower (upper untouched)', float((torch.triu(C, 1) - torch.triu(C0, 1)).abs().max()), 0.0)
# triangular-k flags
L = torch.tril(torch.randn(M, M, dtype=torch.float64, device=dev))
X = torch.randn(M, 130, dtype=torch.float64, device=dev)
Xa = _ops.as_aligned(X)
Y = _ops.aligned_empty(M, 130, dev)
_ops.dgemm(_ops.as_aligned(L), Xa, Y, a_kmajor=True, b_kmajor=False, M=M, N=130, K=M, flags=_lib.GEMM_BETA0 | _lib.GEMM_A_LOWER_K)
report('trmm A_LOWER_K', relerr(Y, L @ X), 1e-13)
_ops.dgemm(_ops.as_aligned(L), Xa, Y, a_kmajor=False, b_kmajor=False, M=M, N=130, K=M, flags=_lib.GEMM_BETA0 | _lib.GEMM_A_UPPER_K)
report('trmm A_UPPER_K (L^T X)', relerr(Y, L.T @ X), 1e-13)

# ---- Cholesky
def spd(n, seed):
    g = torch.Generator(device='cpu').manual_seed(seed)
    x = torch.rand(n, 2, generator=g, dtype=torch.float64) * 10
    d2 = ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)
    return (1.7 * torch.exp(-0.5 * d2 / 1.5 ** 2) + 0.05 * torch.eye(n, dtype=torch.float64))

for n in [1, 2, 10, 127, 128, 129, 300, 1000, 2500]:
    Kc = spd(n, n)
    K = Kc.to(dev)
    st = _ops.chol_factor(K)
    info = int(st.info.item())
    Lg = _ops.chol_get_factor(st).cpu().numpy()
    # oracle restatement (A.2)
    Kn = Kc.numpy()
    d = np.diag(Kn)
    s = np.where(d != 0, np.exp2(np.rint(0.5 * np.log2(d))), 1)
    Kt = Kn / s / s[:, None]
    eps = n * np.finfo(float).eps * np.max(np.sum(np.abs(Kt), axis=1))
    Kt[np.diag_indices(n)] += eps
    import scipy.linalg as sl
    Lr = sl.cholesky(Kt, lower=True) * s[:, None]
    report(f'chol n={n} info={info} factor', float(np.abs(Lg - Lr).max() / np.abs(Lr).max()), 1e-12)
    sc = st.scalars().cpu().numpy()
    report(f'chol n={n} eps', abs(sc[1] - eps) / eps, 1e-14)
    report(f'chol n={n} logdet', abs(sc[4] - np.sum(np.log(np.diag(Lr)))) / max(1, abs(np.sum(np.log(np.diag(Lr))))), 1e-12)
    for m in (1, 3, 130):
        B = torch.randn(n, m, dtype=torch.float64)
        x1 = _ops.chol_solve(st, B.to(dev), False).cpu().numpy()
        r1 = sl.solve_triangular(Lr, B.numpy(), lower=True)
        report(f'chol n={n} solve lower m={m}', float(np.abs(x1 - r1).max() / np.abs(r1).max()), 1e-10)
        x2 = _ops.chol_solve(st, B.to(dev), True).cpu().numpy()
        r2 = sl.solve_triangular(Lr.T, B.numpy(), lower=False)
        report(f'chol n={n} solve upper m={m}', float(np.abs(x2 - r2).max() / np.abs(r2).max()), 1e-10)
        y1 = _ops.chol_mult(st, B.to(dev), False).cpu().numpy()
        report(f'chol n={n} mult m={m}', float(np.abs(y1 - Lr @ B.numpy()).max() / np.abs(Lr @ B.numpy()).max()), 1e-12)
        y2 = _ops.chol_mult(st, B.to(dev), True).cpu().numpy()
        report(f'chol n={n} multT m={m}', float(np.abs(y2 - Lr.T @ B.numpy()).max() / np.abs(Lr.T @ B.numpy()).max()), 1e-12)
    Ki = _ops.chol_inverse(st).cpu().numpy()
    Kir = np.linalg.inv(Lr @ Lr.T)
    report(f'chol n={n} inverse(lower)', float(np.abs(np.tril(Ki) - np.tril(Kir)).max() / np.abs(Kir).max()), 1e-9)

# not positive definite -> info
Kbad = torch.eye(200, dtype=torch.float64, device=dev); Kbad[150, 150] = -1.0
st = _ops.chol_factor(Kbad, epsrel=0.0)
print('info for non-PD (expect 151):', int(st.info.item()))
ok &= int(st.info.item()) == 151

# ---- Gram
n, m, d = 333, 257, 3
x = torch.rand(d, n, dtype=torch.float64) * 10
y = torch.rand(d, m, dtype=torch.float64) * 10
def np_r2(x, y, scale, loc=0.0):
    u = (x.numpy() - loc) / scale; v = (y.numpy() - loc) / scale
    r2 = None
    for f in range(u.shape[0]):
        t = (u[f][:, None] - v[f][None, :]) ** 2
        r2 = t if r2 is None else r2 + t
    return r2
full = (1 << d) - 1
descs = [dict(kind=_lib.K_EXPQUAD, term=0, dimmask=full, scale_x=1.5, scale_y=1.5, amp=1.0)]
Kg = _ops.gram_iso(descs, x.to(dev), y.to(dev)).cpu().numpy()
Kr = np.exp(-0.5 * np_r2(x, y, 1.5))
report('gram expquad', float(np.max(np.abs(Kg - Kr) / np.abs(Kr))), 1e-13)
descs = [dict(kind=_lib.K_MATERNP, term=0, dimmask=full, ipar=2, par0=1e-30, scale_x=2.0, scale_y=2.0, amp=1.3),
         dict(kind=_lib.K_WHITE, term=1, dimmask=full, amp=0.01)]
Kg = _ops.gram_iso(descs, x.to(dev), x.to(dev)).cpu().numpy()
r2 = np_r2(x, x, 2.0)
xx = np.sqrt(5 * r2 + 1e-30)
poly = 1
p = 2
for k in reversed(range(p)):
    poly = 1 + poly * ((p - k) / ((2 * p - k) * (k + 1))) * 2 * xx
Kr = 1.3 * (np.exp(-xx) * poly) + 0.01 * np.eye(n)
report('gram maternp2+white', float(np.max(np.abs(Kg - Kr) / np.abs(Kr))), 1e-13)
descs = [dict(kind=_lib.K_CAUCHY, term=0, dimmask=0b101, par0=2.0, par1=3.0, scale_x=1.0, scale_y=1.0, amp=1.0),
         dict(kind=_lib.K_EXPQUAD, term=0, dimmask=0b010, scale_x=0.7, scale_y=0.7, amp=2.0)]
Kg = _ops.gram_iso(descs, x.to(dev), y.to(dev)).cpu().numpy()
ra = np_r2(x[[0, 2]], y[[0, 2]], 1.0); rb = np_r2(x[[1]], y[[1]], 0.7)
Kr = (1 + ra / 3.0) ** (-3.0 / 2.0) * (2.0 * np.exp(-0.5 * rb))
report('gram cauchy*expquad(dim subsets)', float(np.max(np.abs(Kg - Kr) / np.abs(Kr))), 1e-13)

# ---- timings
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)

res = {}
for nn in (4096, 8192, 16384):
    A = torch.randn(nn, nn, dtype=torch.float64, device=dev)
    B = torch.randn(nn, nn, dtype=torch.float64, device=dev)
    C = torch.zeros(nn, nn, dtype=torch.float64, device=dev)
    t = timeit(lambda: _ops.dgemm(A, B, C, a_kmajor=True, b_kmajor=True, M=nn, N=nn, K=nn, flags=_lib.GEMM_BETA0))
    print(f'lgp_dgemm NT n={nn}: {t:.2f} ms  {2*nn**3/t/1e9:.2f} TFLOP/s', flush=True)
    res[f'dgemm_nt_{nn}'] = 2*nn**3/t/1e9
    if nn == 8192:
        for akm, bkm in ((True, False), (False, True), (False, False)):
            t = timeit(lambda: _ops.dgemm(A, B, C, a_kmajor=akm, b_kmajor=bkm, M=nn, N=nn, K=nn, flags=_lib.GEMM_BETA0))
            print(f'lgp_dgemm akm={akm} bkm={bkm} n={nn}: {t:.2f} ms  {2*nn**3/t/1e9:.2f} TFLOP/s', flush=True)
        t = timeit(lambda: _ops.dgemm(A, A, C, a_kmajor=True, b_kmajor=True, M=nn, N=nn, K=nn, alpha=-1.0, flags=_lib.GEMM_LOWER))
        print(f'lgp syrk lower n={nn}: {t:.2f} ms  {nn**3/t/1e9:.2f} TFLOP/s', flush=True)
    t = timeit(lambda: torch.matmul(A, B.T, out=C))
    print(f'cuBLAS dgemm NT n={nn}: {t:.2f} ms  {2*nn**3/t/1e9:.2f} TFLOP/s', flush=True)
    res[f'cublas_nt_{nn}'] = 2*nn**3/t/1e9
    del A, B, C

for nn in (4096, 10000, 20000):
    x = torch.rand(3, nn, dtype=torch.float64, device=dev) * 10
    descs = [dict(kind=_lib.K_MATERNP, term=0, dimmask=7, ipar=2, par0=0.0, scale_x=1.5, scale_y=1.5, amp=1.0),
             dict(kind=_lib.K_WHITE, term=1, dimmask=7, amp=0.01)]
    K = _ops.aligned_empty(nn, nn, dev)
    t = timeit(lambda: _ops.gram_iso(descs, x, x, out=K))
    print(f'gram matern52+white n={nn}: {t:.3f} ms  {8*nn*nn/t/1e6:.1f} GB/s', flush=True)
    d1 = [dict(kind=_lib.K_EXPQUAD, term=0, dimmask=7, scale_x=1.5, scale_y=1.5, amp=1.0)]
    t = timeit(lambda: _ops.gram_iso(d1, x, x, out=K))
    print(f'gram expquad n={nn}: {t:.3f} ms  {8*nn*nn/t/1e6:.1f} GB/s', flush=True)
    _ops.gram_iso(descs, x, x, out=K)
    hold = {}
    def fac():
        hold['st'] = _ops.chol_factor(K)
    t = timeit(fac)
    st = hold['st']
    print(f'lgp chol_factor n={nn}: {t:.2f} ms  {nn**3/3/t/1e9:.2f} TFLOP/s  info={int(st.info.item())}', flush=True)
    res[f'chol_{nn}'] = nn**3/3/t/1e9
    t2 = timeit(lambda: torch.linalg.cholesky(K))
    print(f'cusolver potrf n={nn}: {t2:.2f} ms  {nn**3/3/t2/1e9:.2f} TFLOP/s', flush=True)
    res[f'cusolver_{nn}'] = nn**3/3/t2/1e9
    b = torch.randn(nn, 1, dtype=torch.float64, device=dev)
    t = timeit(lambda: _ops.chol_solve(st, b, False))
    print(f'lgp trsv n={nn}: {t:.2f} ms', flush=True)
    t = timeit(lambda: _ops.chol_inverse(st))
    print(f'lgp chol_inverse n={nn}: {t:.2f} ms  {2*nn**3/3/t/1e9:.2f} TFLOP/s', flush=True)
    # residual check at size
    Lg = _ops.chol_get_factor(st)
    v = torch.randn(nn, 4, dtype=torch.float64, device=dev)
    r = Lg @ (Lg.T @ v) - K @ v
    print(f'   residual |LL^Tv-Kv|/|Kv| = {float(r.norm()/ (K@v).norm()):.2e}', flush=True)
    del K, st, Lg, hold

print(json.dumps(res))
print('ALL OK' if ok else 'SOME FAILED')
sys.exit(0 if ok else 1)
