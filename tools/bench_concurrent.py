"""Throughput of logML+gradient evaluations at size n with C evaluations in flight on one GPU (one host thread and one
CUDA stream per slot): the panel chains / recursion leaves of one evaluation overlap the big GEMMs of another.
    python tools/bench_concurrent.py 20000 6 1,2,3
"""
import math
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from lsqfitgp_b200 import _lib, _ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 6
CS = [int(c) for c in sys.argv[3].split(',')] if len(sys.argv) > 3 else [1, 2, 3]
dev = torch.device('cuda:0')
rng = np.random.default_rng(2002)
X = rng.uniform(0, 10, (n, 3))
y = np.sin(X[:, 0]) + np.cos(X[:, 1]) * X[:, 2] / 10 + 0.1 * rng.standard_normal(n)
xd = torch.tensor(np.ascontiguousarray(X.T)).to(dev)
yd = torch.tensor(y).to(dev)


def descs_for(theta):
    ell, sf, sn = np.exp(theta)
    return [dict(kind=_lib.K_MATERNP, term=0, dimmask=7, ipar=2, par0=0.0, scale_x=ell, scale_y=ell, amp=sf ** 2),
            dict(kind=_lib.K_WHITE, term=1, dimmask=7, amp=sn ** 2)]


def evaluate(theta, K):
    descs = descs_for(theta)
    _ops.gram_iso(descs, xd, xd, out=K, symmetric=True)
    st = _ops.chol_factor(K)
    a = _ops.chol_solve(st, yd[:, None], False)
    ldq = _ops.chol_logdet_quad(st, a[:, 0].contiguous())
    b = _ops.chol_solve(st, a, True, inplace=True)
    Kinv = _ops.chol_inverse(st)
    vjp = _ops.gram_iso_vjp(descs, xd, Kinv, b[:, 0].contiguous())
    ld, q = ldq.cpu().numpy()
    v = vjp.cpu().numpy()
    return 0.5 * (n * math.log(2 * math.pi) + 2 * ld + q), v


thetas = [np.array([math.log(1.5), 0.0, math.log(0.1)]) + 0.01 * i for i in range(S)]
Ks = [_ops.aligned_empty(n, n, dev) for _ in range(max(CS))]
ref = [evaluate(t, Ks[0]) for t in thetas[:2]]  # warm-up (also creates the library's internal streams single-threaded)
torch.cuda.synchronize()
for C in CS:
    out = [None] * S

    def worker(slot):
        stream = torch.cuda.Stream(dev)
        with torch.cuda.stream(stream):
            for i in range(slot, S, C):
                out[i] = evaluate(thetas[i], Ks[slot])
            stream.synchronize()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=worker, args=(c,)) for c in range(C)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ok = all(abs(out[i][0] - ref[i][0]) <= 1e-12 * abs(ref[i][0]) for i in range(2))
    print(f'n={n} in-flight={C}: {S} evaluations in {dt * 1e3:.1f} ms = {S / dt:.3f} evals/s  (values agree: {ok})', flush=True)
