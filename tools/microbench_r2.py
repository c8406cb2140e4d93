"""Kernel-level timings (CUDA events around `reps` back-to-back launches, after warm-up): Gram (symmetric fast path),
Gram-VJP, BART Gram / derivative / VJP, equilibration pass.  Usage: python tools/microbench_r2.py [n]"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from lsqfitgp_b200 import _lib, _ops  # noqa: E402
import lsqfitgp_b200 as lgp  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
dev = torch.device('cuda:0')
out = {}


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


rng = np.random.default_rng(2002)
x = torch.tensor(rng.uniform(0, 10, (3, n)), device=dev)
K = _ops.aligned_empty(n, n, dev)
for name, descs in [
    ('matern52_white', [dict(kind=_lib.K_MATERNP, term=0, dimmask=7, ipar=2, par0=0.0, scale_x=1.5, scale_y=1.5, amp=1.0),
                        dict(kind=_lib.K_WHITE, term=1, dimmask=7, amp=0.01)]),
    ('expquad_white', [dict(kind=_lib.K_EXPQUAD, term=0, dimmask=7, scale_x=1.5, scale_y=1.5, amp=1.0),
                       dict(kind=_lib.K_WHITE, term=1, dimmask=7, amp=0.01)]),
]:
    ms = timeit(lambda: _ops.gram_iso(descs, x, x, out=K, symmetric=True))
    out[f'gram_{name}_ms'] = ms
    out[f'gram_{name}_GBps'] = 8 * n * n / ms / 1e6
    b = torch.tensor(rng.standard_normal(n), device=dev)
    ms = timeit(lambda: _ops.gram_iso_vjp(descs, x, K, b))
    out[f'vjp_{name}_ms'] = ms
    out[f'vjp_{name}_GBps'] = 4 * n * n / ms / 1e6
st = _ops.chol_factor(K)
del st
# BART, BASELINE configs[3]
n4 = 5000
X4 = np.concatenate([rng.standard_normal((n4, 8)), rng.integers(0, 2, (n4, 2)).astype(float)], axis=1)
splits = lgp.BART.splits_from_coord(X4)
idx = lgp.BART.indices_from_coord(X4, splits)
ix = torch.tensor(np.ascontiguousarray(idx.T.astype(np.int32)), device=dev)
spec = lgp.BART(splits=splits, indices=True, alpha=0.95, beta=2, maxd=10, reset=[2, 4, 6, 8])._bart[0]
widths, nrows, rows, drows, gamma = spec.stages()
w = np.ones(10)
Kb = _ops.aligned_empty(n4, n4, dev)
out['bart_gram_sym_ms'] = timeit(lambda: _ops.gram_bart_stages(splits[0], w, widths, nrows, rows, drows, gamma, 1.0, ix, ix,
                                                               out=Kb, symmetric=True))
iy = ix.clone()
out['bart_gram_full_ms'] = timeit(lambda: _ops.gram_bart_stages(splits[0], w, widths, nrows, rows, drows, gamma, 1.0, ix, iy,
                                                                out=Kb, symmetric=False))
out['bart_deriv_sym_ms'] = timeit(lambda: _ops.gram_bart_stages(splits[0], w, widths, nrows, rows, drows, gamma, 1.0, ix, ix,
                                                                deriv=True, symmetric=True), reps=5)
G = torch.randn(n4, n4, dtype=torch.float64, device=dev)
G = _ops.as_aligned(G + G.T)
bb = torch.randn(n4, dtype=torch.float64, device=dev)
out['bart_vjp_symlower_ms'] = timeit(lambda: _ops.gram_bart_vjp(splits[0], w, widths, nrows, rows, drows, gamma, 1.0, ix, ix,
                                                                G, b=bb, symlower=True))
out['bart_pairs_per_s_sym'] = n4 * n4 / out['bart_gram_sym_ms'] / 1e-3
print(json.dumps(out))
