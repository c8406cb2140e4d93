"""Device timings of the BASELINE.json configurations other than the headline one, through the public API:
C1 (ExpQuad n=1000: marginal_likelihood + predfromdata), C3 (ExpQuad+noise n=10000: logML+gradient, 1 and 4 in flight),
C4 (BART n=5000 p=10: Gram build, Cholesky, recipe logML).  Prints one JSON object."""
import json
import math
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
import lsqfitgp_b200 as lgp  # noqa: E402
from lsqfitgp_b200 import _dist, _ops  # noqa: E402

dev = torch.device('cuda:0')
out = {}


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3


# ---- C1
rng = np.random.default_rng(1001)
x = np.sort(rng.uniform(0, 100, 1000))
y = np.sin(x / 3) + 0.1 * rng.standard_normal(1000)
xp = np.linspace(-5, 105, 500)


ycov1 = {('d', 'd'): 0.01 * np.eye(1000)}


def c1():
    gp = lgp.GP(lgp.ExpQuad(scale=3), checkpos=False, checksym=False).addx(x, 'd').addx(xp, 'p')
    ml = gp.marginal_likelihood({'d': y}, ycov1)
    m, c = gp.predfromdata({'d': y}, 'p', ycov1, raw=True)
    return ml
out['C1_n1000_ml_plus_pred_ms'] = timed(c1)

# ---- C3
rng = np.random.default_rng(3003)
n3 = 10000
X3 = rng.uniform(0, 100, (n3, 2))
y3 = np.sin(X3[:, 0] / 5) + np.cos(X3[:, 1] / 7) + 0.1 * rng.standard_normal(n3)
x3 = lgp.unstructured_to_structured(X3, names=['a', 'b'])


def c3(theta):
    th = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
    k = torch.exp(th[1]) ** 2 * lgp.ExpQuad(scale=torch.exp(th[0])) + torch.exp(th[2]) ** 2 * lgp.White()
    gp = lgp.GP(k, checkpos=False, checksym=False, checkfinite=False).addx(x3, 'data')
    ml = gp.marginal_likelihood({'data': y3})
    g, = torch.autograd.grad(ml, th)
    return np.r_[float(ml.detach()), g.numpy()]
thetas = np.array([np.log(3), 0.0, np.log(0.1)]) + 0.5 * np.random.default_rng(3004).standard_normal((16, 3))
for c in (1, 4):
    ms = timed(lambda: _dist.eval_batch_sharded(c3, thetas, device=dev, in_flight=c), reps=2)
    out[f'C3_n10000_logml_grad_evals_per_s_inflight{c}'] = len(thetas) / (ms * 1e-3)

# ---- C4
rng = np.random.default_rng(4004)
n4 = 5000
X4 = np.concatenate([rng.standard_normal((n4, 8)), rng.integers(0, 2, (n4, 2)).astype(float)], axis=1)
splits = lgp.BART.splits_from_coord(X4)
idx = lgp.BART.indices_from_coord(X4, splits)
xi = lgp.unstructured_to_structured(idx.astype(np.int32), names=[f'c{i}' for i in range(10)])
kb = lgp.BART(splits=splits, indices=True, alpha=0.95, beta=2, maxd=10, reset=[2, 4, 6, 8], gamma=1)
gp4 = lgp.GP(kb, checkpos=False, checksym=False, checkfinite=False).addx(xi, 'train')
elem = gp4._elements['train']


def bart_gram():
    return kb._gram_device(elem.xd, elem.xd, elem.labels)
ms = timed(bart_gram)
out['C4_bart_gram_n5000_p10_ms'] = ms
out['C4_bart_gram_Gpairs_per_s'] = n4 * n4 / (ms * 1e-3) / 1e9
Kb = bart_gram() + 0.25 * torch.eye(n4, dtype=torch.float64, device=dev)
out['C4_chol_n5000_ms'] = timed(lambda: _ops.chol_factor(Kb))
y4 = rng.standard_normal(n4)


noise4 = torch.diag(torch.full((n4,), 0.25, dtype=torch.float64, device=dev))   # built on the device, as bayestree.bart does


def c4():
    gp5 = (lgp.GP(1.3 ** 2 * kb, checkpos=False, checksym=False, checkfinite=False, epsrel=0)
           .addx(xi, 'trainmean').addcov(noise4, 'trainnoise').addcov(0.49, 'mean')
           .addtransf({'trainmean': 1, 'trainnoise': 1, 'mean': 1}, 'train'))
    return gp5.marginal_likelihood({'train': y4})
out['C4_recipe_logml_n5000_ms'] = timed(c4, reps=2)
print(json.dumps(out))
