"""Where the time of BASELINE configs[0] (n = 1000: marginal_likelihood + predfromdata on 500 points) goes: wall clock and
CUDA-event time of each public-API call, kernel launches of the library, and a cProfile of the host side."""
import cProfile
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
import lsqfitgp_b200 as lgp  # noqa: E402
from lsqfitgp_b200 import _lib  # noqa: E402

lib = _lib.load()
rng = np.random.default_rng(1001)
x1 = np.sort(rng.uniform(0, 100, 1000))
y1 = np.sin(x1 / 3) + 0.1 * rng.standard_normal(1000)
xp = np.linspace(-5, 105, 500)
ycov1 = {('d', 'd'): 0.01 * np.eye(1000)}


def seg(fn):
    torch.cuda.synchronize()
    l0 = lib.lgp_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t0) * 1e3, e0.elapsed_time(e1), lib.lgp_launch_count() - l0


def run(verbose):
    gp, a = seg(lambda: lgp.GP(lgp.ExpQuad(scale=3), checkpos=False, checksym=False).addx(x1, 'd').addx(xp, 'p'))[0:2]
    r1 = seg(lambda: gp.marginal_likelihood({'d': y1}, ycov1))
    r2 = seg(lambda: gp.predfromdata({'d': y1}, 'p', ycov1, raw=True))
    if verbose:
        print(f'build+addx {a:.3f} ms | marginal_likelihood wall {r1[1]:.3f} dev {r1[2]:.3f} launches {r1[3]} | '
              f'predfromdata wall {r2[1]:.3f} dev {r2[2]:.3f} launches {r2[3]}')


for _ in range(3):
    run(False)
for _ in range(3):
    run(True)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    run(False)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
