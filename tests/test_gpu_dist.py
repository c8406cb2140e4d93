"""GPU parity of the block-cyclic Cholesky (lsqfitgp_b200._dist.DistChol, C-ABI lgp_dist_*/lgp_tile_*) against the
oracle restatement of Chol (oracle/decomp.py <- reference _linalg/_decomp.py:380-439).
Tolerances: logdet / eps 1e-12 relative, solves 1e-9 relative (north_star: logML and posterior mean 1e-9).
Multi-rank cases launch torchrun (NCCL) and are skipped on boxes with fewer GPUs."""
import os
import pathlib
import subprocess
import sys

import numpy as np
import pytest
import torch

from lsqfitgp_b200 import _lib, _ops, _dist
from oracle import gp as ogp, decomp as odecomp

pytestmark = pytest.mark.gpu
ROOT = pathlib.Path(__file__).resolve().parent.parent


def _problem(n, seed=5005):
    rng = np.random.default_rng(seed)
    X = rng.uniform(0, 40, (n, 2))
    b = rng.standard_normal(n)
    descs = [dict(kind=_lib.K_EXPQUAD, term=0, dimmask=3, scale_x=5.0, scale_y=5.0, amp=1.7),
             dict(kind=_lib.K_WHITE, term=1, dimmask=3, amp=0.01)]
    terms = [(1.7, [dict(kind='expquad', scale=5.0)]), (0.01, [dict(kind='white')])]
    return X, b, descs, terms


@pytest.mark.parametrize('n,T', [(1500, 256), (1000, 128), (513, 512), (2048, 512), (77, 128)])
def test_distchol_single_rank_vs_oracle(n, T):
    X, b, descs, terms = _problem(n)
    dev = torch.device('cuda:0')
    x = torch.tensor(np.ascontiguousarray(X.T)).to(dev)
    dc = _dist.DistChol(descs, x, tile=T)
    K = ogp.gram(terms, X.T.copy(), X.T.copy())
    ref = odecomp.Chol(K)
    ld_ref = 2 * np.sum(np.log(np.diag(ref._L)))
    assert abs(dc.logdet() - ld_ref) <= 1e-12 * abs(ld_ref)
    np.testing.assert_allclose(dc.eps, ref.eps, rtol=1e-13)
    sol = dc.solve(b).cpu().numpy()
    np.testing.assert_allclose(sol, ref.ginv_linear(b), rtol=1e-9, atol=1e-9 * np.abs(sol).max())
    assert abs(dc.quad(b) - ref.ginv_quad(b)) <= 1e-9 * abs(ref.ginv_quad(b))
    v, *_ = ref.minus_log_normal_density(b, value=True)
    assert abs(dc.minus_log_normal_density(b) - v) <= 1e-9 * abs(v)
    # the same factor as the single-GPU path (lgp_chol_factor), to rounding
    st = _ops.chol_factor(torch.tensor(K).to(dev))
    ld1 = 2 * float(st.scalars()[4].item())
    assert abs(dc.logdet() - ld1) <= 1e-12 * abs(ld1)


def test_distchol_not_posdef():
    n = 600
    X, b, descs, terms = _problem(n)
    descs = [dict(kind=_lib.K_EXPQUAD, term=0, dimmask=3, scale_x=500.0, scale_y=500.0, amp=1.0)]  # numerically singular
    x = torch.tensor(np.ascontiguousarray(X.T)).to('cuda:0')
    with pytest.raises(np.linalg.LinAlgError):
        _dist.DistChol(descs, x, tile=128, epsrel=0.0)


def _torchrun(nproc, *args, timeout=600):
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={nproc}',
           '--master-addr', '127.0.0.1', '--master-port', '29517', str(ROOT / 'tools' / 'dist_check.py'), *args]
    return subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout, cwd=str(ROOT))


@pytest.mark.parametrize('nproc,grid', [(2, '2x1'), (2, '1x2'), (4, '2x2'), (8, '2x4')])
@pytest.mark.parametrize('peer', ['off', 'on'])
def test_distchol_multi_rank_vs_oracle(nproc, grid, peer):
    """ peer=off: NCCL panel broadcasts; peer=on: fused TRSM -> peer-memory stores (multimem / NVLink) with counters """
    if torch.cuda.device_count() < nproc:
        pytest.skip(f'needs {nproc} GPUs')
    res = _torchrun(nproc, '--size', '3000', '--tile', '256', '--grid', grid, '--oracle', '--peer', peer)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert 'DIST_CHECK_OK' in res.stdout
