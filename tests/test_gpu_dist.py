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


@pytest.mark.parametrize('storage', ['lower', 'dense'])
@pytest.mark.parametrize('n,T', [(1500, 256), (1000, 128), (513, 512), (2048, 512), (77, 128)])
def test_distchol_single_rank_vs_oracle(n, T, storage):
    X, b, descs, terms = _problem(n)
    dev = torch.device('cuda:0')
    x = torch.tensor(np.ascontiguousarray(X.T)).to(dev)
    dc = _dist.DistChol(descs, x, tile=T, storage=storage)
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
    # products with the factor and the size-independent check L (L^T v) = (K + eps S^2) v
    np.testing.assert_allclose(dc.correlate(b).cpu().numpy(), ref.correlate(b), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(dc.back_correlate(b).cpu().numpy(), ref.back_correlate(b), rtol=1e-10, atol=1e-12)
    bd = torch.tensor(b).to(dev)
    kv = dc.matvec(bd)
    assert float(((dc.correlate(dc.back_correlate(bd)) - kv).norm() / kv.norm()).item()) <= 1e-12


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
@pytest.mark.parametrize('peer,storage', [('off', 'lower'), ('on', 'lower'), ('on', 'dense')])
def test_distchol_multi_rank_vs_oracle(nproc, grid, peer, storage):
    """ peer=off: NCCL panel broadcasts; peer=on: fused TRSM -> peer-memory stores (multimem / NVLink) with counters;
    storage: lower-packed panels (default) or the dense local matrix.  Also drives the operator API on the same ranks
    (GP(..., solver='chol-dist'), GP.decompose) and the L (L^T v) = K v check (tools/dist_check.py --oracle). """
    if torch.cuda.device_count() < nproc:
        pytest.skip(f'needs {nproc} GPUs')
    res = _torchrun(nproc, '--size', '3000', '--tile', '256', '--grid', grid, '--oracle', '--peer', peer, '--storage',
                    storage)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert 'DIST_CHECK_OK' in res.stdout


def test_trsm_bcast_epilogue_and_flags_single_gpu():
    """ the fused panel solve -> broadcast entry points on ONE GPU (the multi-rank cases above need more):
    lgp_tile_trsm_right_bcast with local buffers as the extra destinations must give lgp_tile_trsm_right's result in
    place and the same values in every destination; lgp_flag_signal / lgp_flag_wait order streams through counters in
    device memory and give up (err = 1) instead of spinning forever """
    import ctypes
    dev = torch.device('cuda:0')
    ops = _dist.CudaTileOps(dev)
    T, rows = 256, 300
    rng = np.random.default_rng(9)
    A = rng.standard_normal((T, T))
    A = A @ A.T + T * np.eye(T)
    tile = _ops.as_aligned(torch.tensor(A).to(dev))
    invd = ops.empty(T // 128 * 128 * 128)
    dvec = ops.zeros(T)
    info = ops.zeros(1, dtype=torch.int32)
    info.fill_(2 ** 31 - 1)
    ops.potrf_tile(tile, invd, dvec, info, 0)
    B0 = torch.tensor(rng.standard_normal((rows, T))).to(dev)
    B1 = _ops.as_aligned(B0.clone())
    ops.trsm_right(tile, invd, B1)
    L = np.linalg.cholesky(A)
    np.testing.assert_allclose(B1.cpu().numpy(), np.linalg.solve(L, B0.cpu().numpy().T).T, rtol=1e-10, atol=1e-12)
    B2 = _ops.as_aligned(B0.clone())
    d1 = torch.zeros(rows, T, dtype=torch.float64, device=dev)
    ops.trsm_right_bcast(tile, invd, B2, [d1.data_ptr()], T, False)
    assert torch.equal(B2, B1) and torch.equal(d1, B1)
    # two destinations with a wider leading dimension: the columns beyond T stay untouched
    B3 = _ops.as_aligned(B0.clone())
    w1 = torch.full((rows, T + 2), 7.0, dtype=torch.float64, device=dev)
    w2 = torch.full((rows, T + 2), 7.0, dtype=torch.float64, device=dev)
    ops.trsm_right_bcast(tile, invd, B3, [w1.data_ptr(), w2.data_ptr()], T + 2, False)
    for w in (w1, w2):
        assert torch.equal(w[:, :T], B1) and bool((w[:, T:] == 7.0).all())
    # argument errors are reported, not executed
    lib = _lib.load()

    def call(ndst, ptrs, ld, mm):
        arr = (ctypes.c_void_p * max(len(ptrs), 1))(*ptrs)
        return lib.lgp_tile_trsm_right_bcast(_lib.stream_ptr(), _lib.ptr(tile), tile.stride(0), _lib.ptr(invd), T,
                                             _lib.ptr(B3), B3.stride(0), rows, ndst, arr, ld, mm)
    assert call(9, [w1.data_ptr()] * 9, T + 2, 0) == -1          # more than 8 destinations
    assert call(1, [w1.data_ptr()], T - 2, 0) == -1              # leading dimension shorter than the panel
    assert call(2, [w1.data_ptr(), w2.data_ptr()], T + 2, 1) == -1   # multicast takes exactly one address
    assert call(1, [w1.data_ptr() + 8], T + 2, 0) != 0           # misaligned destination
    torch.cuda.synchronize()

    # counters: a release store on one stream lets the acquire-wait on another stream pass (the signal is enqueued
    # first: a spinning kernel must never sit in front of the work that releases it in a shared hardware queue);
    # an unreachable value times out with err = 1 instead of hanging
    import time
    flags = torch.zeros(4, dtype=torch.int64, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    side = torch.cuda.Stream(dev)
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        ops.flag_signal([flags.data_ptr(), flags.data_ptr() + 8], 5)
    ops.flag_wait(flags.data_ptr(), 2, 5, 20000, err)
    t0 = time.perf_counter()
    torch.cuda.synchronize()
    assert time.perf_counter() - t0 < 10.0
    assert flags.tolist() == [5, 5, 0, 0] and int(err.item()) == 0
    ops.flag_wait(flags.data_ptr(), 2, 4, 20000, err)   # counters are monotone: a smaller value passes at once
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    t0 = time.perf_counter()
    ops.flag_wait(flags.data_ptr(), 3, 5, 300, err)   # flags[2] never reaches 5: gives up after 300 ms
    torch.cuda.synchronize()
    assert time.perf_counter() - t0 >= 0.25
    assert int(err.item()) == 1
