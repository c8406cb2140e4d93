"""Host logic of the block-cyclic multi-GPU Cholesky (lsqfitgp_b200._dist.DistChol) on CPU: layout maps, and the
real orchestration driven over gloo (world_size 2) with the NumPy tile provider of tests/_numpy_tile_ops.py,
checked against the oracle restatement of Chol (oracle/decomp.py, reference _linalg/_decomp.py:380-439)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from lsqfitgp_b200 import _dist
from lsqfitgp_b200._dist import Layout, tiles_before, default_grid


def _matrix(n, seed=0):
    rng = np.random.default_rng(seed)
    x = np.sort(rng.uniform(0, 30, n))
    amp = np.exp(rng.normal(0, 2, n))          # non-trivial equilibration
    K = np.exp(-0.5 * (x[:, None] - x[None, :]) ** 2) + 0.05 * np.eye(n)
    return amp[:, None] * K * amp[None, :]


def test_default_grid():
    assert default_grid(1) == (1, 1)
    assert default_grid(2) == (2, 1)
    assert default_grid(4) == (2, 2)
    assert default_grid(8) == (2, 4)
    assert default_grid(6) == (2, 3)


@pytest.mark.parametrize('n,T,Pr,Pc', [(1000, 128, 2, 1), (1000, 128, 2, 4), (640, 128, 3, 2), (100, 128, 2, 2),
                                       (150000, 512, 2, 4)])
def test_layout_partition(n, T, Pr, Pc):
    lays = [Layout(n, T, Pr, Pc, r) for r in range(Pr * Pc)]
    NT = lays[0].NT
    assert NT == -(-n // T)
    # every tile owned exactly once, at a distinct local position
    seen = {}
    for lay in lays:
        rt, ct = lay.row_tiles(), lay.col_tiles()
        assert len(rt) == lay.LR and len(ct) == lay.LC
        for li, I in enumerate(rt):
            assert tiles_before(I, lay.pr, Pr) == li
            for lj, J in enumerate(ct):
                assert lay.owner(I, J) == lay.rank
                assert (I, J) not in seen
                seen[(I, J)] = (lay.rank, li, lj)
    assert len(seen) == NT * NT
    # panel bookkeeping: the slabs broadcast at step k cover the tiles I > k exactly once
    for k in sorted({0, min(1, NT - 1), NT // 2, NT - 1}):
        tiles = []
        for r in range(Pr):
            first, cnt = lays[0].panel_first(k, r), lays[0].panel_count(k, r)
            rows = lays[0].row_tiles(r)[first:first + cnt]
            assert len(rows) == cnt
            tiles += rows
        assert sorted(tiles) == list(range(k + 1, NT))
    if n <= 2000:
        g = np.sort(np.concatenate([lays[lays[0].rank_of(r, 0)].global_rows() for r in range(Pr)]))
        np.testing.assert_array_equal(g, np.arange(NT * T))


def _check_against_oracle(dc, K, b):
    from oracle import decomp as odecomp
    ref = odecomp.Chol(K)
    n = len(K)
    ld_ref = 2 * np.sum(np.log(np.diag(ref._L)))
    assert abs(dc.logdet() - ld_ref) <= 1e-11 * abs(ld_ref)
    np.testing.assert_allclose(dc.eps, ref.eps, rtol=1e-14)
    sol = dc.solve(b).numpy()
    np.testing.assert_allclose(sol, ref.ginv_linear(b), rtol=1e-8, atol=1e-10 * np.abs(sol).max())
    q_ref = ref.ginv_quad(b)
    assert abs(dc.quad(b) - q_ref) <= 1e-10 * abs(q_ref)
    v, *_ = ref.minus_log_normal_density(b, value=True)
    assert abs(dc.minus_log_normal_density(b) - v) <= 1e-10 * abs(v)
    np.testing.assert_allclose(dc.pinv_correlate(b).numpy(), ref.pinv_correlate(b), rtol=1e-8,
                               atol=1e-10 * np.abs(b).max())
    # products with the factor, and the size-independent check of SURVEY 8(d): L (L^T v) = (K + eps S^2) v
    lc, lbc = ref.correlate(b), ref.back_correlate(b)
    np.testing.assert_allclose(dc.correlate(b).numpy(), lc, rtol=1e-10, atol=1e-12 * np.abs(lc).max())
    np.testing.assert_allclose(dc.back_correlate(b).numpy(), lbc, rtol=1e-10, atol=1e-12 * np.abs(lbc).max())
    kv = dc.matvec(b).numpy()
    llv = dc.correlate(dc.back_correlate(b)).numpy()
    assert np.linalg.norm(llv - kv) <= 1e-12 * np.linalg.norm(kv)
    s = np.asarray(dc.s[:n])
    np.testing.assert_allclose(kv, K @ b + float(dc._epsout[1]) * s ** 2 * b, rtol=1e-12, atol=1e-13 * np.abs(kv).max())


@pytest.mark.parametrize('storage', ['lower', 'dense'])
@pytest.mark.parametrize('n,T', [(300, 128), (700, 256), (128, 128), (1, 128)])
def test_single_process_orchestration(n, T, storage):
    from _numpy_tile_ops import NumpyTileOps
    K = _matrix(n)
    b = np.random.default_rng(1).standard_normal(n)
    x = torch.zeros(1, n, dtype=torch.float64)
    dc = _dist.DistChol(None, x, tile=T, ops=NumpyTileOps(K), storage=storage)
    _check_against_oracle(dc, K, b)


def test_not_posdef_raises():
    from _numpy_tile_ops import NumpyTileOps
    K = _matrix(300)
    K[200, 201] = K[201, 200] = 10 * np.sqrt(K[200, 200] * K[201, 201])
    x = torch.zeros(1, 300, dtype=torch.float64)
    with pytest.raises(np.linalg.LinAlgError):
        _dist.DistChol(None, x, tile=128, ops=NumpyTileOps(K), epsrel=0)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, T, grid, q, storage='lower'):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from _numpy_tile_ops import NumpyTileOps
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        K = _matrix(n)
        b = np.random.default_rng(1).standard_normal(n)
        x = torch.zeros(1, n, dtype=torch.float64)
        dc = _dist.DistChol(None, x, tile=T, grid=grid, ops=NumpyTileOps(K), storage=storage)
        try:
            _check_against_oracle(dc, K, b)
            q.put((rank, 'ok', dc.logdet()))
        except AssertionError as e:
            q.put((rank, 'fail: ' + str(e)[:500], None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n,T,grid,storage', [(700, 128, (2, 1), 'lower'), (700, 128, (1, 2), 'lower'),
                                              (520, 256, (2, 1), 'lower'), (700, 128, (2, 1), 'dense')])
def test_two_ranks_gloo(n, T, grid, storage):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, T, grid, q, storage)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, status, ld in res:
        assert status == 'ok', (rank, status)
    assert res[0][2] == res[1][2]  # replicated result is bit-identical on the two ranks


def _peer_count(n, T, grid):
    """ doubles of the symmetric allocation DistChol asks for (mirrors DistChol.__init__) """
    Pr, Pc = grid
    lay = Layout(n, T, Pr, Pc, 0)
    slab_cap = max(max(lay.panel_count(0, r) for r in range(Pr)), 1) * T * T
    diag_cap = T * T + (T // 128) * 128 * 128 if Pr > 1 else 0
    return 2 * Pr * slab_cap + 2 * diag_cap


def _peer_worker(rank, world, port, n, T, grid, multicast, bufs, q):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from _numpy_tile_ops import NumpyPeerTileOps
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        K = _matrix(n)
        b = np.random.default_rng(1).standard_normal(n)
        x = torch.zeros(1, n, dtype=torch.float64)
        dc = _dist.DistChol(None, x, tile=T, grid=grid, ops=NumpyPeerTileOps(K, bufs, rank, multicast), peer=True)
        try:
            assert dc.peer_mode == ('multimem' if multicast else 'p2p')
            _check_against_oracle(dc, K, b)
            q.put((rank, 'ok', dc.logdet()))
        except AssertionError as e:
            q.put((rank, 'fail: ' + str(e)[:500], None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n,T,grid,multicast', [(900, 128, (2, 1), True), (900, 128, (2, 1), False),
                                                (900, 128, (1, 2), True), (520, 256, (2, 1), False)])
def test_two_ranks_fused_panel_path_shared_memory(n, T, grid, multicast):
    """ the peer-memory branch of the factorisation (slabs and diagonal tile stored into every rank's buffer by the
    producer, READY / READYD / DONE counters instead of NCCL) with two concurrent processes over shared memory:
    same results as the broadcast path and as the oracle """
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    count = _peer_count(n, T, grid)
    bufs = [torch.zeros(64 + count, dtype=torch.float64).share_memory_() for _ in range(2)]
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, n, T, grid, multicast, bufs, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, status, ld in res:
        assert status == 'ok', (rank, status)
    assert res[0][2] == res[1][2]
