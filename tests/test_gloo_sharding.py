"""N>1 path on CPU: world_size-2 gloo run of the batch-of-hyperparameters sharding (config 3 layout)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lsqfitgp_b200 import _dist


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _fun(theta):
    # stand-in for (logML, grad): any deterministic function of the hyperparameter point
    return np.array([np.sum(theta ** 2), *np.sin(theta)])


def _worker(rank, world, port, B, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        thetas = np.random.default_rng(3004).standard_normal((B, 3))
        calls = []

        def fun(t):
            calls.append(1)
            return _fun(t)
        out = _dist.eval_batch_sharded(fun, thetas)
        q.put((rank, len(calls), out))
    finally:
        dist.destroy_process_group()


def _run(B, world=2):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_shard_indices():
    assert _dist.shard_indices(7, 0, 2) == [0, 2, 4, 6]
    assert _dist.shard_indices(7, 1, 2) == [1, 3, 5]
    allidx = sorted(sum((_dist.shard_indices(64, r, 8) for r in range(8)), []))
    assert allidx == list(range(64))


def test_single_process():
    thetas = np.random.default_rng(1).standard_normal((5, 3))
    out = _dist.eval_batch_sharded(_fun, thetas)
    np.testing.assert_array_equal(out, np.stack([_fun(t) for t in thetas]))


def test_two_ranks_gloo():
    for B in (7, 2, 1):
        res = _run(B)
        thetas = np.random.default_rng(3004).standard_normal((B, 3))
        expect = np.stack([_fun(t) for t in thetas])
        ncalls = 0
        for rank, calls, out in res:
            np.testing.assert_array_equal(out, expect)  # bit-identical on every rank
            assert calls == len(_dist.shard_indices(B, rank, 2))
            ncalls += calls
        assert ncalls == B  # every point evaluated exactly once


def test_eval_concurrent_host_logic():
    """ slot assignment, ordering and error propagation of the in-flight evaluator (CPU: threads only, no streams) """
    import threading
    seen = {}

    def fun(t):
        seen.setdefault(threading.current_thread().name, []).append(int(t[0]))
        return _fun(t)
    thetas = np.arange(21.0).reshape(7, 3)
    out = _dist.eval_concurrent(fun, thetas, 3)
    np.testing.assert_array_equal(np.stack(out), np.stack([_fun(t) for t in thetas]))
    assert sorted(sorted(v) for v in seen.values()) == [[0, 9, 18], [3, 12], [6, 15]]
    assert _dist.eval_concurrent(_fun, [], 4) == []
    out1 = _dist.eval_batch_sharded(_fun, thetas, in_flight=2)
    np.testing.assert_array_equal(out1, np.stack([_fun(t) for t in thetas]))

    def bad(t):
        if t[0] == 9:
            raise ValueError('boom')
        return _fun(t)
    with pytest.raises(ValueError):
        _dist.eval_concurrent(bad, thetas, 2)
