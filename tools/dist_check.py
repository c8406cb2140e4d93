"""torchrun entry: block-cyclic Cholesky on N GPUs (NCCL), checked against the oracle (small n) and/or through
size-independent properties (residual |Kx - b|/|b| with K regenerated slab-wise on the fly), with timings.

    torchrun --nproc-per-node N tools/dist_check.py --size 3000 --tile 256 --grid 2x1 --oracle
    torchrun --nproc-per-node 8 tools/dist_check.py --size 150000 --tile 512
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsqfitgp_b200 import _lib, _ops, _dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', dest='n', type=int, default=3000)
    ap.add_argument('--tile', type=int, default=512)
    ap.add_argument('--grid', default=None)
    ap.add_argument('--oracle', action='store_true')
    ap.add_argument('--reps', type=int, default=1)
    ap.add_argument('--scale', type=float, default=5.0)
    ap.add_argument('--box', type=float, default=None, help='points are U(0, box)^2; default keeps the density of C5')
    ap.add_argument('--out', default=None)
    ap.add_argument('--peer', default='auto', choices=['auto', 'on', 'off'],
                    help='panel broadcast: fused TRSM -> peer-memory stores (on), NCCL broadcasts (off), or auto')
    ap.add_argument('--storage', default='lower', choices=['lower', 'dense'])
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        opts = None
        try:
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        except Exception:
            pass
        dist.init_process_group('nccl', device_id=dev, pg_options=opts)
    grid = tuple(int(v) for v in args.grid.split('x')) if args.grid else None

    n = args.n
    # X = U(0,1000,(150000,2)), K = ExpQuad(scale=5) + 0.01 I, b = N(0,1);
    # smaller n keeps the point density (box ~ sqrt(n)) so that conditioning is comparable
    box = args.box if args.box is not None else 1000.0 * np.sqrt(n / 150000.0)
    rng = np.random.default_rng(5005)
    X = rng.uniform(0, box, (n, 2))
    b = rng.standard_normal(n)
    descs = [dict(kind=_lib.K_EXPQUAD, term=0, dimmask=3, scale_x=args.scale, scale_y=args.scale, amp=1.0),
             dict(kind=_lib.K_WHITE, term=1, dimmask=3, amp=0.01)]
    x = torch.tensor(np.ascontiguousarray(X.T)).to(dev)
    bd = torch.tensor(b).to(dev)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    times = []
    dc = None
    for rep in range(args.reps):
        del dc
        sync()
        timers = []
        t0 = time.perf_counter()
        dc = _dist.DistChol(descs, x, tile=args.tile, grid=grid, timers=timers,
                            peer={'auto': 'auto', 'on': True, 'off': False}[args.peer], storage=args.storage)
        sync()
        t1 = time.perf_counter()
        ph = {timers[i][0]: timers[i][1] - timers[i - 1][1] for i in range(1, len(timers))}
        times.append(dict(total=t1 - t0, **ph))
    sync()
    t0 = time.perf_counter()
    sol = dc.solve(bd)
    sync()
    t_solve = time.perf_counter() - t0
    logdet = dc.logdet()

    # residual |(K + jitter) x - b| / |b| with K regenerated slab-wise (rows sharded over ranks); checker code, not product
    rows = np.array_split(np.arange(n), world)[rank]
    res2 = torch.zeros(1, dtype=torch.float64, device=dev)
    slab = 2048
    epsj = float(dc._epsout[1].item())
    for r0 in range(0, len(rows), slab):
        idx = torch.as_tensor(rows[r0:r0 + slab], device=dev)
        Ks = _ops.gram_iso(descs, x.index_select(1, idx).contiguous(), x)
        # the factor is that of K + eps*diag(s^2) (Chol.__init__ jitter, reference _decomp.py:384-387)
        r = Ks @ sol + epsj * (dc.s[idx] ** 2) * sol[idx] - bd[idx]
        res2 += (r * r).sum()
    if world > 1:
        dist.all_reduce(res2)
    resid = float(res2.sqrt().item()) / float(np.linalg.norm(b))

    ok = resid <= 1e-10
    out = dict(n=n, tile=args.tile, world=world, grid=[dc.lay.Pr, dc.lay.Pc], logdet=logdet, resid=resid,
               t_solve=t_solve, times=times, info=dc._info, peer_mode=getattr(dc, 'peer_mode', 'nccl'), storage=args.storage,
               factor_ms_device=dc.factor_ms(),
               factor_tflops=[n ** 3 / 3 / t['factor'] / 1e12 for t in times])
    if args.oracle:
        from oracle import gp as ogp, decomp as odecomp
        terms = [(1.0, [dict(kind='expquad', scale=args.scale)]), (0.01, [dict(kind='white')])]
        K = ogp.gram(terms, X.T.copy(), X.T.copy())
        ref = odecomp.Chol(K)
        ld_ref = 2 * np.sum(np.log(np.diag(ref._L)))
        sol_ref = ref.ginv_linear(b)
        e_ld = abs(logdet - ld_ref) / abs(ld_ref)
        e_sol = float(np.max(np.abs(sol.cpu().numpy() - sol_ref)) / np.max(np.abs(sol_ref)))
        e_eps = abs(dc.eps - ref.eps) / ref.eps
        out.update(err_logdet=e_ld, err_solve=e_sol, err_eps=e_eps)
        ok = ok and e_ld <= 1e-12 and e_sol <= 1e-9 and e_eps <= 1e-13
        # products with the factor and the size-independent L (L^T v) = (K + eps S^2) v check
        lc = ref.correlate(b)
        e_cor = float(np.max(np.abs(dc.correlate(bd).cpu().numpy() - lc)) / np.max(np.abs(lc)))
        kv = dc.matvec(bd)
        e_llt = float(((dc.correlate(dc.back_correlate(bd)) - kv).norm() / kv.norm()).item())
        out.update(err_correlate=e_cor, err_llt=e_llt)
        ok = ok and e_cor <= 1e-11 and e_llt <= 1e-12
        # the operator API on the same ranks: GP(..., solver='chol-dist') (tiles generated from the kernel, no n x n
        # matrix anywhere) and GP.decompose(K, solver='chol-dist') (replicated matrix cut into tiles)
        import lsqfitgp_b200 as lgp
        xs = lgp.unstructured_to_structured(X, names=['a', 'b'])
        kern = lgp.ExpQuad(scale=args.scale) + 0.01 * lgp.White()
        gp = lgp.GP(kern, solver='chol-dist', tile=args.tile, grid=grid, checkpos=False, checksym=False).addx(xs, 'd') \
            .addx(xs[:25], 'p')
        ml = gp.marginal_likelihood({'d': b})
        ml_ref = ogp.logml(K, b)
        mean, cov = gp.predfromdata({'d': b}, 'p', raw=True)
        mean_ref, cov_ref = ogp.pred(K, K[:, :25], K[:25, :25], b)
        dd = lgp.GP.decompose(K, solver='chol-dist', tile=args.tile, grid=grid)
        e_api = max(abs(ml - ml_ref) / abs(ml_ref), float(np.max(np.abs(mean - mean_ref)) / np.max(np.abs(mean_ref))),
                    float(np.max(np.abs(dd.ginv_linear(b) - sol_ref)) / np.max(np.abs(sol_ref))))
        out.update(err_api=e_api)
        ok = ok and e_api <= 1e-9
    if rank == 0:
        print(json.dumps(out))
        if args.out:
            with open(args.out, 'w') as f:
                json.dump(out, f)
        print('DIST_CHECK_OK' if ok else 'DIST_CHECK_FAIL')
    if world > 1:
        dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
