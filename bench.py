#!/usr/bin/env python
"""bench.py -- headline benchmark of the GP-fitting hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--n 20000]

Workload (config.workload): BASELINE.json configs[1] -- Matern(nu=2.5), 3-D covariates, n = 20000, float64,
one step = one evaluation of log marginal likelihood AND its gradient w.r.t. (log scale, log sigma_f, log sigma_n):
Gram build -> equilibrated/jittered Cholesky -> triangular solves + log-determinant -> inverse from the factor ->
Gram-VJP contraction.  Metric: logML(+gradient) evaluations per second, whole job.

  value : device-timed (CUDA events), inputs (x, y) already resident in HBM, through the C ABI.
  e2e   : the same metric through the public API (lgp.GP(...).marginal_likelihood + torch.autograd.grad) from HOST
          arrays, host->device copies of x, y and the device->host read of (logML, gradient) inside the timed region.
  N > 1 : one process per GPU (torchrun); every rank evaluates its own hyperparameter point of a batch on replicated
          data (independent units, no data-path collective) -> weak scaling; time = max over ranks.
  --impl reference : the CPU restatement of the reference path (oracle/, NumPy/SciPy/OpenBLAS on all host cores;
          jax/gvar are not installable here so the reference itself cannot run) on a bounded sample of the same workload.
"""

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'logML+gradient evaluations/s at n=20000 fp64 (Gram+Chol+solve+inverse+VJP)'
UNIT = 'evals/s'
FP64_DMMA_PEAK_TFLOPS = 37.0   # measured on this pool's B200 with tools/peaks_fp64.cu (profiles/peaks_fp64_r1.log);
                               # MEASURED_PEAKS.json has no FP64 entry


CHOL_TRAFFIC_BYTES_N20000 = 62.88e9  # measured, see roofline.traffic_source


def make_data(n, seed=2002):
    """ SURVEY.md section 8(d), config C2 """
    rng = np.random.default_rng(seed)
    X = rng.uniform(0, 10, (n, 3))
    y = np.sin(X[:, 0]) + np.cos(X[:, 1]) * X[:, 2] / 10 + 0.1 * rng.standard_normal(n)
    return X, y


def theta_for(rank, step):
    """ hyperparameter point of the batch evaluated by `rank` at `step`: (log ell, log sigma_f, log sigma_n) """
    rng = np.random.default_rng(3004 + 1000 * rank + step)
    return np.array([np.log(1.5), 0.0, np.log(0.1)]) + 0.05 * rng.standard_normal(3)


# --------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# --------------------------------------------------------------------------------------------------------------

def cpu_eval(X, y, theta):
    from oracle import gp as ogp
    ell, sf, sn = np.exp(theta)
    terms = [(sf ** 2, [dict(kind='matern', nu=2.5, scale=ell)]), (sn ** 2, [dict(kind='white')])]
    val, g = ogp.logml_and_grad(terms, X.T.copy(), y, [('logscale', 0, 0), ('amp', 0), ('amp', 1)])
    return val, np.array([g[0], g[1] * 2 * sf ** 2, g[2] * 2 * sn ** 2])


def cpu_threads():
    try:
        import threadpoolctl
        infos = threadpoolctl.threadpool_info()
        if infos:
            return max(i.get('num_threads', 1) for i in infos)
    except Exception:
        pass
    return os.cpu_count() or 1


def cpu_baseline(n_full, n_sample, reps=1):
    """ time the oracle on a bounded sample (n_sample points) and scale by (n_full/n_sample)^3 (the path is O(n^3)) """
    X, y = make_data(n_sample)
    cpu_eval(X[:500], y[:500], theta_for(0, 0))  # warm up BLAS threads
    ts = []
    for r in range(reps):
        t0 = time.perf_counter()
        cpu_eval(X, y, theta_for(0, r))
        ts.append(time.perf_counter() - t0)
    t = min(ts)
    scale = (n_full / n_sample) ** 3
    return dict(value=1.0 / (t * scale), unit=UNIT, cores=cpu_threads(), kind='port',
                sample=f'oracle (NumPy/SciPy OpenBLAS restatement of the reference path) logML+gradient at n={n_sample} '
                       f'in {t:.2f} s, extrapolated to n={n_full} by (n/n_sample)^3 = {scale:.1f}x',
                seconds_sample=t)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    n_sample = args.ref_n
    X, y = make_data(n_sample)
    cpu_eval(X[:500], y[:500], theta_for(0, 0))
    for w in range(args.warmup):
        cpu_eval(X, y, theta_for(0, w))
    t0 = time.perf_counter()
    for s in range(args.steps):
        cpu_eval(X, y, theta_for(0, s))
    t = (time.perf_counter() - t0) / args.steps
    scale = (args.n / n_sample) ** 3
    value = 1.0 / (t * scale)
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=t * scale * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64',
                data='synthetic', impl='reference',
                config=dict(workload=f'Matern(nu=2.5) 3-D n={args.n} fp64 logML+gradient (BASELINE configs[1])',
                            n=args.n, d=3, kernel='sf^2*Matern(nu=2.5, scale=ell) + sn^2*White'),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cpu_threads(), kind='port',
                                  sample=f'each step = oracle logML+gradient at n={n_sample} ({t:.2f} s), scaled to '
                                         f'n={args.n} by (n/n_sample)^3 = {scale:.1f}x; the reference itself needs '
                                         'jax+gvar, not installable here'),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------------

def dist_chol_measure(n, tile, dev, rank, world, reps=1):
    """ Block-cyclic multi-GPU Cholesky (BASELINE.json configs[4] / SURVEY.md 8(d) C5): X = U(0,1000)^2 (density kept
    for other n), K = ExpQuad(scale=5) + 0.01 I generated tile-wise in place, factor, solve K x = b, logdet.
    Device-timed (CUDA events on each rank's main stream), max over ranks. """
    import torch
    import torch.distributed as dist
    from lsqfitgp_b200 import _lib, _dist
    box = 1000.0 * math.sqrt(n / 150000.0)
    rng = np.random.default_rng(5005)
    X = rng.uniform(0, box, (n, 2))
    b = rng.standard_normal(n)
    descs = [dict(kind=_lib.K_EXPQUAD, term=0, dimmask=3, scale_x=5.0, scale_y=5.0, amp=1.0),
             dict(kind=_lib.K_WHITE, term=1, dimmask=3, amp=0.01)]
    Xh = torch.tensor(np.ascontiguousarray(X.T)).pin_memory()
    bh = torch.tensor(b).pin_memory()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up at a small size: NCCL broadcast channels, kernel attributes, allocator
    w = _dist.DistChol(descs, Xh[:, :4096].to(dev), tile=min(tile, 512))
    w.solve(bh[:4096].to(dev))
    del w
    _dist.peer_reserve(n, tile, dev)  # pooled peer-mapped slab buffers of the fused TRSM -> broadcast path (allocation)
    best = None
    for _ in range(reps):
        sync()
        t0 = time.perf_counter()
        x = Xh.to(dev, non_blocking=True)                       # e2e: host -> device copy of the points
        dc = _dist.DistChol(descs, x, tile=tile)
        ld = dc.logdet()                                        # device -> host read of the result
        torch.cuda.synchronize()
        t_e2e = time.perf_counter() - t0
        fms = dc.factor_ms()
        sync()
        t0 = time.perf_counter()
        sol = dc.solve(bh.to(dev))
        torch.cuda.synchronize()
        t_solve = time.perf_counter() - t0
        tt = torch.tensor([fms, t_e2e * 1e3, t_solve * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        fms, e2e_ms, solve_ms = (float(v) for v in tt.cpu())
        if best is None or fms < best['factor_ms']:
            best = dict(factor_ms=fms, e2e_ms=e2e_ms, solve_ms=solve_ms, logdet=ld)
        # cheap size-independent check: residual of the jittered system on 512 sampled rows
        idx = torch.as_tensor(np.random.default_rng(1).choice(n, 512, replace=False), device=dev)
        from lsqfitgp_b200 import _ops
        Ks = _ops.gram_iso(descs, x.index_select(1, idx).contiguous(), x)
        r = Ks @ sol + float(dc._epsout[1].item()) * (dc.s[idx] ** 2) * sol[idx] - bh.to(dev)[idx]
        best['resid_sampled'] = float(r.norm().item() / bh[idx.cpu()].norm().item())
        grid = [dc.lay.Pr, dc.lay.Pc]
        peer_mode = getattr(dc, 'peer_mode', 'nccl')
        del dc, sol, Ks
        torch.cuda.empty_cache()
    flops = n ** 3 / 3
    tf = flops / (best['factor_ms'] * 1e-3) / 1e12
    return dict(n=n, tile=tile, grid=grid, n_gpus=world, panel_broadcast=peer_mode, factor_ms=best['factor_ms'],
                factor_TFLOPs=tf,
                per_gpu_TFLOPs=tf / world, frac_of_dmma_peak=tf / world / FP64_DMMA_PEAK_TFLOPS,
                e2e_ms=best['e2e_ms'], e2e_TFLOPs=flops / (best['e2e_ms'] * 1e-3) / 1e12,
                solve_ms=best['solve_ms'], logdet=best['logdet'], resid_sampled=best['resid_sampled'],
                h2d_bytes=n * 2 * 8, d2h_bytes=8,
                note='n^3/3 flop; Gram generated in place by tile owners; e2e = H2D of points + Gram + equilibration + '
                     'factor + logdet readback; panel_broadcast: multimem = TRSM epilogue stores through the NVSwitch '
                     'multicast mapping, p2p = one NVLink store per peer, nccl = copy + ncclBroadcast')


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = threading.Event()
        self.maxmhz = None

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--id={self.index}', f'--query-gpu={q}',
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(',')]
                self.samples.append(float(parts[0]))
                self.maxmhz = float(parts[1])
                for nm, v in zip(names, parts[2:]):
                    if v.lower().startswith('active'):
                        self.reasons.add(nm)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=self.maxmhz, reasons=sorted(self.reasons))
        return dict(sm_mhz=float(np.median(self.samples)), sm_max_mhz=self.maxmhz, reasons=sorted(self.reasons),
                    samples=len(self.samples))


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import lsqfitgp_b200 as lgp
    from lsqfitgp_b200 import _lib, _ops

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()

    n = args.n
    X, y = make_data(n)
    xd = torch.tensor(np.ascontiguousarray(X.T)).to(dev)          # (3, n) field-major, resident in HBM
    yd = torch.tensor(y).to(dev)
    names = ['f0', 'f1', 'f2']

    def descs_for(theta):
        ell, sf, sn = np.exp(theta)
        return [dict(kind=_lib.K_MATERNP, term=0, dimmask=7, ipar=2, par0=0.0, scale_x=ell, scale_y=ell, amp=sf ** 2),
                dict(kind=_lib.K_WHITE, term=1, dimmask=7, amp=sn ** 2)]

    phases = ['gram', 'chol', 'solve', 'inverse', 'vjp']  # 'inverse' = what is left of it after the overlapped solves
    K = _ops.aligned_empty(n, n, dev)
    side = torch.cuda.Stream(dev)

    def step_device(theta, ev=None):
        """ one logML+gradient evaluation with device-resident inputs, straight through the C ABI """
        descs = descs_for(theta)

        def mark(i):
            if ev is not None:
                ev[i].record()
        mark(0)
        _ops.gram_iso(descs, xd, xd, out=K, symmetric=True)
        mark(1)
        st = _ops.chol_factor(K)
        mark(2)
        # inverse-from-factor on a side stream right behind the factorisation: the latency-bound triangular solves on
        # the main stream overlap its first GEMMs (the public API does the same in _GP._FusedNegLogMLFn.forward)
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            Kinv = _ops.chol_inverse(st)
        a = _ops.chol_solve(st, yd[:, None], False)
        ldq = _ops.chol_logdet_quad(st, a[:, 0].contiguous())
        b = _ops.chol_solve(st, a, True, inplace=True)
        mark(3)
        main.wait_stream(side)
        Kinv.record_stream(main)
        mark(4)
        vjp = _ops.gram_iso_vjp(descs, xd, Kinv, b[:, 0].contiguous())
        mark(5)
        return ldq, vjp, st

    def finish(theta, ldq, vjp):
        ld, q = ldq.cpu().numpy()
        v = vjp.cpu().numpy()
        ell, sf, sn = np.exp(theta)
        val = 0.5 * (n * math.log(2 * math.pi) + 2 * ld + q)
        grad = 0.5 * np.array([v[0, 1], v[0, 0] * 2 * sf ** 2, v[1, 0] * 2 * sn ** 2])
        return val, grad

    # pinned host staging for the end-to-end arm
    Xh = torch.empty((n, 3), dtype=torch.float64).pin_memory()
    Xh.copy_(torch.tensor(X))
    yh = torch.empty(n, dtype=torch.float64).pin_memory()
    yh.copy_(torch.tensor(y))
    Xnp, ynp = Xh.numpy(), yh.numpy()

    def step_e2e(theta):
        """ public API from host buffers: H2D of x and y, D2H of (logML, gradient) """
        th = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
        ell, sf, sn = torch.exp(th[0]), torch.exp(th[1]), torch.exp(th[2])
        kern = sf ** 2 * lgp.Matern(nu=2.5, scale=ell) + sn ** 2 * lgp.White()
        xs = lgp.unstructured_to_structured(Xnp, names=names)
        gp = lgp.GP(kern, checkpos=False, checksym=False, checkfinite=False).addx(xs, 'data')
        ml = gp.marginal_likelihood({'data': ynp})
        g, = torch.autograd.grad(ml, th)
        return float(ml.detach()), g.numpy()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness guard: the two arms agree (and, at rank 0, with the oracle on a subsample in tests/)
    th0 = theta_for(rank, 0)
    ldq, vjp, st = step_device(th0)
    v_dev, g_dev = finish(th0, ldq, vjp)
    assert int(st.info.item()) == 0
    del st
    v_api, g_api = step_e2e(th0)
    assert abs(v_dev + v_api) <= 1e-9 * abs(v_dev), (v_dev, v_api)
    assert np.max(np.abs(g_dev + g_api)) <= 1e-9 * np.max(np.abs(g_dev)), (g_dev, g_api)

    # ---- device-resident arm
    for w in range(args.warmup):
        ldq, vjp, st = step_device(theta_for(rank, w))
        del st
    barrier()
    launches0 = lib.lgp_launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    results = []
    e0.record()
    for s in range(args.steps):
        ldq, vjp, st = step_device(theta_for(rank, s), evs[s])
        results.append((ldq, vjp))
        del st
    e1.record()
    barrier()
    sampler.stop_flag.set()
    launches = lib.lgp_launch_count() - launches0
    t_dev = e0.elapsed_time(e1) / 1e3
    phase_ms = {p: float(np.mean([evs[s][i].elapsed_time(evs[s][i + 1]) for s in range(args.steps)]))
                for i, p in enumerate(phases)}

    # ---- end-to-end arm
    for w in range(min(args.warmup, 2)):
        step_e2e(theta_for(rank, w))
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        step_e2e(theta_for(rank, s))
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    barrier()

    # ---- batch arm (extra key, not the headline): independent hyperparameter points of one batch kept `in_flight` at a
    # time on this GPU (one host thread + stream per slot, lsqfitgp_b200._dist.eval_concurrent): throughput of the same
    # evaluations when the panel chains of one factorisation overlap the GEMMs of another
    batch = None
    if args.in_flight > 1:
        from lsqfitgp_b200 import _dist
        nb = max(args.steps, 2 * args.in_flight)
        Ks = [K] + [_ops.aligned_empty(n, n, dev) for _ in range(args.in_flight - 1)]

        def batch_fun(item):
            i, theta = item
            descs = descs_for(theta)
            Ki = Ks[i % args.in_flight]
            _ops.gram_iso(descs, xd, xd, out=Ki, symmetric=True)
            st = _ops.chol_factor(Ki)
            a = _ops.chol_solve(st, yd[:, None], False)
            ldq = _ops.chol_logdet_quad(st, a[:, 0].contiguous())
            b = _ops.chol_solve(st, a, True, inplace=True)
            Kinv = _ops.chol_inverse(st)
            vjp = _ops.gram_iso_vjp(descs, xd, Kinv, b[:, 0].contiguous())
            return finish(theta, ldq, vjp)
        items = [(i, theta_for(rank, i)) for i in range(nb)]
        _dist.eval_concurrent(batch_fun, items[:args.in_flight], args.in_flight, dev)  # warm-up: per-stream setup
        barrier()
        t0 = time.perf_counter()
        res = _dist.eval_concurrent(batch_fun, items, args.in_flight, dev)
        torch.cuda.synchronize()
        t_batch = time.perf_counter() - t0
        barrier()
        v0, g0 = finish(items[0][1], *results[0]) if args.steps else (None, None)
        if v0 is not None:  # same point as step 0 of the device arm: same value
            assert abs(res[0][0] - v0) <= 1e-12 * abs(v0), (res[0][0], v0)
        del Ks
        if world > 1:
            tb = torch.tensor([t_batch], dtype=torch.float64, device=dev)
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
            t_batch = float(tb.item())
        batch = dict(value=nb * world / t_batch, unit=UNIT, in_flight=args.in_flight, evaluations_per_gpu=nb,
                     timing='host wall clock between device synchronisations (several streams)')

    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = (float(v) for v in tt.cpu())
    sampler.join(timeout=2)

    dist_chol = None
    if world > 1 and args.dist_n > 0:
        del K
        torch.cuda.empty_cache()
        try:
            dist_chol = dist_chol_measure(args.dist_n, args.dist_tile, dev, rank, world)
        except Exception as e:  # never lose the headline line to the secondary measurement
            dist_chol = dict(error=repr(e)[:300])

    if rank == 0:
        value = args.steps * world / t_dev
        e2e_value = args.steps * world / t_e2e
        chol_tflops = n ** 3 / 3 / (phase_ms['chol'] * 1e-3) / 1e12
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=t_dev / args.steps * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None,
            dtype='f64', data='synthetic',
            config=dict(workload=f'Matern(nu=2.5) 3-D n={n} fp64 logML+gradient (BASELINE configs[1])', n=n, d=3,
                        kernel='sf^2*Matern(nu=2.5, scale=ell) + sn^2*White', parallelism=f'hyperparameter-batch x{world}',
                        l2='working set (3.2 GB matrix) exceeds the 126 MB L2; no flush needed'),
            clocks=sampler.summary(),
            e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=n * 3 * 8 + n * 8, d2h_bytes_per_step=4 * 8,
                     ms_per_step=t_e2e / args.steps * 1e3,
                     api='lgp.GP(kernel).addx(x).marginal_likelihood({..}) + torch.autograd.grad'),
            gpu_launches=int(launches),
            roofline=dict(bound='tensor', kernel='gemm_dmma_kernel (Cholesky phase: lgp_chol_factor)',
                          achieved=chol_tflops, peak=FP64_DMMA_PEAK_TFLOPS, unit='TFLOP/s',
                          frac=chol_tflops / FP64_DMMA_PEAK_TFLOPS,
                          traffic=CHOL_TRAFFIC_BYTES_N20000 if n == 20000 else None,
                          traffic_source='ncu dram__bytes_read.sum + dram__bytes_write.sum over the kernels of one '
                                         'lgp_chol_factor call at n=20000 (profiles/traffic_chol20k_r1.txt): 42.8 GB read + '
                                         '20.1 GB written; the rank-512 right-looking update inherently re-reads the '
                                         'trailing matrix once per panel (~43 GB); tensor-bound, not HBM-bound',
                          peak_source='measured FP64 DMMA.8x8x4 register-resident loop, tools/peaks_fp64.cu '
                                      '(MEASURED_PEAKS.json has no FP64 entry)',
                          algorithmic_flops_per_launch=n ** 3 / 3),
            phases_ms=phase_ms,
            phase_rates=dict(gram_GBps=8 * n * n / (phase_ms['gram'] * 1e-3) / 1e9,
                             chol_TFLOPs=chol_tflops,
                             inverse_TFLOPs=2 * n ** 3 / 3 / ((phase_ms['solve'] + phase_ms['inverse']) * 1e-3) / 1e12,  # its whole span
                             vjp_GBps=4 * n * n / (phase_ms['vjp'] * 1e-3) / 1e9,
                             step_TFLOPs=n ** 3 / (t_dev / args.steps) / 1e12 / world * world),
        )
        if batch is not None:
            line['batch_throughput'] = batch
        if dist_chol is not None:
            line['dist_chol'] = dist_chol
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline(n, args.ref_n)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--n', type=int, default=20000)
    ap.add_argument('--ref-n', type=int, default=5000, help='sample size of the CPU baseline / reference arm')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--dist-n', type=int, default=150000,
                    help='N > 1 only: size of the block-cyclic multi-GPU Cholesky reported under "dist_chol" (0 = skip)')
    ap.add_argument('--dist-tile', type=int, default=1024)
    ap.add_argument('--in-flight', type=int, default=2,
                    help='extra key batch_throughput: the same evaluations with this many in flight per GPU (0: skip)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    return run_gpu(args)


if __name__ == '__main__':
    sys.exit(main())
