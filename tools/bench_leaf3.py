"""Leaf version 3 (chol_leaf3.cuh) store-mode experiments: correctness, phase clocks and time per call.
variants: 3 thread stores during the next panel, 31 all stores at the end, 32 no stores, 33 bulk stores."""
import sys, ctypes
import numpy as np
import torch
sys.path.insert(0, '.')
from lsqfitgp_b200 import _lib
lib = _lib.load()
lib.lgp_debug_leaf.restype = ctypes.c_int
lib.lgp_debug_leaf.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_int64] + [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_int]
lib.lgp_debug_leaf3_clocks.argtypes = [ctypes.c_void_p]
dev = torch.device('cuda:0')
torch.manual_seed(0)
A = torch.randn(128, 128, dtype=torch.float64, device=dev)
K = A @ A.T + 128 * torch.eye(128, dtype=torch.float64, device=dev)
L = torch.linalg.cholesky(K)
Linv = torch.linalg.inv(L)
variants = [int(a) for a in sys.argv[1].split(',')] if len(sys.argv) > 1 else [1, 3, 31, 32, 33]
clk = (ctypes.c_longlong * 32)()
names3 = ['load'] + sum([[f'panel{J}', f'colupd{J}'] for J in range(4)], [])[:-1] + ['store']
for variant in variants:
    invd = torch.full((128, 128), 7.0, dtype=torch.float64, device=dev)
    dvec = torch.empty(128, dtype=torch.float64, device=dev)
    info = torch.full((1,), 2**31 - 1, dtype=torch.int32, device=dev)
    W = K.clone()
    lib.lgp_debug_leaf(_lib.stream_ptr(), _lib.ptr(W), 128, _lib.ptr(invd), _lib.ptr(dvec), _lib.ptr(info), 1, variant)
    torch.cuda.synchronize()
    print(f'variant={variant}: L err {float((W - L).abs().max() / L.abs().max()):.2e} upper zero {bool((torch.triu(W, 1) == 0).all())} '
          f'inverse err {float((invd - Linv).abs().max() / Linv.abs().max()):.2e} diag err {float((dvec - torch.diagonal(L)).abs().max()):.2e} info {int(info.item())}')
    if variant >= 3:
        for rep in range(2):
            W = K.clone()
            lib.lgp_debug_leaf(_lib.stream_ptr(), _lib.ptr(W), 128, _lib.ptr(invd), _lib.ptr(dvec), _lib.ptr(info), 1, variant)
            torch.cuda.synchronize()
            lib.lgp_debug_leaf3_clocks(clk)
            c = np.array(list(clk)[:10], dtype=np.int64)
            print('   phases (cycles):', ', '.join(f'{n} {v}' for n, v in zip(names3, np.diff(c))), '| total', c[-1] - c[0])
    for trial in range(2):
        Ws = [K.clone() for _ in range(50)]
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for Wi in Ws:
            lib.lgp_debug_leaf(_lib.stream_ptr(), _lib.ptr(Wi), 128, _lib.ptr(invd), _lib.ptr(dvec), _lib.ptr(info), 1, variant)
        e1.record(); torch.cuda.synchronize()
        print(f'   {e0.elapsed_time(e1)*1e3/50:.1f} us per leaf (50 fresh inputs)')
