"""128x128 leaf (Cholesky + inverse of the factor in one CTA): correctness against torch and time per call for
version 1 (unblocked, register-resident) and version 2 (blocked 4 x 32, DMMA updates; chol_leaf2.cuh)."""
import sys, ctypes
import torch
sys.path.insert(0, '.')
from lsqfitgp_b200 import _lib, _ops
lib = _lib.load()
lib.lgp_debug_leaf.restype = ctypes.c_int
lib.lgp_debug_leaf.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_int64] + [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_int]
dev = torch.device('cuda:0')
torch.manual_seed(0)
A = torch.randn(128, 128, dtype=torch.float64, device=dev)
K = A @ A.T + 128 * torch.eye(128, dtype=torch.float64, device=dev)
L = torch.linalg.cholesky(K)
Linv = torch.linalg.inv(L)
for variant in (1, 2, 3):
    invd = torch.full((128, 128), 7.0, dtype=torch.float64, device=dev)
    dvec = torch.empty(128, dtype=torch.float64, device=dev)
    info = torch.full((1,), 2**31 - 1, dtype=torch.int32, device=dev)
    W = K.clone()
    rc = lib.lgp_debug_leaf(_lib.stream_ptr(), _lib.ptr(W), 128, _lib.ptr(invd), _lib.ptr(dvec), _lib.ptr(info), 1, variant)
    torch.cuda.synchronize()
    print(f'variant={variant} rc={rc}: L err {float((W - L).abs().max() / L.abs().max()):.2e} (upper zero: {bool((torch.triu(W, 1) == 0).all())}), '
          f'inverse err {float((invd - Linv).abs().max() / Linv.abs().max()):.2e}, diag err {float((dvec - torch.diagonal(L)).abs().max()):.2e}, info {int(info.item())}')
    # failure reporting: a negative pivot at column 70
    Kb = K.clone(); Kb[70, 70] = -1.0
    info.fill_(2**31 - 1)
    lib.lgp_debug_leaf(_lib.stream_ptr(), _lib.ptr(Kb), 128, _lib.ptr(invd), _lib.ptr(dvec), _lib.ptr(info), 1, variant)
    torch.cuda.synchronize()
    print(f'   bad pivot reported at {int(info.item())} (expect 71)')
    reps = 50
    W = K.clone()
    Ws = [K.clone() for _ in range(20)]
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.lgp_debug_leaf(_lib.stream_ptr(), _lib.ptr(W), 128, _lib.ptr(invd), _lib.ptr(dvec), _lib.ptr(info), reps, variant)
    e1.record(); torch.cuda.synchronize()
    print(f'   {e0.elapsed_time(e1)*1e3/reps:.1f} us per leaf')

import numpy as np
clk = (ctypes.c_longlong * 32)()
lib.lgp_debug_leaf2_clocks.argtypes = [ctypes.c_void_p]
W = K.clone()
lib.lgp_debug_leaf(_lib.stream_ptr(), _lib.ptr(W), 128, _lib.ptr(invd), _lib.ptr(dvec), _lib.ptr(info), 1, 2)
torch.cuda.synchronize()
lib.lgp_debug_leaf2_clocks(clk)
c = np.array(list(clk)[:18], dtype=np.int64)
names = ['load', 'sync'] + sum([[f'potrf{J}', f'trtri/trsm{J}', f'update{J}'] for J in range(4)], []) + ['X d=1', 'X d=2', 'X d=3', 'store']
d = np.diff(c)
print('leaf2 phases (cycles):', ', '.join(f'{n} {v}' for n, v in zip(names[1:], d)), '| total', c[-1] - c[0])

lib.lgp_debug_leaf3_clocks.argtypes = [ctypes.c_void_p]
for rep in range(3):
    W = K.clone()
    lib.lgp_debug_leaf(_lib.stream_ptr(), _lib.ptr(W), 128, _lib.ptr(invd), _lib.ptr(dvec), _lib.ptr(info), 1, 3)
    torch.cuda.synchronize()
    lib.lgp_debug_leaf3_clocks(clk)
    c = np.array(list(clk)[:10], dtype=np.int64)
    names3 = ['load'] + sum([[f'panel{J}', f'colupd{J}'] for J in range(4)], [])[:-1] + ['store']
    print('leaf3 phases (cycles):', ', '.join(f'{n} {v}' for n, v in zip(names3, np.diff(c))), '| total', c[-1] - c[0])
# fresh (valid) input for every launch: time 20 leaves on 20 different matrices
for variant in (1, 3):
    Ws = [K.clone() for _ in range(20)]
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for Wi in Ws:
        lib.lgp_debug_leaf(_lib.stream_ptr(), _lib.ptr(Wi), 128, _lib.ptr(invd), _lib.ptr(dvec), _lib.ptr(info), 1, variant)
    e1.record(); torch.cuda.synchronize()
    print(f'variant={variant}: {e0.elapsed_time(e1)*1e3/20:.1f} us per leaf (fresh inputs), L err {float((Ws[-1] - L).abs().max() / L.abs().max()):.2e}')
