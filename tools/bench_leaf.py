import sys, ctypes
import torch
sys.path.insert(0, '.')
from lsqfitgp_b200 import _lib, _ops
lib = _lib.load()
lib.lgp_debug_leaf.restype = ctypes.c_int
lib.lgp_debug_leaf.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_int64] + [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_int]
dev = torch.device('cuda:0')
A = torch.randn(128, 128, dtype=torch.float64, device=dev)
K = A @ A.T + 128 * torch.eye(128, dtype=torch.float64, device=dev)
invd = torch.empty(128, 128, dtype=torch.float64, device=dev)
dvec = torch.empty(128, dtype=torch.float64, device=dev)
info = torch.full((1,), 2**31 - 1, dtype=torch.int32, device=dev)
W = K.clone()
lib.lgp_debug_leaf(_lib.stream_ptr(), _lib.ptr(W), 128, _lib.ptr(invd), _lib.ptr(dvec), _lib.ptr(info), 1, 0)
torch.cuda.synchronize()
L = torch.linalg.cholesky(K)
print('leaf err', float((torch.tril(W) - L).abs().max()), 'inv err', float((invd - torch.linalg.inv(L)).abs().max()))
for variant in (0,):
    reps = 50
    W = K.clone()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.lgp_debug_leaf(_lib.stream_ptr(), _lib.ptr(W), 128, _lib.ptr(invd), _lib.ptr(dvec), _lib.ptr(info), reps, variant)
    e1.record(); torch.cuda.synchronize()

    print(f'variant={variant}: {e0.elapsed_time(e1)*1e3/reps:.1f} us per leaf')
