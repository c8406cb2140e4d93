"""Probe what peer-memory plumbing works on the GPU box (torchrun, >= 2 ranks): torch symmetric memory (buffer_ptrs,
multicast_ptr) and whether peer pointers are directly addressable from a kernel (copy through a peer view)."""
import os, sys, traceback
import torch
import torch.distributed as dist

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); lr = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
dev = torch.device('cuda', lr)
ok = {}
try:
    import torch.distributed._symmetric_memory as sm
    t = sm.empty(1 << 20, dtype=torch.float64, device=dev)
    h = sm.rendezvous(t, dist.group.WORLD)
    ok['rendezvous'] = True
    ok['buffer_ptrs'] = [hex(p) for p in h.buffer_ptrs]
    ok['multicast_ptr'] = hex(h.multicast_ptr) if h.multicast_ptr else 0
    ok['signal_pad_size'] = h.signal_pad_size
    t.fill_(rank + 1.0)
    h.barrier()
    peer = h.get_buffer((rank + 1) % world, (1 << 20,), torch.float64)
    ok['peer_read'] = float(peer[:10].sum().item())
    h.barrier()
    # write into the peer
    peer[100:110] = 100.0 + rank
    torch.cuda.synchronize()
    h.barrier()
    ok['peer_written_seen'] = float(t[100].item())
except Exception as e:
    ok['error'] = repr(e)
    traceback.print_exc()
print(f'[rank {rank}] {ok}', flush=True)
dist.barrier()
dist.destroy_process_group()
