"""Multi-GPU: sharding of independent hyperparameter evaluations (one process per GPU, torch.distributed).

The reference has no parallelism of any kind; its optimiser evaluates one hyperparameter
point at a time (src/lsqfitgp/_fit.py:338).  Batches of evaluations (multi-start fits, line-search fans, grids)
are independent units: X and y are replicated (n*d*8 bytes), every rank runs the full single-GPU path on its
share of the points, and one all_gather of (1 + k) doubles per point assembles the result.  There is no
data-path collective: the Gram/Cholesky kernels never talk across GPUs here.
"""

import numpy
import torch
import torch.distributed as dist

__all__ = ['shard_indices', 'eval_batch_sharded']


def shard_indices(nitems, rank, world):
    """ indices of the items evaluated by `rank`: round-robin, so that any prefix of the batch is balanced """
    return list(range(rank, nitems, world))


def eval_batch_sharded(fun, thetas, *, group=None, device=None):
    """Evaluate ``fun(theta) -> 1-d array of fixed length`` on every row of `thetas`, sharded over the ranks of
    the process group; every rank returns the full (B, len) array.

    Works without an initialised process group (single process).  With the NCCL backend pass the rank's CUDA
    device as `device`; with gloo leave it None (CPU tensors)."""
    thetas = numpy.asarray(thetas, dtype=float)
    B = thetas.shape[0]
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    mine = shard_indices(B, rank, world)
    local = [numpy.asarray(fun(thetas[i]), dtype=float).reshape(-1) for i in mine]
    width = None
    if local:
        width = local[0].size
    if world == 1:
        out = numpy.empty((B, width or 0))
        for i, v in zip(mine, local):
            out[i] = v
        return out
    # all ranks must agree on the row width even if a rank has no items
    wt = torch.tensor([width or 0], dtype=torch.int64, device=device)
    dist.all_reduce(wt, op=dist.ReduceOp.MAX, group=group)
    width = int(wt.item())
    per = (B + world - 1) // world
    buf = torch.full((per, width), float('nan'), dtype=torch.float64, device=device)
    for j, v in enumerate(local):
        buf[j] = torch.as_tensor(v, dtype=torch.float64, device=device)
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf, group=group)
    out = numpy.empty((B, width))
    for r in range(world):
        idx = shard_indices(B, r, world)
        out[idx] = gathered[r][:len(idx)].cpu().numpy()
    return out
