"""torch.distributed plumbing for the source-sharded registration (BASELINE config 5).

The C-ABI takes the all-reduce as a callback on raw device buffers; here it is NCCL through
torch.distributed, enqueued on the library's own CUDA stream (wrapped as an ExternalStream) so that the
exchange is ordered with the kernels before and after it without any host synchronisation.
"""
from __future__ import annotations

import numpy as np


class _DevBuf:
    """Zero-copy view of a raw device pointer for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr: int, count: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}


def tensor_from_ptr(ptr: int, count: int, dtype: int, device):
    import torch

    typestr = "<f8" if dtype == 0 else "<i4"
    return torch.as_tensor(_DevBuf(ptr, count, typestr), device=device)


def torch_allreduce(group=None):
    """Returns allreduce(ptr, count, dtype, op, stream) backed by torch.distributed (NCCL)."""
    import torch
    import torch.distributed as dist

    def allreduce(ptr, count, dtype, op, stream):
        if count == 0:
            return 0
        dev = torch.device("cuda", torch.cuda.current_device())
        t = tensor_from_ptr(ptr, count, dtype, dev)
        with torch.cuda.stream(torch.cuda.ExternalStream(stream, device=dev)):
            dist.all_reduce(t, op=dist.ReduceOp.SUM if op == 0 else dist.ReduceOp.MIN, group=group)
        return 0

    return allreduce


def shard_sources(src_clouds, rank: int, world: int):
    """Contiguous index ranges per class (keeps "first in source order" meaningful).
    Returns (shards, index_base, global_n)."""
    shards, base, glob = [], [], []
    for c in src_clouds:
        n = len(c)
        lo = (n * rank) // world
        hi = (n * (rank + 1)) // world
        shards.append(np.ascontiguousarray(c[lo:hi]))
        base.append(lo)
        glob.append(n)
    return shards, base, glob
