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


def nccl_unique_id() -> bytes:
    """mulls_nccl_unique_id: 128 bytes to be shipped from rank 0 to every rank (see nccl_init_from_torch)."""
    import ctypes as C

    from . import abi

    buf = C.create_string_buffer(128)
    rc = abi.load_library().mulls_nccl_unique_id(buf)
    if rc != 0:
        raise RuntimeError(f"mulls_nccl_unique_id failed ({rc}): libnccl.so.2 not available")
    return buf.raw


def nccl_init_from_torch(ctx, group=None):
    """Give `ctx` its own NCCL communicator over the ranks of a torch.distributed group: rank 0 creates the id, the
    group's broadcast ships it (any backend), every rank calls mulls_nccl_init."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    uid = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0, group=group)
    ctx.nccl_init(rank, world, uid[0])


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


def run_sharded_local(pair, world: int, device: int, max_src: int, max_tgt: int, want_trace: bool = False):
    """BASELINE config 5's partitioning on ONE device (tests, single-GPU boxes): `world` contexts each hold the
    full target and one contiguous slice of every source class, are driven from `world` host threads, and meet in
    an all-reduce that combines the ranks' device buffers (sum / min) exactly where ncclAllReduce sits in a
    multi-GPU run. Returns ([result dict per rank], [trace dict per rank] or None)."""
    import threading

    import torch

    from .registration import Context

    ctxs = [Context(device, 1, max_src, max_tgt) for _ in range(world)]
    barrier = threading.Barrier(world)
    slots = [None] * world
    dev = torch.device("cuda", device)

    def make_hook(rank):
        def hook(ptr, count, dtype, op, stream):
            torch.cuda.ExternalStream(stream, device=dev).synchronize()
            slots[rank] = tensor_from_ptr(ptr, count, dtype, dev)
            barrier.wait(timeout=120)
            if rank == 0:
                stack = torch.stack([s.clone() for s in slots])
                red = stack.sum(0) if op == 0 else stack.min(0).values
                for s in slots:
                    s.copy_(red)
                torch.cuda.synchronize()
            barrier.wait(timeout=120)
            return 0

        return hook

    out = [None] * world
    err = [None] * world

    def worker(rank):
        try:
            shards, base, glob = shard_sources(pair["src"], rank, world)
            out[rank] = ctxs[rank].run_sharded(dict(pair, src=shards), base, glob, make_hook(rank), want_trace=want_trace)
        except Exception as exc:  # noqa: BLE001
            err[rank] = exc
            barrier.abort()

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    for c in ctxs:
        c.close()
    for e in err:
        if e is not None:
            raise e
    if any(t.is_alive() for t in threads):
        raise RuntimeError("sharded run did not finish")
    return [o[0] for o in out], ([o[1] for o in out] if want_trace else None)
