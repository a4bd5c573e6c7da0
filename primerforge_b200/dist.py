"""Multi-GPU plumbing (one process per GPU, torch.distributed).

The stage shards by GENOME: every rank holds genome 0 -- it alone defines the key table, so the
table is identical on every rank without any exchange -- plus its own share of the other genomes.
What has to be agreed between ranks is tiny: one byte per key ("still unique in every genome this
rank scanned", MIN all-reduce) and one byte per record ("picked in some genome", MAX all-reduce).
No k-mer ever crosses NVLink.  See DESIGN.md for why the all-to-all of north_star's sketch
is not needed once the key table is replicated.
"""
from __future__ import annotations

import numpy as np


def shard_genomes(n_genomes: int, world: int, rank: int) -> list:
    """global genome indices held by `rank`: genome 0 everywhere, the rest dealt round-robin"""
    if n_genomes <= 0:
        return []
    rest = list(range(1, n_genomes))
    return [0] + rest[rank::world]


class _DeviceBytes:
    """exposes a raw device pointer through __cuda_array_interface__ (zero copy into torch)"""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def reduce_bytes(t, op: str):
    """in-place all-reduce of a uint8 tensor over the default process group (any backend)"""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return t
    dist.all_reduce(t, op=dist.ReduceOp.MIN if op == "min" else dist.ReduceOp.MAX)
    return t


def allreduce_stage_buffers(eng, which: int, op: str):
    """all-reduce one of the engine's per-key / per-record byte vectors in place (NCCL on device memory).
    The reduction is enqueued behind the stage's kernels on the engine's own stream (ExternalStream), so
    the host never waits between the stages."""
    import torch
    ptr, n = eng.device_buffer(which)
    if n == 0:
        return
    dev = torch.device("cuda", torch.cuda.current_device())
    ext = getattr(eng, "_torch_stream", None)
    if ext is None:
        ext = torch.cuda.ExternalStream(eng.stream_handle(), device=dev)
        eng._torch_stream = ext
    t = torch.as_tensor(_DeviceBytes(ptr, n), device=dev)
    with torch.cuda.stream(ext):
        reduce_bytes(t, op)


def run_sharded(eng, upload: bool = False):
    """the three compute stages with the two reductions in between; with upload=True the H2D copies of
    the rank's genomes are queued first and overlap stage 0"""
    from . import _lib
    if upload:
        eng.upload_async()
    eng.compute_stage(0)
    allreduce_stage_buffers(eng, _lib.PFB_BUF_ALIVE, "min")
    eng.compute_stage(1)
    allreduce_stage_buffers(eng, _lib.PFB_BUF_PICK, "max")
    eng.compute_stage(2)
    if upload:
        eng.upload_wait()
