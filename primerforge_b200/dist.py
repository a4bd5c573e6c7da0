"""Multi-GPU plumbing, one process per GPU (torchrun / torch.distributed for the rendezvous only).

The stage shards by GENOME: every rank holds genome 0 -- it alone defines the key table and the ids, so both are
identical on every rank -- plus its own share of the other genomes; genome 0's own pipeline (candidate windows,
its scan) is shared out by position.  The exchanges (an all-gather of genome 0's candidate lists, a SUM
all-reduce of its packed rows, a MIN all-reduce of one byte per id and a MAX all-reduce of one byte per id) are
NCCL calls INSIDE libpfb200.so on each context's own stream; torch.distributed only carries the 128-byte NCCL
id from rank 0 to the others.  See DESIGN.md for why the all-to-all of north_star's sketch is not needed once
the key table is replicated.  One process for all GPUs: primerforge_b200.engine.MultiEngine.
"""
from __future__ import annotations


def shard_genomes(n_genomes: int, world: int, rank: int) -> list:
    """global genome indices held by `rank`: genome 0 everywhere, the rest dealt round-robin"""
    if n_genomes <= 0:
        return []
    rest = list(range(1, n_genomes))
    return [0] + rest[rank::world]


def broadcast_unique_id(make_id) -> bytes:
    """rank 0 calls make_id() (-> 128 bytes); every rank returns those bytes (any torch.distributed backend)"""
    import torch.distributed as dist
    rank = dist.get_rank()
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = bytes(box[0])
    if len(uid) != 128:
        raise ValueError("an NCCL unique id is 128 bytes")
    return uid


def attach_comm(eng) -> bool:
    """joins `eng` (this rank's Engine) to a communicator spanning the default process group; afterwards eng.run()
    and eng.compute() are collective calls.  Ranks > 0 leave the per-id record columns on their device (they are
    identical on every rank).  Returns False when there is nothing to attach (one rank)."""
    import torch.distributed as dist
    from . import _lib
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return False
    world, rank = dist.get_world_size(), dist.get_rank()
    uid = broadcast_unique_id(type(eng).unique_id)
    eng.comm_init(world, rank, uid)
    eng.set_option(_lib.OPT_FETCH_RECORDS, 1 if rank == 0 else 0)
    return True
