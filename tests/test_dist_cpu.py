"""Multi-rank logic on CPU (gloo, world_size 2): genome sharding, the rendezvous of the in-library communicator
(the 128-byte id travels over torch.distributed) and the algebra of the exchanges.  The per-rank compute is stood
in for by the oracle; the tests check what the multi-GPU path relies on: shared(all genomes) = AND over ranks of
shared(genome 0 + the rank's shard), picks = OR over ranks, and genome 0's packed rows SUM to the rows of a
single pass."""
import os
import socket

import numpy as np
import pytest

from primerforge_b200.dist import shard_genomes


def test_shard_genomes_partition():
    for n in (1, 2, 5, 10, 200):
        for world in (1, 2, 3, 8):
            seen = []
            for rank in range(world):
                part = shard_genomes(n, world, rank)
                assert part[0] == 0
                seen += part[1:]
            assert sorted(seen) == list(range(1, n))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from oracle import pf_oracle as po
    from primerforge_b200 import synth
    from primerforge_b200.dist import shard_genomes

    def reduce_bytes(t, op):          # what the library does with NCCL on the device buffers
        dist.all_reduce(t, op=dist.ReduceOp.MIN if op == "min" else dist.ReduceOp.MAX)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gens = synth.make_genomes(5, 15000, seed=77, snp_rate=0.004, n_inversions=2, n_contigs=2)
    par = po.default_params(table_size=65521)
    universe = po.run(gens[:1], par)                      # genome 0 alone defines the key universe
    keys = {(int(w), int(k)): i for i, (w, k) in enumerate(zip(universe.word, universe.k))}
    mine = [gens[i] for i in shard_genomes(len(gens), world, rank)]
    local = po.run(mine, par)
    alive = np.zeros(len(keys), dtype=np.uint8)
    picked = np.zeros(len(keys), dtype=np.uint8)
    for w, k, f in zip(local.word, local.k, local.flags):
        alive[keys[(int(w), int(k))]] = 1
    t = torch.from_numpy(alive)
    reduce_bytes(t, "min")
    # picks: per rank on the globally alive set (what stage 1 does), then MAX
    for w, k, f in zip(local.word, local.k, local.flags):
        i = keys[(int(w), int(k))]
        if alive[i] and (f & 8):
            picked[i] = 1
    tp = torch.from_numpy(picked)
    reduce_bytes(tp, "max")
    if rank == 0:
        full = po.run(gens, par)
        want = np.zeros(len(keys), dtype=np.uint8)
        for w, k in zip(full.word, full.k):
            want[keys[(int(w), int(k))]] = 1
        np.save(os.path.join(out_dir, "ok.npy"), np.array([int(np.array_equal(want, alive)), int(want.sum()), int(alive.sum())]))
    dist.destroy_process_group()


def test_two_rank_alive_reduction_equals_single_run(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ok, n_want, n_got = np.load(os.path.join(str(tmp_path), "ok.npy"))
    assert ok == 1 and n_want == n_got and n_want > 0


class _FakeEngine:
    """stands in for primerforge_b200.engine.Engine where no GPU exists: records what attach_comm hands it"""
    made = 0

    def __init__(self):
        self.comm = None
        self.options = {}

    @staticmethod
    def unique_id() -> bytes:
        _FakeEngine.made += 1
        return bytes((7 * i + os.getpid()) % 256 for i in range(128))

    def comm_init(self, n_ranks, rank, uid):
        self.comm = (n_ranks, rank, bytes(uid))

    def set_option(self, option, value):
        self.options[option] = value


def _worker_comm(rank, world, port, out_dir):
    import torch.distributed as dist
    from primerforge_b200 import _lib
    from primerforge_b200.dist import attach_comm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = _FakeEngine()
    assert attach_comm(eng) is True
    n_ranks, r, uid = eng.comm
    assert (n_ranks, r) == (world, rank) and len(uid) == 128
    assert _FakeEngine.made == (1 if rank == 0 else 0)                 # only rank 0 makes the id
    assert eng.options[_lib.OPT_FETCH_RECORDS] == (1 if rank == 0 else 0)
    with open(os.path.join(out_dir, f"uid{rank}.bin"), "wb") as fh:
        fh.write(uid)
    dist.destroy_process_group()


def test_every_rank_joins_with_rank_zeros_id(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker_comm, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = open(os.path.join(str(tmp_path), "uid0.bin"), "rb").read()
    b = open(os.path.join(str(tmp_path), "uid1.bin"), "rb").read()
    assert a == b and len(a) == 128


def test_attach_comm_is_a_no_op_without_a_process_group():
    from primerforge_b200.dist import attach_comm
    eng = _FakeEngine()
    assert attach_comm(eng) is False and eng.comm is None


def test_packed_rows_sum_to_the_single_pass_rows():
    """the 32-bit row format of the shared-out genome-1 pass (k_rows0_pack / k_rows0_unpack): count saturated at 3
    in bits 26.., start of a lone window below; the plain SUM over ranks must give 'exactly one window, and
    where' whenever a single pass would, and 'two or more' otherwise -- for every way the windows of a bin can
    fall on 8 ranks"""
    rng = np.random.default_rng(5)
    mask = (1 << 26) - 1
    for _ in range(20000):
        world = int(rng.integers(2, 9))
        n_win = int(rng.integers(0, 6))
        starts = rng.integers(0, 1 << 26, size=n_win)
        owner = rng.integers(0, world, size=n_win)
        total = 0
        for r in range(world):
            mine = starts[owner == r]
            c = min(3, mine.size)
            total += (c << 26) | (int(mine[0]) if c == 1 else 0)
        assert total < (1 << 32)
        c = total >> 26
        if n_win == 0:
            assert total == 0
        elif n_win == 1:
            assert c == 1 and (total & mask) == int(starts[0])
        else:
            assert c >= 2
