"""Multi-rank logic on CPU (gloo, world_size 2): genome sharding + the two byte-vector reductions.
The per-rank compute is stood in for by the oracle; the test checks the algebra the multi-GPU path
relies on: shared(all genomes) = AND over ranks of shared(genome 0 + the rank's shard)."""
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
    from primerforge_b200.dist import reduce_bytes, shard_genomes
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
