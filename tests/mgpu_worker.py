"""torchrun worker of tests/test_gpu_multi.py: the genome-sharded path on N GPUs must reproduce the
single-GPU result bit for bit (run on rank 0's GPU for comparison)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from primerforge_b200 import synth
    from primerforge_b200.dist import attach_comm, shard_genomes
    from primerforge_b200.engine import Engine
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_genomes, length = int(sys.argv[1]), int(sys.argv[2])
    n_contigs = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    as_fasta = len(sys.argv) > 4 and sys.argv[4] == "fasta"       # hand the ranks FASTA file images (device ingest)
    gens = synth.make_genomes(n_genomes, length, seed=909, snp_rate=0.004, n_inversions=3, n_contigs=n_contigs)
    mine = shard_genomes(n_genomes, world, rank)
    eng = Engine(device=local)
    attach_comm(eng)          # from here on eng.run() is a collective call: the exchanges happen inside the library
    for gi in mine:
        if as_fasta:
            img = bytearray()
            for cid, arr in zip(gens[gi].contig_ids, gens[gi].contigs):
                img += b">" + cid.encode() + b"\n"
                raw = arr.tobytes()
                for i in range(0, len(raw), 70):
                    img += raw[i:i + 70] + b"\n"
            eng.add_fasta(np.frombuffer(bytes(img), dtype=np.uint8).copy())
        else:
            eng.add_genome(gens[gi].contigs)
    eng.run()
    tmg = eng.timing()
    assert tmg["sharded_first_genome"] == 1               # the ranks shared genome 1's own pipeline
    want_mode = 1 if os.environ.get("PFB200_XCHG", "").lower().startswith("n") else 2
    assert tmg["exchange_mode"] == want_mode, tmg["exchange_mode"]     # 2: through each other's arenas over NVLink, 1: NCCL
    rec, ss, ct = eng.result_picked()
    rec, ss = rec.copy(), ss.copy()
    ct = None if ct is None else ct.copy()
    # gather every rank's coordinates on rank 0 (test only)
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, rec["word"].tolist(), ss, ct))
    ok = True
    if rank == 0:
        ref = Engine(device=local)
        for g in gens:
            ref.add_genome(g.contigs)
        ref.run()
        r_rec, r_ss, r_ct = ref.result_picked()
        for owner, words, s2, c2 in gathered:
            ok &= words == r_rec["word"].tolist()
            for j, gi in enumerate(owner):
                ok &= bool(np.array_equal(s2[j], r_ss[gi]))
                if r_ct is not None:
                    ok &= bool(np.array_equal(c2[j], r_ct[gi]))
        print(f"MGPU world={world} genomes={n_genomes} records={r_rec.size} ok={ok}")
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
