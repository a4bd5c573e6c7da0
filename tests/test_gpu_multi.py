import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("xchg", ["peer", "nccl"])
@pytest.mark.parametrize("n_contigs,ingest", [(1, "bases"), (5, "bases"), (3, "fasta")])
def test_sharded_equals_single_gpu(n_contigs, ingest, xchg):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "mgpu_worker.py"), "7", "300000", str(n_contigs), ingest]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, PFB200_XCHG=xchg))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "ok=True" in out.stdout


@pytest.mark.parametrize("n_contigs,ingest", [(1, "bases"), (4, "fasta")])
def test_one_process_for_all_gpus_equals_single_gpu(n_contigs, ingest):
    """MultiEngine (pfb_multi_*): one host thread, N contexts, NCCL inside the library"""
    import numpy as np
    import torch
    from primerforge_b200 import synth
    from primerforge_b200.engine import Engine, MultiEngine
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    gens = synth.make_genomes(7, 300_000, seed=911, snp_rate=0.004, n_inversions=3, n_contigs=n_contigs)

    def image(g):
        img = bytearray()
        for cid, arr in zip(g.contig_ids, g.contigs):
            img += b">" + cid.encode() + b"\n"
            raw = arr.tobytes()
            for i in range(0, len(raw), 70):
                img += raw[i:i + 70] + b"\n"
        return np.frombuffer(bytes(img), dtype=np.uint8).copy()

    with Engine(device=0) as ref:
        for g in gens:
            ref.add_genome(g.contigs)
        ref.run()
        r_rec, r_ss, r_ct = ref.result_picked()
        r_rec, r_ss = r_rec.copy(), r_ss.copy()
        r_ct = None if r_ct is None else r_ct.copy()
    with MultiEngine(list(range(world))) as me:
        for g in gens:
            if ingest == "fasta":
                me.add_fasta(image(g))
            else:
                me.add_genome(g.contigs)
        me.run()
        assert me.engines[0].timing()["sharded_first_genome"] == 1
        assert me.engines[0].timing()["exchange_mode"] == 2          # peer access between the contexts of one process
        for gi in range(len(gens)):
            ei, li = me.where(gi)
            rec, ss, ct = me.engines[ei].result_picked()
            assert np.array_equal(rec["word"], r_rec["word"]) and np.array_equal(rec["tm"], r_rec["tm"])
            assert np.array_equal(ss[li], r_ss[gi])
            if r_ct is not None:
                assert np.array_equal(ct[li], r_ct[gi])
