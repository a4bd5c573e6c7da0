"""small end-to-end runs for compute-sanitizer (memcheck / racecheck): both scan kernels, multi-contig,
FASTA ingest (sync and pipelined), the picked-output path"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primerforge_b200 import synth
from primerforge_b200.engine import Engine

gens = synth.make_genomes(3, 30_000, seed=11, snp_rate=0.003, n_inversions=2, n_contigs=3, n_frac=0.0005)
for prefilter in (0, 1):
    for mode in (1, 2):
        with Engine(device=0, prefilter=prefilter, k_min=14, k_max=22) as eng:
            eng.set_scan_mode(mode)
            for g in gens:
                eng.add_genome(g.contigs)
            eng.run()
            rec, ss, ct = eng.result_picked()
            print("prefilter", prefilter, "mode", mode, "records", eng.counts(), rec.size)

def img(g):
    out = bytearray()
    for cid, arr in zip(g.contig_ids, g.contigs):
        out += b">" + cid.encode() + b" x\n"
        s = arr.tobytes()
        for i in range(0, len(s), 70):
            out += s[i:i + 70] + b"\n"
    return np.frombuffer(bytes(out), dtype=np.uint8).copy()

for piped in (False, True):
    with Engine(device=0) as eng:
        for g in gens:
            eng.add_fasta(img(g))
        if piped:
            eng.run()
        else:
            eng.upload(); eng.compute()
        print("fasta piped", piped, eng.counts(), [eng.genome_info(i) for i in range(3)])
print("done")
