"""helpers shared by the CPU and GPU test suites"""
from __future__ import annotations

import glob
import os
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_cases() -> list:
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_golden(name: str) -> SimpleNamespace:
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    from primerforge_b200.synth import Genome
    names = [str(x) for x in z["genome_names"]]
    gens = []
    for gi, nm in enumerate(names):
        bases = z[f"g{gi}_bases"]
        off = z[f"g{gi}_offsets"]
        contigs = [np.ascontiguousarray(bases[off[i]:off[i + 1]]) for i in range(off.size - 1)]
        gens.append(Genome(name=nm, contig_ids=[str(x) for x in z[f"g{gi}_contig_ids"]], contigs=contigs))
    k_min, k_max, max_repeat = (int(x) for x in z["params"])
    min_gc, max_gc, min_tm, max_tm, tol = (float(x) for x in z["fparams"])
    return SimpleNamespace(
        z=z, genomes=gens, names=names, k_min=k_min, k_max=k_max, max_repeat=max_repeat, min_gc=min_gc,
        max_gc=max_gc, min_tm=min_tm, max_tm=max_tm, temp_tolerance=tol, table_size=int(z["table_size"]),
        thal=str(z["thal"]), shared=[s.decode() for s in z["shared_kmers"]])


def golden_candidate_set(gold, gi: int) -> set:
    z = gold.z
    km = z[f"g{gi}_cand_kmer"]
    return {(gold.shared[int(km[i])], int(z[f"g{gi}_cand_contig"][i]), int(z[f"g{gi}_cand_start"][i]),
             int(z[f"g{gi}_cand_end"][i]), int(z[f"g{gi}_cand_strand"][i])) for i in range(km.size)}


def candidate_set_from_arrays(kmers, k, picked, contig, start, strand) -> set:
    """(kmer, contig idx, Primer.start, Primer.end, strand) with the Primer.py:42-45 start/end flip"""
    out = set()
    for r in np.flatnonzero(picked):
        kk = int(k[r])
        s = int(start[r])
        if strand[r] == 0:
            out.add((kmers[r], int(contig[r]), s, s + kk - 1, 0))
        else:
            out.add((kmers[r], int(contig[r]), s + kk - 1, s, 1))
    return out
