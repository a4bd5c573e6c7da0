"""Generate tests/golden/pp_*.npz by running the reference's UNMODIFIED bin/getPrimerPairs.py::_getPrimerPairs (from
/root/reference, present only in the build container) on the candidates its own unmodified
bin/getCandidateKmers.py finds in synthetic ingroup genomes (stand-in dependencies: primer3.calc_heterodimer_tm is
0.0, so no pair is rejected as a dimer).  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_pairs.py            # rewrites every fixture

Each fixture stores the candidate dict as columns -- the S candidate strings with their Tm, and per genome the
candidates in the order of its dict of lists (candidate index, contig index in dict order, leftmost coordinate,
strand) -- the pair-search parameters, and the reference's result: the pairs in dict order (forward / reverse
primer strings) with every field of their Product in every genome.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from primerforge_b200 import synth  # noqa: E402
from oracle.run_reference import run_reference, run_reference_pairs  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (ingroup genome kwargs, stage kwargs, pair-search kwargs)
    "pp_basic": (dict(n_genomes=3, length=12000, seed=401, snp_rate=0.002, n_inversions=2), {},
                 dict(min_pcr=120, max_pcr=2400, max_tm_diff=5.0, max_bin_size=64)),
    "pp_contigs": (dict(n_genomes=4, length=16000, seed=402, snp_rate=0.003, n_inversions=3, n_contigs=4), {},
                   dict(min_pcr=100, max_pcr=1200, max_tm_diff=2.0, max_bin_size=40)),
    "pp_widek": (dict(n_genomes=3, length=10000, seed=403, snp_rate=0.002, n_inversions=2, n_contigs=2),
                 dict(min_len=12, max_len=30, max_repeats=3), dict(min_len=12, max_len=30, min_pcr=80, max_pcr=900, max_tm_diff=8.0, max_bin_size=30)),
    "pp_tightbins": (dict(n_genomes=2, length=9000, seed=404, snp_rate=0.001, n_inversions=1), dict(min_gc=30.0, max_gc=70.0, min_tm=50.0, max_tm=72.0),
                     dict(min_pcr=60, max_pcr=300, max_tm_diff=1.0, max_bin_size=22)),
}


def make_one(name):
    ikw, skw, pkw = CASES[name]
    gens = synth.make_genomes(**ikw)
    with tempfile.TemporaryDirectory() as d:
        paths = synth.write_fasta(gens, d)
        os.environ.pop("PF_STANDIN_TABLE", None)
        _, cands = run_reference(paths, thal_mode="pass", **skw)
        # the candidate dict as columns, BEFORE the pair search touches the objects
        first = gens[0].name
        seqs, tm = [], []
        for plist in cands[first].values():
            for p in plist:
                seqs.append(str(p.seq)); tm.append(p.Tm)
        index = {s: i for i, s in enumerate(seqs)}
        assert len(index) == len(seqs)
        cols = {}
        for gi, g in enumerate(gens):
            ids, ct, lf, st = [], [], [], []
            for ci, (cname, plist) in enumerate(cands[g.name].items()):
                for p in plist:
                    ids.append(index[str(p.seq)]); ct.append(ci); lf.append(min(p.start, p.end)); st.append(int(p.strand == "-"))
            assert sorted(ids) == list(range(len(seqs)))
            cols[gi] = (ids, ct, lf, st, list(cands[g.name].keys()))
        pairs = run_reference_pairs(cands, paths, **pkw)
    names = [g.name for g in gens]
    rows, prods = [], []
    for (f, r), per in pairs.items():
        rows.append((str(f.seq), str(r.seq)))
        for gi, nm in enumerate(names):
            pr = per[nm]
            cidx = cols[gi][4].index(pr.contig)
            prods.append((cidx, int(pr.size), pr.fbin, pr.rbin, pr.fStart, pr.fEnd, int(pr.fStrand == "-"), pr.rStart, pr.rEnd,
                          int(pr.rStrand == "-")))
            assert pr.dimerTm == 0.0
    out = dict(genome_names=np.array(names), cand=np.array(seqs, dtype="S32"), cand_tm=np.array(tm, dtype=np.float64),
               params=np.array([pkw.get("max_bin_size", 64), pkw.get("min_len", 16), pkw["min_pcr"], pkw["max_pcr"]], dtype=np.int64),
               max_tm_diff=np.float64(pkw["max_tm_diff"]),
               pair_fwd=np.array([a for a, _ in rows], dtype="S32"), pair_rev=np.array([b for _, b in rows], dtype="S32"),
               products=np.array(prods, dtype=np.int64).reshape(len(rows), len(names), 10))
    for gi in range(len(gens)):
        ids, ct, lf, st, cnames = cols[gi]
        out[f"g{gi}_id"] = np.array(ids, dtype=np.uint32); out[f"g{gi}_contig"] = np.array(ct, dtype=np.uint32)
        out[f"g{gi}_left"] = np.array(lf, dtype=np.uint32); out[f"g{gi}_strand"] = np.array(st, dtype=np.uint8)
        out[f"g{gi}_contig_names"] = np.array(cnames)
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {len(seqs)} candidates, {len(rows)} pairs -> {os.path.getsize(path) // 1024} KiB")


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(CASES)):
        make_one(nm)
