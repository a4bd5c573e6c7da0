"""Generate tests/golden/og_*.npz by running the reference's UNMODIFIED bin/removeOutgroupPrimers.py (from
/root/reference, present only in the build container) against the stand-in dependencies.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_outgroup.py            # rewrites every fixture

The primer pairs come from the reference's own upstream stages run on synthetic ingroup genomes
(bin/getCandidateKmers.py -> bin/getPrimerPairs.py, both unmodified).  The outgroup genomes are diverged copies of
the ingroup ancestor -- so that some primers still bind -- with planted duplications (several products per pair),
reverse-complemented contigs, N bases, a lower-case stretch and contigs shorter than a primer.  Each fixture stores
the pairs (primer strings), the outgroup genomes, params.disallowedLens, and what the reference did: which pairs it
kept, the (contig, size) sets of its Product objects per outgroup genome, its debug log lines, and
__extractKmerData's start lists for every contig and strand.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from primerforge_b200 import synth  # noqa: E402
from oracle.run_reference import (reference_extract_kmer_data, run_reference, run_reference_outgroup,  # noqa: E402
                                  run_reference_pairs)

GOLDEN = os.path.join(ROOT, "tests", "golden")
_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)


def _diverge(anc: np.ndarray, rate: float, rng) -> np.ndarray:
    out = anc.copy()
    sites = np.flatnonzero(rng.random(anc.size) < rate)
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    cur = np.array([code[int(b)] for b in out[sites]], dtype=np.int64)
    out[sites] = _ASCII[(cur + rng.integers(1, 4, size=sites.size)) & 3]
    return out


def _outgroup(anc, seed, rates, n_contigs=1, dup=0, n_frac=0.0, lower=False, tiny=False):
    rng = np.random.default_rng(seed)
    gens = []
    for g, rate in enumerate(rates):
        seq = _diverge(anc, rate, rng)
        for _ in range(dup):                       # a copied segment: the primers inside it bind twice
            n = int(rng.integers(200, 900))
            src = int(rng.integers(0, seq.size - n))
            dst = int(rng.integers(0, seq.size - n))
            seq[dst:dst + n] = seq[src:src + n].copy()
        if n_frac:
            seq[rng.random(seq.size) < n_frac] = ord("N")
        if lower:
            a = int(rng.integers(0, seq.size - 400))
            seq[a:a + 300] = np.frombuffer(seq[a:a + 300].tobytes().lower(), dtype=np.uint8)
        if n_contigs > 1:
            cuts = np.sort(rng.choice(np.arange(1, seq.size), size=n_contigs - 1, replace=False))
            pieces = [np.ascontiguousarray(p) for p in np.split(seq, cuts)]
            pieces = [synth.revcomp_ascii(p) if (i % 3 == 1) else p for i, p in enumerate(pieces)]
            pieces = [pieces[i] for i in rng.permutation(len(pieces))]
        else:
            pieces = [seq]
        if tiny:
            pieces.insert(1, np.frombuffer(b"ACGTACG", dtype=np.uint8).copy())
            pieces.append(np.frombuffer(b"GGC", dtype=np.uint8).copy())
        gens.append(synth.Genome(name=f"outgroup_{g:03d}.fasta", contig_ids=[f"o{g:03d}_c{c:03d}" for c in range(len(pieces))],
                                 contigs=pieces))
    return gens


CASES = {
    # name: (ingroup kwargs, stage kwargs, pair kwargs, outgroup kwargs, disallowed range)
    "og_basic": (dict(n_genomes=3, length=12000, seed=301, snp_rate=0.002, n_inversions=2), {},
                 dict(min_pcr=120, max_pcr=2400), dict(seed=11, rates=[0.02, 0.04, 0.08]), (120, 2400)),
    "og_multi": (dict(n_genomes=3, length=14000, seed=302, snp_rate=0.002, n_inversions=2, n_contigs=3), {},
                 dict(min_pcr=120, max_pcr=2400),
                 dict(seed=12, rates=[0.003, 0.01, 0.02, 0.06], n_contigs=5, dup=6, n_frac=0.0008, lower=True, tiny=True), (400, 1500)),
    "og_widek": (dict(n_genomes=3, length=9000, seed=303, snp_rate=0.002, n_inversions=2, n_contigs=2),
                 dict(min_len=12, max_len=30, max_repeats=3), dict(min_len=12, max_len=30, min_pcr=100, max_pcr=1500),
                 dict(seed=13, rates=[0.005, 0.03], n_contigs=3, dup=4), (600, 1500)),
    "og_allgone": (dict(n_genomes=2, length=6000, seed=304, snp_rate=0.001, n_inversions=0), {},
                   dict(min_pcr=120, max_pcr=2400), dict(seed=14, rates=[0.0]), (1, 100000)),
}


def _product_set(prod, contig_index):
    """the (contig, size) set behind a Product made by __processOutgroupResults (:125-178)"""
    if prod.contig == "NA" and prod.size == 0:
        return []
    if isinstance(prod.size, str):
        return sorted(zip((contig_index[c] for c in prod.contig.split(",")), (int(s) for s in prod.size.split(","))))
    return [(contig_index[prod.contig], int(prod.size))]


def make_one(name):
    ikw, skw, pkw, okw, bad = CASES[name]
    gens = synth.make_genomes(**ikw)
    with tempfile.TemporaryDirectory() as d:
        paths = synth.write_fasta(gens, d)
        os.environ.pop("PF_STANDIN_TABLE", None)
        _, cands = run_reference(paths, thal_mode="pass", **skw)
        pairs = run_reference_pairs(cands, paths, **pkw)
        anc = np.concatenate(gens[0].contigs)
        outs = _outgroup(anc, **okw)
        opaths = synth.write_fasta(outs, os.path.join(d, "out"))
        pair_list = list(pairs.keys())
        ingroup_names = [g.name for g in gens]
        err = ""
        try:
            debug = run_reference_outgroup(opaths, pairs, range(bad[0], bad[1] + 1))
        except RuntimeError as e:
            err, debug = str(e), []
    fwd = [str(f) for f, _ in pair_list]
    rev = [str(r) for _, r in pair_list]
    kept = np.array([p in pairs for p in pair_list], dtype=np.uint8)
    rows = []
    if not err:
        for pi, p in enumerate(pair_list):
            if p not in pairs:
                continue
            assert set(pairs[p].keys()) == set(ingroup_names) | {o.name for o in outs}
            for gi, o in enumerate(outs):
                cidx = {c: i for i, c in enumerate(o.contig_ids)}
                pr = pairs[p][o.name]
                assert pr.dimerTm == pairs[p][ingroup_names[0]].dimerTm and pr.fbin == -1 and pr.fStrand == ""
                rows += [(pi, gi, c, s) for c, s in _product_set(pr, cidx)]
    keys = list(dict.fromkeys(fwd + rev))
    hits = []
    for gi, o in enumerate(outs):
        for ci, arr in enumerate(o.contigs):
            s = arr.tobytes().decode("ascii")
            rc = synth.revcomp_ascii(arr).tobytes().decode("ascii")
            for strand, text in (("+", s), ("-", rc)):
                got = reference_extract_kmer_data(text, strand, keys)
                for kmer, starts in got.items():
                    hits += [(keys.index(kmer), gi, ci, st, int(strand == "-")) for st in starts]
    out = dict(fwd=np.array(fwd, dtype="S32"), rev=np.array(rev, dtype="S32"), keys=np.array(keys, dtype="S32"),
               disallowed=np.array(bad, dtype=np.int64), kept=kept, error=np.array(err),
               products=np.array(sorted(set(rows)), dtype=np.int64).reshape(-1, 4),
               hits=np.array(sorted(hits), dtype=np.int64).reshape(-1, 5), debug=np.array(debug),
               outgroup_names=np.array([o.name for o in outs]))
    for gi, o in enumerate(outs):
        out[f"o{gi}_contig_ids"] = np.array(o.contig_ids)
        out[f"o{gi}_bases"] = np.concatenate(o.contigs)
        out[f"o{gi}_offsets"] = np.cumsum([0] + [c.size for c in o.contigs]).astype(np.int64)
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, f"{name}.npz")
    np.savez_compressed(path, **out)
    multi = sum(1 for r in set((a, b) for a, b, _, _ in rows) if sum(1 for x in rows if x[:2] == r) > 1)
    print(f"{name}: {len(pair_list)} pairs, {int(kept.sum())} kept, {len(rows)} products ({multi} (pair, genome) with several), "
          f"{len(hits)} hits, error={err!r} -> {os.path.getsize(path) // 1024} KiB")


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(CASES)):
        make_one(nm)
