"""helpers shared by the CPU and GPU test suites"""
from __future__ import annotations

import glob
import os
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_cases() -> list:
    """fixtures of the candidate-k-mer stage (og_* = outgroup screen, pp_* = pair search)"""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
    return [n for n in names if not n.startswith(("og_", "pp_"))]


def og_cases() -> list:
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "og_*.npz")))


def load_og(name: str) -> SimpleNamespace:
    """an outgroup-screen fixture (oracle/make_golden_outgroup.py): the pairs, the outgroup genomes and what the
    reference's unmodified bin/removeOutgroupPrimers.py did with them"""
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    from primerforge_b200.synth import Genome
    names = [str(x) for x in z["outgroup_names"]]
    gens = []
    for gi, nm in enumerate(names):
        bases, off = z[f"o{gi}_bases"], z[f"o{gi}_offsets"]
        gens.append(Genome(name=nm, contig_ids=[str(x) for x in z[f"o{gi}_contig_ids"]],
                           contigs=[np.ascontiguousarray(bases[off[i]:off[i + 1]]) for i in range(off.size - 1)]))
    keys = [s.decode() for s in z["keys"]]
    index = {k: i for i, k in enumerate(keys)}
    fwd = [s.decode() for s in z["fwd"]]
    rev = [s.decode() for s in z["rev"]]
    return SimpleNamespace(z=z, genomes=gens, names=names, keys=keys, fwd=fwd, rev=rev,
                           pair_fwd=np.array([index[s] for s in fwd], dtype=np.uint32),
                           pair_rev=np.array([index[s] for s in rev], dtype=np.uint32),
                           bad_lo=int(z["disallowed"][0]), bad_hi=int(z["disallowed"][1]), kept=z["kept"].astype(bool),
                           products=z["products"], hits=z["hits"], error=str(z["error"]), debug=[str(x) for x in z["debug"]])


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


def assert_equals_oracle(res, o, prefilter: bool, n_genomes: int):
    """bit-for-bit comparison of a StageResult with an OracleResult of the same inputs: words, order, lengths,
    flags (GC / Tm / repeat / pick), GC counts, Tm (identical doubles; the contract is 0.01), and contig / leftmost
    start / strand in every genome.  With prefilter the GPU reports only the k-mers that pass the three filters."""
    rec = res.records
    keep = ((o.flags & 7) == 7) if prefilter else np.ones(o.n_shared, dtype=bool)
    assert rec.size == int(keep.sum()), f"{rec.size} records, oracle has {int(keep.sum())}"
    assert np.array_equal(rec["word"], o.word[keep]), "k-mer set / dict order differs"
    assert np.array_equal(rec["k"], o.k[keep])
    assert np.array_equal(rec["flags"] & 15, o.flags[keep]), "filter flags / picks differ"
    assert np.array_equal(rec["gc_count"].astype(np.int32), o.gc_count[keep])
    assert np.max(np.abs(rec["tm"] - o.tm[keep]), initial=0.0) <= 0.01
    assert np.array_equal(rec["tm"], o.tm[keep]), "Tm doubles differ"
    for gi in range(n_genomes):
        pos = res.positions[gi]
        assert np.array_equal(pos["contig"].astype(np.int32), o.contig[gi][keep]), f"contig differs in genome {gi}"
        assert np.array_equal(pos["start"].astype(np.int64), o.start[gi][keep]), f"start differs in genome {gi}"
        assert np.array_equal(pos["strand"].astype(np.uint8), o.strand[gi][keep]), f"strand differs in genome {gi}"


def pp_cases() -> list:
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "pp_*.npz")))


_CODE = {65: 0, 84: 1, 67: 2, 71: 3}


def str_to_word(s) -> int:
    w = 0
    for b in (s if isinstance(s, bytes) else s.encode()):
        w = (w << 2) | _CODE[b]
    return w


def load_pp(name: str) -> SimpleNamespace:
    """a pair-search fixture (oracle/make_golden_pairs.py): the candidate dict as columns, the parameters, and the
    pairs + Products the reference's unmodified bin/getPrimerPairs.py made of them"""
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    cand = [s.decode() for s in z["cand"]]
    G = len(z["genome_names"])
    genomes = [(z[f"g{g}_id"], z[f"g{g}_contig"], z[f"g{g}_left"], z[f"g{g}_strand"]) for g in range(G)]
    index = {s: i for i, s in enumerate(cand)}
    rc = str.maketrans("ACGT", "TGCA")
    fwd = np.array([index[s.decode()] for s in z["pair_fwd"]], dtype=np.uint32)
    rev = np.array([index[s.decode()[::-1].translate(rc)] for s in z["pair_rev"]], dtype=np.uint32)   # p2 = rev.reverseComplement()
    max_bin, min_len, min_pcr, max_pcr = (int(x) for x in z["params"])
    return SimpleNamespace(z=z, cand=cand, word=np.array([str_to_word(s) for s in cand], dtype=np.uint64),
                           k=np.array([len(s) for s in cand], dtype=np.uint8), tm=z["cand_tm"], genomes=genomes,
                           names=[str(x) for x in z["genome_names"]],
                           contig_names=[[str(x) for x in z[f"g{g}_contig_names"]] for g in range(G)],
                           max_bin=max_bin, min_len=min_len, min_pcr=min_pcr, max_pcr=max_pcr, max_tm_diff=float(z["max_tm_diff"]),
                           pair_fwd=fwd, pair_rev=rev, products=z["products"])


def expected_products(fix, pairs, shared, kept):
    """the Product fields (getPrimerPairs.py:336, 463, 501-502) of the kept pairs, derived from a pair-search result
    (pairs [n, 6], shared [n, G, 4]) and the fixture's candidate columns: [n_kept, G, 10] like the fixture's array"""
    G = len(fix.genomes)
    S = fix.word.size
    out = np.zeros((int(kept.sum()), G, 10), dtype=np.int64)
    for g in range(G):
        gid, gct, glf, gst = fix.genomes[g]
        ct = np.zeros(S, np.int64); lf = np.zeros(S, np.int64); st = np.zeros(S, np.int64)
        ct[gid] = gct; lf[gid] = glf; st[gid] = gst
        for row, q in enumerate(np.flatnonzero(kept)):
            a, b = int(pairs[q, 0]), int(pairs[q, 1])
            ka, kb = int(fix.k[a]), int(fix.k[b])
            # k1 as it is, k2 reverse complemented (:458): Primer.start is the leftmost base on (+), the rightmost on (-)
            f_start, f_end = (lf[a], lf[a] + ka - 1) if st[a] == 0 else (lf[a] + ka - 1, lf[a])
            r_strand = 1 - st[b]
            r_start, r_end = (lf[b], lf[b] + kb - 1) if r_strand == 0 else (lf[b] + kb - 1, lf[b])
            out[row, g] = (ct[a], shared[q, g, 1], shared[q, g, 2], shared[q, g, 3], f_start, f_end, st[a], r_start, r_end, r_strand)
    return out


def synthetic_candidates(seed, S, L, G, kmin=16, kmax=20, n_contigs=3, nested=0.3):
    """S candidates on a random genome of L bases (many nested: same start or same end, other length), placed in G
    genomes whose blocks are shuffled, flipped and dealt to contigs"""
    rng = np.random.default_rng(seed)
    code = rng.integers(0, 4, L).astype(np.uint64)
    starts, ks = [], []
    while len(starts) < S:
        a = int(rng.integers(0, L - kmax)); k = int(rng.integers(kmin, kmax + 1))
        starts.append(a); ks.append(k)
        if rng.random() < nested and k < kmax:                       # a longer one that contains it
            k2 = int(rng.integers(k + 1, kmax + 1))
            a2 = a if rng.random() < 0.5 else max(0, a + k - k2)
            starts.append(a2); ks.append(k2)
    # distinct (start, k) -> distinct strings with overwhelming probability; drop duplicates of the string
    seen, cand = set(), []
    for a, k in zip(starts, ks):
        w = 0
        for c in code[a:a + k].tolist():
            w = (w << 2) | c
        if (w, k) not in seen:
            seen.add((w, k)); cand.append((a, k, w))
    cand = cand[:S]
    S = len(cand)
    start = np.array([c[0] for c in cand], dtype=np.int64); k = np.array([c[1] for c in cand], dtype=np.uint8)
    word = np.array([c[2] for c in cand], dtype=np.uint64)
    tm = np.round(rng.uniform(55.0, 62.0, S), 3)
    genomes = []
    for g in range(G):
        if g == 0:
            contig = np.minimum(start * n_contigs // L, n_contigs - 1)
            left, strand = start.copy(), np.zeros(S, dtype=np.uint8)
            bounds = np.array([c * L // n_contigs for c in range(n_contigs)], dtype=np.int64)
            left = left - bounds[contig]
            # a candidate must not straddle a contig boundary: those that would are put at the end of their contig
        else:
            nblk = 12
            cuts = np.sort(rng.choice(np.arange(100, L - 100), nblk - 1, replace=False))
            edges = np.r_[0, cuts, L]
            order = rng.permutation(nblk); flip = rng.random(nblk) < 0.4
            blk = np.searchsorted(edges, start, side="right") - 1
            # candidates that straddle a block edge keep their block's transform (coordinates may overlap the next block: harmless)
            new_off = np.zeros(nblk, dtype=np.int64); acc = 0
            for b in order:
                new_off[b] = acc; acc += edges[b + 1] - edges[b]
            rel = start - edges[blk]
            blen = edges[blk + 1] - edges[blk]
            left = np.where(flip[blk], new_off[blk] + blen - rel - k, new_off[blk] + rel)
            left = np.maximum(left, 0)
            strand = flip[blk].astype(np.uint8)
            contig = np.minimum(left * n_contigs // L, n_contigs - 1)
            bounds = np.array([c * L // n_contigs for c in range(n_contigs)], dtype=np.int64)
            left = np.maximum(left - bounds[contig], 0)
        # the genome's lists: per contig sorted by leftmost start; ties in a seeded random order (the reference's is arbitrary)
        tie = rng.permutation(S)
        order_p = np.lexsort((tie, left, contig))
        genomes.append((order_p.astype(np.uint32), contig[order_p].astype(np.uint32), left[order_p].astype(np.uint32), strand[order_p]))
    return word, k, tm, genomes
