"""The CPU restatements against the reference's UNMODIFIED Python run live on seeded random inputs (pair search, outgroup
screen).  Needs /root/reference, which exists only in the build container: skipped elsewhere (the committed fixtures in
tests/golden are what travels).  This is what pins oracle/pf_oracle_pairs.c and oracle/pf_oracle_outgroup.c beyond the
fixtures: many more bins, ties, nested primers and deletions than the fixtures hold."""
import os
import sys

import numpy as np
import pytest

from oracle import pf_oracle as po
from oracle.run_reference import reference_available
from tests.util import ROOT, synthetic_candidates

pytestmark = pytest.mark.skipif(not reference_available(), reason="the reference is only present in the build container")

_LET = "ATCG"


def _str(w, k):
    w, k = int(w), int(k)
    return "".join(_LET[(w >> (2 * (k - 1 - i))) & 3] for i in range(k))


def _reference_modules():
    for pth in ("/root/reference", os.path.join(ROOT, "oracle", "standins")):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    os.environ.setdefault("CI", "1")
    import bin.getPrimerPairs as gpp
    import bin.removeOutgroupPrimers as rop
    from bin.Primer import Primer
    from bin.Product import Product
    from Bio.Seq import Seq
    return gpp, rop, Primer, Product, Seq


@pytest.mark.parametrize("seed,S,L,G,kw", [
    (11, 2500, 12000, 3, dict(max_bin=64, min_primer_len=16, min_pcr=120, max_pcr=2400, max_tm_diff=5.0)),
    (12, 6000, 9000, 2, dict(max_bin=30, min_primer_len=16, min_pcr=60, max_pcr=500, max_tm_diff=2.0)),
    (13, 1500, 20000, 4, dict(max_bin=100, min_primer_len=12, min_pcr=100, max_pcr=4000, max_tm_diff=8.0)),
    (14, 5000, 6000, 2, dict(max_bin=18, min_primer_len=16, min_pcr=30, max_pcr=300, max_tm_diff=1.5)),
])
def test_pair_search_oracle_equals_the_live_reference(seed, S, L, G, kw, capsys):
    from types import SimpleNamespace
    gpp, _rop, Primer, _Product, Seq = _reference_modules()
    kmin, kmax = (12, 30) if seed == 13 else (16, 20)
    word, k, _tm, genomes = synthetic_candidates(seed, S, L, G, kmin, kmax)
    seqs = [_str(w, kk) for w, kk in zip(word, k)]
    names = [f"g{g}.fasta" for g in range(G)]
    cands, tm = {}, np.zeros(len(seqs))
    for g, (gid, gct, glf, gst) in enumerate(genomes):
        per = {}
        for p in range(gid.size):
            s = int(gid[p])
            pr = Primer(Seq(seqs[s]), f"c{int(gct[p])}", int(glf[p]), int(k[s]), "-" if gst[p] else "+")       # getCandidateKmers.py:473-487
            tm[s] = pr.Tm
            per.setdefault(f"c{int(gct[p])}", []).append(pr)
        cands[names[g]] = per
    log = SimpleNamespace(rename=lambda *a: None, info=lambda *a: None, debug=lambda *a: None, error=lambda *a: None)
    prm = SimpleNamespace(ingroupFns=names, maxBinSize=kw["max_bin"], minLen=kw["min_primer_len"], minPcr=kw["min_pcr"], maxPcr=kw["max_pcr"],
                          maxTmDiff=kw["max_tm_diff"], numThreads=1, tempTolerance=5.0, p3_mvConc=50.0, p3_dvConc=1.5, p3_dntpConc=0.6,
                          p3_dnaConc=50.0, p3_tempC=37.0, p3_maxLoop=30, log=log)
    o = po.run_pairs(word, k, tm, genomes, **kw)
    kept = o.kept
    try:
        pairs = gpp._getPrimerPairs(cands, prm)
    except RuntimeError:
        pairs = {}
    capsys.readouterr()
    index = {s: i for i, s in enumerate(seqs)}
    rc = str.maketrans("ACGT", "TGCA")
    got = [(index[str(f.seq)], index[str(r.seq)[::-1].translate(rc)]) for f, r in pairs]
    assert got == [(int(a), int(b)) for a, b in o.pairs[kept, :2]] and len(got) > 0
    for row, per in zip(np.flatnonzero(kept), pairs.values()):
        for g, nm in enumerate(names):
            pr = per[nm]
            assert (pr.size, pr.fbin, pr.rbin) == tuple(int(x) for x in o.shared[row, g, 1:]), (row, g)


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_outgroup_oracle_equals_the_live_reference(seed, capsys):
    from types import SimpleNamespace
    _gpp, rop, _Primer, Product, Seq = _reference_modules()
    from Bio.SeqRecord import SeqRecord
    from primerforge_b200 import synth
    rng = np.random.default_rng(seed)
    L = 30000
    anc = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, L)]
    # primers: windows of the ancestor; pairs: a window and the reverse complement of one downstream (or any two)
    keys, pairs_idx = [], []
    for _ in range(600):
        a = int(rng.integers(0, L - 3000)); ka = int(rng.integers(16, 21)); kb = int(rng.integers(16, 21))
        size = int(rng.integers(60, 2500))
        f = anc[a:a + ka].tobytes().decode()
        r = synth.revcomp_ascii(anc[a + size - kb:a + size]).tobytes().decode()
        for s in (f, r):
            if s not in keys:
                keys.append(s)
        if (keys.index(f), keys.index(r)) not in pairs_idx:
            pairs_idx.append((keys.index(f), keys.index(r)))
    outs = []
    for g, rate in enumerate((0.0, 0.01, 0.03)):
        seq = anc.copy()
        sites = np.flatnonzero(rng.random(L) < rate)
        seq[sites] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, sites.size)]
        for _ in range(5):                                   # copied segments: several products per pair
            n = int(rng.integers(300, 1500)); src = int(rng.integers(0, L - n)); dst = int(rng.integers(0, L - n))
            seq[dst:dst + n] = seq[src:src + n].copy()
        seq[rng.random(L) < 0.0005] = ord("N")
        cuts = np.sort(rng.choice(np.arange(1, L), size=3, replace=False))
        pieces = [synth.revcomp_ascii(p) if i == 2 else np.ascontiguousarray(p) for i, p in enumerate(np.split(seq, cuts))]
        outs.append(synth.Genome(name=f"out{g}", contig_ids=[f"o{g}c{i}" for i in range(4)], contigs=pieces))
    lo, hi = 300, 1200
    o = po.run_outgroup(keys, [a for a, _ in pairs_idx], [b for _, b in pairs_idx], lo, hi, outs)

    class P:                                                  # dict keys with the reference's Primer semantics: str() and hash by sequence
        def __init__(self, s):
            self.s = s

        def __str__(self):
            return self.s

        def __hash__(self):
            return hash(self.s)

        def __eq__(self, other):
            return self.s == other.s
    ref_pairs = {(P(keys[a]), P(keys[b])): {"ingroup": Product("c", 100, 0, 1, 3.5, 0, 1, "+", 2, 3, "-")} for a, b in pairs_idx}
    order = list(ref_pairs.keys())
    outgroup = {g.name: iter([SeqRecord(Seq(c.tobytes().decode()), id=cid) for cid, c in zip(g.contig_ids, g.contigs)]) for g in outs}
    lines = []
    log = SimpleNamespace(rename=lambda *a: None, info=lambda *a: None, error=lambda *a: None, debug=lines.append)
    rop._removeOutgroupPrimers(outgroup, ref_pairs, SimpleNamespace(disallowedLens=range(lo, hi + 1), log=log))
    capsys.readouterr()
    kept = o.removed_at == 0xFFFFFFFF
    assert [pk in ref_pairs for pk in order] == kept.tolist() and kept.any() and not kept.all()
    want = {}
    for p, g, c, size in o.products.tolist():
        want.setdefault((p, g), set()).add((outs[g].contig_ids[c], size))
    for pi, pk in enumerate(order):
        if not kept[pi]:
            continue
        for g, gen in enumerate(outs):
            pr = ref_pairs[pk][gen.name]
            got = set() if (pr.contig, pr.size) == ("NA", 0) else (
                set(zip(pr.contig.split(","), map(int, pr.size.split(",")))) if isinstance(pr.size, str) else {(pr.contig, pr.size)})
            assert got == want.get((pi, g), set()), (pi, g)
    remaining = kept.size
    for g, gen in enumerate(outs):
        gone = int(np.count_nonzero(o.removed_at == g)); remaining -= gone
        assert lines[g] == f"removed {gone} pairs after processing {gen.name} ({remaining} pairs remaining)"


@pytest.mark.parametrize("seed,table,kw", [
    (31, 9999991, dict()),
    (32, 40009, dict(n_contigs=5)),                                   # small sketch: collisions and junction pollution
    (33, 65521, dict(n_contigs=3, n_frac=0.002)),                      # N bases
])
def test_candidate_stage_oracle_equals_the_live_reference(seed, table, kw, tmp_path, monkeypatch):
    """oracle/pf_oracle.c against bin/getCandidateKmers.py on genomes no fixture holds: the shared k-mer dict (keys, order,
    coordinates, strands) and the candidate sets"""
    from oracle.run_reference import run_reference
    from primerforge_b200 import synth
    from tests.util import candidate_set_from_arrays
    gens = synth.make_genomes(3, 9000, seed=seed, snp_rate=0.003, n_inversions=2, **kw)
    paths = synth.write_fasta(gens, str(tmp_path))
    if table != 9999991:
        monkeypatch.setenv("PF_STANDIN_TABLE", str(table))
    else:
        monkeypatch.delenv("PF_STANDIN_TABLE", raising=False)
    shared, cands = run_reference(paths, thal_mode="pass")
    o = po.run(gens, po.default_params(table_size=table))
    kmers = o.kmer_strings()
    assert kmers == list(shared.keys()) and len(kmers) > 100
    for gi, gen in enumerate(gens):
        cidx = {c: i for i, c in enumerate(gen.contig_ids)}
        assert [cidx[shared[s][gen.name][0]] for s in kmers] == o.contig[gi].tolist()
        assert [shared[s][gen.name][1] for s in kmers] == o.start[gi].tolist()
        assert [int(shared[s][gen.name][2] == "-") for s in kmers] == o.strand[gi].tolist()
        mine = candidate_set_from_arrays(kmers, o.k, (o.flags & 8) != 0, o.contig[gi], o.start[gi], o.strand[gi])
        theirs = {(str(p.seq), cidx[c], p.start, p.end, int(p.strand == "-")) for c, lst in cands[gen.name].items() for p in lst}
        assert mine == theirs
