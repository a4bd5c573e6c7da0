"""CUDA pair search (pfb_pp_*, through the C ABI) against the reference's unmodified bin/getPrimerPairs.py (committed
fixtures tests/golden/pp_*.npz) and against the CPU restatement oracle/pf_oracle_pairs.c on seeded inputs."""
import os
from types import SimpleNamespace

import numpy as np
import pytest

from tests.util import expected_products, load_pp, pp_cases, synthetic_candidates as _synthetic

pytestmark = pytest.mark.gpu


def _search(word, k, tm, genomes, **kw):
    from primerforge_b200.pairs import PairSearch
    with PairSearch(device=0) as ps:
        ps.set_candidates(word, k, tm)
        for g in genomes:
            ps.add_genome(*g)
        ps.run(**kw)
        res = ps.result()
        bins = res.bin_of
        reduced = ps.bins(0, with_reduced=True)[1] if genomes and len(word) else None
    return res, bins, reduced


def _as_arrays(res):
    p = res.pairs
    pairs = np.stack([p["fwd"], p["rev"], p["contig"], p["fbin"], p["rbin"], p["pcr_len"]], axis=1).astype(np.uint32).reshape(-1, 6)
    return pairs, res.shared4()


@pytest.mark.parametrize("name", pp_cases())
def test_equals_reference(name):
    f = load_pp(name)
    res, bins, reduced = _search(f.word, f.k, f.tm, f.genomes, max_bin=f.max_bin, min_primer_len=f.min_len, min_pcr=f.min_pcr,
                                 max_pcr=f.max_pcr, max_tm_diff=f.max_tm_diff)
    pairs, shared = _as_arrays(res)
    kept = res.kept
    assert np.array_equal(pairs[kept, 0], f.pair_fwd) and np.array_equal(pairs[kept, 1], f.pair_rev)      # the dict's keys, in order
    assert np.array_equal(expected_products(f, pairs, shared, kept), f.products)                          # every Product field


@pytest.mark.parametrize("seed,S,L,G,kw", [
    (1, 4000, 20000, 3, dict(max_bin=64, min_primer_len=16, min_pcr=120, max_pcr=2400, max_tm_diff=5.0)),
    (2, 20000, 40000, 4, dict(max_bin=40, min_primer_len=16, min_pcr=100, max_pcr=600, max_tm_diff=2.0)),      # dense: long overlap runs
    (3, 3000, 60000, 2, dict(max_bin=25, min_primer_len=12, min_pcr=50, max_pcr=5000, max_tm_diff=10.0)),      # sparse: many bins
    (4, 30000, 30000, 3, dict(max_bin=200, min_primer_len=16, min_pcr=300, max_pcr=900, max_tm_diff=1.0)),     # large bins
    (5, 9000, 10000, 2, dict(max_bin=0, min_primer_len=16, min_pcr=20, max_pcr=200, max_tm_diff=3.0)),         # maxBinSize 0
])
def test_against_oracle(seed, S, L, G, kw):
    from oracle import pf_oracle as po
    kmin, kmax = (12, 30) if seed == 3 else (16, 20)
    word, k, tm, genomes = _synthetic(seed, S, L, G, kmin, kmax)
    o = po.run_pairs(word, k, tm, genomes, **kw)
    res, bins, reduced = _search(word, k, tm, genomes, **kw)
    for g in range(G):
        assert np.array_equal(bins[g], o.bin_of[g]), f"bins of genome {g}"           # __binCandidateKmers
    assert np.array_equal(reduced, o.reduced) and (o.reduced.any() or kw["max_bin"] == 0)     # __reduceBinSize
    pairs, shared = _as_arrays(res)
    assert pairs.shape[0] > 0 and np.array_equal(pairs, o.pairs)                       # bin pairs + first suitable primer pair, in order
    assert np.array_equal(shared, o.shared)                                            # the other genomes
    assert np.array_equal(res.kept, o.kept)


def test_edge_inputs():
    from primerforge_b200.engine import PfbError
    from primerforge_b200.pairs import PairSearch
    e32, e8 = np.zeros(0, np.uint32), np.zeros(0, np.uint8)
    res, bins, _ = _search(np.zeros(0, np.uint64), e8, np.zeros(0), [(e32, e32, e32, e8), (e32, e32, e32, e8)])
    assert res.pairs.size == 0 and res.kept.size == 0
    # one candidate, one genome; two candidates too far apart
    word = np.array([0x1B1B1B1B, 0x2E2E2E2E], np.uint64); k = np.array([16, 16], np.uint8); tm = np.array([60.0, 60.0])
    g0 = (np.array([0, 1], np.uint32), np.zeros(2, np.uint32), np.array([10, 100000], np.uint32), np.zeros(2, np.uint8))
    res, bins, _ = _search(word, k, tm, [g0])
    assert res.pairs.size == 0 and bins[0].tolist() == [0, 1]
    with PairSearch(device=0) as ps:
        ps.set_candidates(word, k, tm)
        with pytest.raises(PfbError):      # a genome must list every candidate
            ps.add_genome(np.array([0], np.uint32), np.zeros(1, np.uint32), np.zeros(1, np.uint32), np.zeros(1, np.uint8))
        with pytest.raises(PfbError):      # the first genome is all (+)
            ps.add_genome(g0[0], g0[1], g0[2], np.array([0, 1], np.uint8))
        ps.add_genome(np.array([0, 0], np.uint32), g0[1], g0[2], g0[3])
        with pytest.raises(PfbError):      # ... exactly once
            ps.run()


class _P:
    """a stand-in for bin.Primer.Primer with what the pair search reads"""
    def __init__(self, seq, contig, left, k, minus, tm):
        self.seq, self.contig, self.strand, self.Tm = seq, contig, "-" if minus else "+", tm
        self.start, self.end = (left + k - 1, left) if minus else (left, left + k - 1)
        self.k = k

    def __str__(self):
        return self.seq

    def reverseComplement(self):
        rc = self.seq[::-1].translate(str.maketrans("ACGT", "TGCA"))
        return _P(rc, self.contig, min(self.start, self.end), self.k, self.strand == "+", self.Tm)


@pytest.mark.parametrize("name", ["pp_contigs", "pp_widek"])
def test_drop_in_entry_point(name):
    """_getPrimerPairs(candidateKmers, params) on a dict of lists of Primer-like objects: the reference's pairs, in
    order, with every Product field"""
    from primerforge_b200.pairs import ERR_NO_PAIRS, _getPrimerPairs
    f = load_pp(name)
    cands = {}
    for g, nm in enumerate(f.names):
        gid, gct, glf, gst = f.genomes[g]
        per = {}
        for p in range(gid.size):
            s = int(gid[p])
            per.setdefault(f.contig_names[g][int(gct[p])], []).append(_P(f.cand[s], f.contig_names[g][int(gct[p])], int(glf[p]), int(f.k[s]), bool(gst[p]), float(f.tm[s])))
        cands[nm] = per
    log = SimpleNamespace(rename=lambda *a: None, info=lambda *a: None, debug=lambda *a: None, error=lambda *a: None)
    prm = SimpleNamespace(ingroupFns=["/some/dir/" + f.names[0]] + f.names[1:], maxBinSize=f.max_bin, minLen=f.min_len, minPcr=f.min_pcr,
                          maxPcr=f.max_pcr, maxTmDiff=f.max_tm_diff, tempTolerance=5.0, log=log)
    out = _getPrimerPairs(cands, prm, dimer="off")
    rc = str.maketrans("ACGT", "TGCA")
    assert [(str(a), str(b)) for a, b in out] == [(f.cand[int(x)], f.cand[int(y)][::-1].translate(rc)) for x, y in zip(f.pair_fwd, f.pair_rev)]
    for row, per in enumerate(out.values()):
        for g, nm in enumerate(f.names):
            pr = per[nm]
            got = (f.contig_names[g].index(pr.contig), pr.size, pr.fbin, pr.rbin, pr.fStart, pr.fEnd, int(pr.fStrand == "-"), pr.rStart,
                   pr.rEnd, int(pr.rStrand == "-"))
            assert got == tuple(int(x) for x in f.products[row, g]) and pr.dimerTm == 0.0
    # a heterodimer check that rejects everything -> the reference's first error
    with pytest.raises(RuntimeError, match=ERR_NO_PAIRS):
        _getPrimerPairs(cands, prm, dimer=lambda a, b: 99.0)
    # ... and one that rejects the pairs of some forward primers only
    some = _getPrimerPairs(cands, prm, dimer=lambda a, b: 99.0 if a[0] == "A" else -10.0)
    assert 0 < len(some) < len(out) and all(str(a)[0] != "A" for a, _ in some)
    assert all(per[f.names[0]].dimerTm == -10.0 for per in some.values())


def test_drop_in_on_the_lazy_candidates_of_this_package():
    """the stage's own lazy output goes to the pair search as columns (no Primer objects but the pairs' members) and gives
    what the generic path gives on the same candidates"""
    import tempfile
    from primerforge_b200 import synth
    from primerforge_b200.candidates import _getAllCandidateKmers
    from primerforge_b200.lazy import LazyPrimerList
    from primerforge_b200.pairs import _getPrimerPairs
    gens = synth.make_genomes(3, 60000, seed=77, snp_rate=0.003, n_inversions=2, n_contigs=2)
    log = SimpleNamespace(rename=lambda *a: None, info=lambda *a: None, debug=lambda *a: None, error=lambda *a: None)
    with tempfile.TemporaryDirectory() as d:
        paths = synth.write_fasta(gens, d)
        prm = SimpleNamespace(ingroupFns=paths, format="fasta", minLen=16, maxLen=20, minGc=40.0, maxGc=60.0, minTm=55.0, maxTm=68.0,
                              maxRepeatLen=4, tempTolerance=5.0, numThreads=1, p3_mvConc=50.0, p3_dvConc=1.5, p3_dntpConc=0.6,
                              p3_dnaConc=50.0, p3_tempC=37.0, p3_maxLoop=30, log=log, pickles={1: "s.p", 2: "c.p"},
                              dumpObj=lambda *a, **k: None, loadObj=lambda *a: None, debug=False, maxBinSize=64, minPcr=120,
                              maxPcr=2400, maxTmDiff=5.0)
        lazy = _getAllCandidateKmers(prm, False, thal="off")
        a = _getPrimerPairs(lazy, prm, dimer="off")
        untouched = [v for per in lazy.values() for v in dict.values(per) if isinstance(v, LazyPrimerList) and v._lazy is not None]
        assert untouched                                     # the other genomes' lists were never materialised
        plain = {n: {c: list(v) for c, v in per.items()} for n, per in lazy.items()}
        b = _getPrimerPairs(plain, prm, dimer="off")

    def norm(pairs):
        return [(str(f), str(r), [(n, p.contig, p.size, p.fbin, p.rbin, p.fStart, p.fEnd, p.fStrand, p.rStart, p.rEnd, p.rStrand)
                                  for n, p in per.items()]) for (f, r), per in pairs.items()]
    assert len(a) > 100 and norm(a) == norm(b)
