"""The columnar, lazy return value of the drop-in (primerforge_b200/lazy.py): same container types and contents as
the reference's dict[genome][contig] -> list[Primer], objects created on first use, pickles in the reference's
layout -- and the reference's own next stage (bin/getPrimerPairs.py, run here when /root/reference is present)
consumes it unchanged."""
import os
import pickle
import sys

import numpy as np
import pytest

from oracle import pf_oracle as po
from primerforge_b200 import synth
from primerforge_b200.lazy import Columns, LazyGenome, LazyPrimerList, build_lazy_output
from primerforge_b200.primer import primer_class

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _columns_from_oracle(gens, o):
    """what candidates._Stage.columns produces, made from an oracle result: the picked k-mers and their places"""
    sel = np.flatnonzero(o.flags & 8)
    pos, wstarts = [], []
    for g, gen in enumerate(gens):
        lens = [c.size for c in gen.contigs]
        ws = None
        if len(lens) > 1:
            ws = np.concatenate([[0], np.cumsum((np.array(lens) + 31) // 32)[:-1]]).astype(np.int64)
        lp = o.start[g][sel] + (32 * ws[o.contig[g][sel]] if ws is not None else 0)
        pos.append((lp.astype(np.uint32) | (o.strand[g][sel].astype(np.uint32) << np.uint32(31))))
        wstarts.append(ws)
    return Columns([g.name for g in gens], [g.contig_ids for g in gens], wstarts, o.word[sel], o.k[sel], o.tm[sel],
                   o.gc_count[sel], pos), sel


@pytest.fixture(scope="module")
def case():
    gens = synth.make_genomes(3, 40_000, seed=71, snp_rate=0.003, n_inversions=2, n_contigs=3)
    o = po.run(gens, po.default_params())
    cols, sel = _columns_from_oracle(gens, o)
    return gens, o, cols, sel


def test_containers_are_the_reference_types_and_fill_on_first_use(case):
    gens, o, cols, sel = case
    out = build_lazy_output(cols)
    assert list(out.keys()) == [g.name for g in gens]
    first = out[gens[0].name]
    assert isinstance(first, dict) and isinstance(first, LazyGenome)
    assert first.n_candidates() == sel.size and first._lazy is not None          # counted from the columns
    contigs = list(first.keys())                                                  # first real use: sorted, split per contig
    assert first._lazy is None and set(contigs) <= set(gens[0].contig_ids)
    lst = first[contigs[0]]
    assert isinstance(lst, list) and isinstance(lst, LazyPrimerList)
    n = len(lst)
    assert lst._lazy is not None and n > 0                                        # len() needs no objects
    p0 = lst[0]
    assert lst._lazy is None and list.__len__(lst) == n                           # now an ordinary populated list
    assert isinstance(p0, primer_class())
    assert sum(len(v) for v in first.values()) == sel.size


def test_contents_equal_the_oracle_candidates(case):
    gens, o, cols, sel = case
    out = build_lazy_output(cols)
    kmers = o.kmer_strings()
    for g, gen in enumerate(gens):
        want = set()
        for r in sel:
            k = int(o.k[r]); s = int(o.start[g][r])
            a, b = (s, s + k - 1) if o.strand[g][r] == 0 else (s + k - 1, s)
            want.add((kmers[r], gen.contig_ids[int(o.contig[g][r])], a, b, "+" if o.strand[g][r] == 0 else "-", float(o.tm[r])))
        got = set()
        for contig, plist in out[gen.name].items():
            keys = [min(p.start, p.end) for p in plist]
            assert keys == sorted(keys)                                            # getCandidateKmers.py:602
            for p in plist:
                assert p.contig == contig and len(p) == len(str(p.seq)) and p.hairpinTm is None
                assert abs(p.gcPer - sum(c in "GC" for c in str(p.seq)) / len(p) * 100) < 1e-12
                got.add((str(p.seq), p.contig, p.start, p.end, p.strand, p.Tm))
        assert got == want


def test_pickles_as_plain_dicts_and_lists(case):
    gens, o, cols, sel = case
    out = build_lazy_output(cols)
    blob = pickle.dumps(out)
    assert b"primerforge_b200.lazy" not in blob                                   # candidates.p loads without this package's containers
    back = pickle.loads(blob)
    for name in out:
        assert type(back[name]) is dict
        for contig, plist in back[name].items():
            assert type(plist) is list
            assert [(str(p.seq), p.start, p.end, p.strand, p.Tm, p.gcPer) for p in plist] == \
                   [(str(p.seq), p.start, p.end, p.strand, p.Tm, p.gcPer) for p in out[name][contig]]


def test_list_and_dict_behaviour(case):
    gens, o, cols, sel = case
    a = build_lazy_output(cols)[gens[1].name]
    b = build_lazy_output(cols)[gens[1].name]
    assert a == b and len(a) == len(b) and (gens[1].contig_ids[0] in a) == (gens[1].contig_ids[0] in b)
    contig = next(iter(a))
    la, lb = a[contig], b[contig]
    assert la == lb and la[-1] == lb[-1] and la[1:3] == lb[1:3] and (la[0] in la)
    assert sorted(la, key=lambda p: -min(p.start, p.end))[0] == la[-1] or len(la) == 1
    assert list(reversed(la))[0] == la[-1] and la + [] == list(lb) and la.index(la[0]) == 0


def test_the_reference_next_stage_consumes_it_unchanged(case, capsys):
    """bin/getPrimerPairs.py::_getPrimerPairs (unmodified, from /root/reference) on the lazy result gives the same
    pairs as on eagerly built plain containers of reference Primers"""
    if not os.path.isfile("/root/reference/bin/getPrimerPairs.py"):
        pytest.skip("the reference is only present in the build container")
    for pth in ("/root/reference", os.path.join(ROOT, "oracle", "standins")):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    os.environ.setdefault("CI", "1")
    from types import SimpleNamespace
    import bin.getPrimerPairs as gpp
    from bin.Primer import Primer as RefPrimer
    gens, o, cols, sel = case
    cols2, _ = _columns_from_oracle(gens, o)
    assert cols2.cls is RefPrimer                                                  # real bin.Primer.Primer objects come out
    lazy = build_lazy_output(cols2)
    eager = pickle.loads(pickle.dumps(build_lazy_output(_columns_from_oracle(gens, o)[0])))     # plain dicts / lists
    log = SimpleNamespace(rename=lambda *a: None, info=lambda *a: None, debug=lambda *a: None, error=lambda *a: None)
    prm = SimpleNamespace(ingroupFns=[g.name for g in gens], maxBinSize=64, minLen=16, maxLen=20, minPcr=120, maxPcr=2400,
                          maxTmDiff=5.0, numThreads=1, tempTolerance=5.0, p3_mvConc=50.0, p3_dvConc=1.5, p3_dntpConc=0.6,
                          p3_dnaConc=50.0, p3_tempC=37.0, p3_maxLoop=30, log=log)
    pairs_lazy = gpp._getPrimerPairs(lazy, prm)
    pairs_eager = gpp._getPrimerPairs(eager, prm)
    capsys.readouterr()

    def norm(pairs):
        return sorted((str(f), str(r), sorted((n, p.contig, p.size, p.fbin, p.rbin) for n, p in prods.items()))
                      for (f, r), prods in pairs.items())
    assert len(pairs_lazy) > 0 and norm(pairs_lazy) == norm(pairs_eager)
