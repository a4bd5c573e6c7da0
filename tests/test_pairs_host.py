"""host logic of the pair-search / outgroup drop-ins that needs no GPU: the candidate dict as columns (lazy containers vs
plain lists of Primers), primer strings as 2-bit words, the oracle's reading of Python's list-iterator quirk"""
import numpy as np
import pytest

from oracle import pf_oracle as po
from primerforge_b200 import synth
from primerforge_b200.lazy import Columns, build_lazy_output
from primerforge_b200.outgroup import strings_to_words
from primerforge_b200.pairs import _columns


@pytest.fixture(scope="module")
def lazy_case():
    gens = synth.make_genomes(3, 30000, seed=55, snp_rate=0.003, n_inversions=2, n_contigs=2)
    o = po.run(gens, po.default_params())
    sel = np.flatnonzero((o.flags & 8) != 0)
    pos = []
    for g, gen in enumerate(gens):
        # padded coordinates as the device reports them: contigs word-aligned in the order added
        ws = np.cumsum([0] + [(c.size + 31) // 32 for c in gen.contigs])[:-1]
        pos.append(((32 * ws[o.contig[g][sel]] + o.start[g][sel]).astype(np.uint32)) | (o.strand[g][sel].astype(np.uint32) << np.uint32(31)))
    ws_all = [np.cumsum([0] + [(c.size + 31) // 32 for c in gen.contigs])[:-1].astype(np.uint32) for gen in gens]
    cols = Columns([g.name for g in gens], [g.contig_ids for g in gens], ws_all, o.word[sel], o.k[sel], o.tm[sel], o.gc_count[sel], pos)
    return gens, cols


def test_lazy_containers_hand_over_their_columns(lazy_case):
    gens, cols = lazy_case
    lazy = build_lazy_output(cols)
    word, k, tm, names, per_genome, lists = _columns(lazy, gens[0].name)
    assert word is cols.word and names[0] == gens[0].name                  # no copy, no Primer object
    assert all(v._lazy is not None for per in lazy.values() for v in dict.values(per))
    plain = {n: {c: list(v) for c, v in per.items()} for n, per in build_lazy_output(cols).items()}
    word2, k2, tm2, names2, per2, _ = _columns(plain, gens[0].name)
    # the generic path numbers the candidates in the first genome's list order; map both to strings to compare
    from primerforge_b200.engine import words_to_strings
    s1, s2 = words_to_strings(word, k), words_to_strings(word2, k2)
    assert sorted(s1) == sorted(s2) and names == names2
    for (i1, c1, l1, t1), (i2, c2, l2, t2) in zip(per_genome, per2):
        assert [s1[i] for i in i1] == [s2[i] for i in i2]
        assert np.array_equal(c1, c2) and np.array_equal(l1, l2) and np.array_equal(t1, t2)
    assert np.array_equal(tm[per_genome[0][0]], tm2[per2[0][0]])


def test_an_edited_lazy_list_is_read_object_by_object(lazy_case):
    gens, cols = lazy_case
    lazy = build_lazy_output(cols)
    first = next(iter(lazy[gens[1].name].values()))
    first.append(first.pop())                                              # same content, but no longer described by the columns
    assert first._view is None
    word, k, tm, names, per_genome, _ = _columns(lazy, gens[0].name)
    assert word is not cols.word and per_genome[1][0].size == cols.n


def test_strings_to_words():
    w, k = strings_to_words(["A", "ACGT", "T" * 32, "GATTACA"])
    assert k.tolist() == [1, 4, 32, 7] and w[0] == 0 and w[1] == 0b00101101 and w[2] == 0x5555555555555555
    for bad in ("", "ACGN", "acgt", "A" * 33):
        with pytest.raises(ValueError):
            strings_to_words([bad])


def test_reduce_bin_size_skips_the_element_after_a_removed_one():
    """getPrimerPairs.py:28-36 removes from the list it iterates: with three 16-mers in one bin that are all substrings of a
    17-mer, Python removes the first and the third and never looks at the second; the oracle does the same"""
    seq = "GCTAGCTAGGATCCGATCGGA"
    big = seq[0:18]
    smalls = [seq[0:16], seq[1:17], seq[2:18]]              # all inside `big`
    cand = smalls + [big]
    word, k = strings_to_words(cand)
    ids = np.arange(4, dtype=np.uint32)
    left = np.array([0, 1, 2, 0], dtype=np.uint32)
    order = np.lexsort((k, left))                          # list order: by start, then length
    g0 = (ids[order], np.zeros(4, np.uint32), left[order], np.zeros(4, np.uint8))
    o = po.run_pairs(word, k, np.full(4, 60.0), [g0], max_bin=64)
    assert o.bin_of[0].tolist() == [0, 0, 0, 0]
    # list order of the 16-mers within the bin is [start 0, start 1, start 2] -> the one at start 1 survives
    assert o.reduced.tolist() == [1, 0, 1, 0]


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


@pytest.mark.parametrize("name", ["pp_contigs", "pp_widek", "pp_tightbins"])
def test_host_half_of_the_drop_in_builds_the_reference_products(name):
    """build_pairs_output on the ORACLE's pair-search result (no GPU): the reference's pairs, in order, with every Product
    field -- the same check tests/test_gpu_pairs.py makes with the CUDA result"""
    from types import SimpleNamespace
    from primerforge_b200 import _lib
    from primerforge_b200.pairs import ERR_NO_PAIRS, PairResult, build_pairs_output
    from tests.util import load_pp
    f = load_pp(name)
    cands = {}
    for g, nm in enumerate(f.names):
        gid, gct, glf, gst = f.genomes[g]
        per = {}
        for p in range(gid.size):
            s = int(gid[p])
            per.setdefault(f.contig_names[g][int(gct[p])], []).append(_P(f.cand[s], f.contig_names[g][int(gct[p])], int(glf[p]), int(f.k[s]), bool(gst[p]), float(f.tm[s])))
        cands[nm] = per
    word, k, tm, names, per_genome, first_lists = _columns(cands, f.names[0])
    o = po.run_pairs(word, k, tm, per_genome, max_bin=f.max_bin, min_primer_len=f.min_len, min_pcr=f.min_pcr, max_pcr=f.max_pcr,
                     max_tm_diff=f.max_tm_diff)
    pairs = np.zeros(o.pairs.shape[0], dtype=_lib.PP_PAIR_DTYPE)
    for j, nm in enumerate(("fwd", "rev", "contig", "fbin", "rbin", "pcr_len")):
        pairs[nm] = o.pairs[:, j]
    res = PairResult(pairs=pairs, valid=o.shared[:, :, 0] != 0, length=o.shared[:, :, 1], kept=o.kept, bin_of=list(o.bin_of), timing={})
    assert np.array_equal(res.shared4(), o.shared)
    log = SimpleNamespace(rename=lambda *a: None, info=lambda *a: None, debug=lambda *a: None, error=lambda *a: None)
    prm = SimpleNamespace(tempTolerance=5.0, log=log)
    contig_names = [list(cands[n].keys()) for n in names]
    out = build_pairs_output(res, k, names, per_genome, first_lists, contig_names, prm, None)
    rc = str.maketrans("ACGT", "TGCA")
    assert [(str(a), str(b)) for a, b in out] == [(f.cand[int(x)], f.cand[int(y)][::-1].translate(rc)) for x, y in zip(f.pair_fwd, f.pair_rev)]
    for row, per in enumerate(out.values()):
        for g, nm in enumerate(f.names):
            pr = per[nm]
            got = (f.contig_names[g].index(pr.contig), pr.size, pr.fbin, pr.rbin, pr.fStart, pr.fEnd, int(pr.fStrand == "-"), pr.rStart,
                   pr.rEnd, int(pr.rStrand == "-"))
            assert got == tuple(int(x) for x in f.products[row, g]) and pr.dimerTm == 0.0
    with pytest.raises(RuntimeError, match=ERR_NO_PAIRS):
        build_pairs_output(res, k, names, per_genome, first_lists, contig_names, prm, lambda a, b: 99.0)
    some = build_pairs_output(res, k, names, per_genome, first_lists, contig_names, prm, lambda a, b: 99.0 if a[0] == "A" else -10.0)
    assert 0 < len(some) < len(out) and all(str(a)[0] != "A" for a, _ in some)


@pytest.mark.parametrize("name", ["og_multi", "og_widek", "og_allgone"])
def test_host_half_of_the_outgroup_drop_in(name, capsys):
    """apply_outgroup_result on the ORACLE's result (no GPU): the reference's deletions, debug lines, error and Products"""
    from types import SimpleNamespace
    from primerforge_b200 import _lib
    from primerforge_b200.outgroup import ERR_ALL_IN_OUTGROUP, OutgroupResult, apply_outgroup_result
    from tests.util import load_og
    g = load_og(name)
    o = po.run_outgroup(g.keys, g.pair_fwd, g.pair_rev, g.bad_lo, g.bad_hi, g.genomes)
    prod = np.zeros(o.products.shape[0], dtype=_lib.OG_PRODUCT_DTYPE)
    for j, nm in enumerate(("pair", "genome", "contig", "size")):
        prod[nm] = o.products[:, j]
    res = OutgroupResult(removed_at=o.removed_at, products=prod, hits=np.zeros(0, dtype=_lib.OG_HIT_DTYPE), timing={})

    class P:
        def __init__(self, s):
            self.s = s

        def __str__(self):
            return self.s
    pair_keys = [(P(f), P(r)) for f, r in zip(g.fwd, g.rev)]
    pairs = {pk: {"ingroup_000.fasta": SimpleNamespace(dimerTm=7.25)} for pk in pair_keys}
    lines = []
    log = SimpleNamespace(rename=lambda *a: None, info=lambda *a: None, error=lambda *a: None, debug=lines.append)
    params = SimpleNamespace(log=log)
    contig_ids = [gen.contig_ids for gen in g.genomes]
    if g.error:
        with pytest.raises(RuntimeError, match=ERR_ALL_IN_OUTGROUP):
            apply_outgroup_result(res, pairs, pair_keys, g.names, contig_ids, params)
        assert not pairs
        return
    apply_outgroup_result(res, pairs, pair_keys, g.names, contig_ids, params)
    assert "filtering primer pairs" in capsys.readouterr().out
    assert list(pairs.keys()) == [pk for pk, keep in zip(pair_keys, g.kept) if keep] and lines == g.debug
    want = {}
    for pi, gi, c, size in g.products.tolist():
        want.setdefault((pi, gi), []).append((g.genomes[gi].contig_ids[c], size))
    for pi, pk in enumerate(pair_keys):
        if not g.kept[pi]:
            continue
        for gi, gen in enumerate(g.genomes):
            pr = pairs[pk][gen.name]
            assert pr.dimerTm == 7.25 and pr.fbin == -1 and pr.rStrand == ""
            exp = want.get((pi, gi))
            if exp is None:
                assert (pr.contig, pr.size) == ("NA", 0)
            elif len(exp) == 1:
                assert (pr.contig, pr.size) == exp[0]
            else:
                assert sorted(zip(pr.contig.split(","), map(int, pr.size.split(",")))) == sorted(exp)
