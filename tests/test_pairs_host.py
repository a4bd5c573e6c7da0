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
