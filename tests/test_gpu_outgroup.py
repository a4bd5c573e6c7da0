"""CUDA outgroup screen (pfb_og_*, through the C ABI) against the reference's unmodified
bin/removeOutgroupPrimers.py (committed fixtures tests/golden/og_*.npz) and against the CPU restatement
oracle/pf_oracle_outgroup.c on seeded inputs."""
from types import SimpleNamespace

import numpy as np
import pytest

from tests.util import load_og, og_cases

pytestmark = pytest.mark.gpu

NOT_REMOVED = 0xFFFFFFFF


def _screen(keys_or_words, pair_fwd, pair_rev, lo, hi, genomes, k=None):
    """both scan kernels (filter in shared memory / in global memory) must agree"""
    a = _screen_mode(keys_or_words, pair_fwd, pair_rev, lo, hi, genomes, k, 0)
    b = _screen_mode(keys_or_words, pair_fwd, pair_rev, lo, hi, genomes, k, 1)
    assert a.timing["filter_in_smem"] == 1 and b.timing["filter_in_smem"] == 0
    assert np.array_equal(a.removed_at, b.removed_at) and np.array_equal(a.products, b.products) and np.array_equal(a.hits, b.hits)
    return a


def _screen_mode(keys_or_words, pair_fwd, pair_rev, lo, hi, genomes, k, mode):
    from primerforge_b200.outgroup import OutgroupScreen, strings_to_words
    with OutgroupScreen(device=0) as scr:
        scr.set_scan_mode(mode)
        if k is None:
            word, k = strings_to_words(keys_or_words)
        else:
            word = keys_or_words
        scr.set_keys_and_pairs(word, k, pair_fwd, pair_rev, lo, hi)
        for g in genomes:
            scr.add_genome(g.contigs)
        scr.run()
        res = scr.result()
        scr.compute()                       # the same again on the resident genomes
        res2 = scr.result()
    assert np.array_equal(res.removed_at, res2.removed_at) and np.array_equal(res.products, res2.products)
    assert np.array_equal(res.hits, res2.hits)
    return res


def _hits_array(res):
    h = res.hits
    a = np.stack([h["key"], h["genome"], h["contig"], h["start"], h["strand"]], axis=1).astype(np.int64)
    return np.array(sorted(map(tuple, a.tolist())), dtype=np.int64).reshape(-1, 5)


def _products_array(res):
    p = res.products
    return np.stack([p["pair"], p["genome"], p["contig"], p["size"]], axis=1).astype(np.int64).reshape(-1, 4)


@pytest.mark.parametrize("name", og_cases())
def test_equals_reference(name):
    g = load_og(name)
    res = _screen(g.keys, g.pair_fwd, g.pair_rev, g.bad_lo, g.bad_hi, g.genomes)
    assert np.array_equal(_hits_array(res), g.hits)                    # __extractKmerData (:34-62)
    if g.error:
        assert not res.kept.any()
        return
    assert np.array_equal(res.kept, g.kept)                            # the deletions (:281-287)
    assert np.array_equal(_products_array(res), g.products)            # the (contig, size) sets (:289-291)
    remaining = res.kept.size
    for gi, nm in enumerate(g.names):
        gone = int(np.count_nonzero(res.removed_at == gi))
        remaining -= gone
        assert g.debug[gi] == f"removed {gone} pairs after processing {nm} ({remaining} pairs remaining)"


def _random_case(seed, n_genomes, length, n_keys, n_pairs, kmin=16, kmax=20, rates=(0.0, 0.02, 0.05), n_contigs=4, n_frac=0.0005):
    from primerforge_b200 import synth
    rng = np.random.default_rng(seed)
    anc = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, length)]
    gens = []
    for g in range(n_genomes):
        seq = anc.copy()
        sites = np.flatnonzero(rng.random(length) < rates[g % len(rates)])
        seq[sites] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, sites.size)]
        seq[rng.random(length) < n_frac] = ord("N")
        cuts = np.sort(rng.choice(np.arange(1, length), size=n_contigs - 1, replace=False))
        pieces = [synth.revcomp_ascii(p) if i & 1 else np.ascontiguousarray(p) for i, p in enumerate(np.split(seq, cuts))]
        gens.append(synth.Genome(name=f"o{g}", contig_ids=[f"c{i}" for i in range(len(pieces))], contigs=pieces))
    keys = {}
    while len(keys) < n_keys:
        k = int(rng.integers(kmin, kmax + 1))
        a = int(rng.integers(0, length - k))
        s = anc[a:a + k]
        if rng.random() < 0.3:
            s = synth.revcomp_ascii(s)
        keys.setdefault(s.tobytes().decode(), None)
    keys = list(keys)
    pf = rng.integers(0, n_keys, n_pairs).astype(np.uint32)
    pr = rng.integers(0, n_keys, n_pairs).astype(np.uint32)
    return gens, keys, pf, pr


def _check_against_oracle(gens, keys, pf, pr, lo, hi):
    from oracle import pf_oracle as po
    o = po.run_outgroup(keys, pf, pr, lo, hi, gens)
    res = _screen(keys, pf, pr, lo, hi, gens)
    mine = np.array(sorted(map(tuple, o.hits.tolist())), dtype=np.int64).reshape(-1, 5)
    assert np.array_equal(_hits_array(res), mine)
    assert np.array_equal(res.removed_at, o.removed_at)
    assert np.array_equal(_products_array(res), o.products.astype(np.int64))
    return res, o


@pytest.mark.parametrize("seed,kmin,kmax", [(1, 16, 20), (2, 12, 30), (3, 20, 20), (4, 28, 32)])
def test_against_oracle(seed, kmin, kmax):
    gens, keys, pf, pr = _random_case(seed, 4, 150_000, 6000, 20000, kmin, kmax)
    res, o = _check_against_oracle(gens, keys, pf, pr, 200, 3000)
    assert res.kept.any() and not res.kept.all() and res.products.size > 0


def test_keys_that_are_each_others_reverse_complement_and_palindromes():
    from primerforge_b200 import synth
    rng = np.random.default_rng(9)
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 20000)].copy()
    pal = np.frombuffer(b"ACGTTGCATGCAACGT", dtype=np.uint8)            # its own reverse complement
    for a in (100, 5000, 12000):
        seq[a:a + 16] = pal
    gen = synth.Genome(name="o", contig_ids=["c0"], contigs=[seq])
    w1 = seq[300:318].tobytes().decode()
    w1rc = synth.revcomp_ascii(seq[300:318]).tobytes().decode()
    keys = [pal.tobytes().decode(), w1, w1rc, seq[700:716].tobytes().decode()]
    pf = np.array([0, 1, 2, 3, 0, 1], dtype=np.uint32)
    pr = np.array([0, 2, 1, 0, 3, 1], dtype=np.uint32)
    res, o = _check_against_oracle([gen], keys, pf, pr, 50, 60)
    # the palindrome binds both strands at each of its three places
    assert int(np.count_nonzero(res.hits["key"] == 0)) == 6


def test_many_hits_and_products_grow_the_buffers():
    """short keys: far more hits / products than the first capacities (the run is repeated with larger ones)"""
    gens, keys, pf, pr = _random_case(5, 2, 120_000, 300, 500, 6, 7, rates=(0.0,), n_contigs=2, n_frac=0.0)
    res, o = _check_against_oracle(gens, keys, pf, pr, 10, 20)
    assert res.timing["n_products"] > 65536 and res.timing["n_reruns"] >= 1
    gens, keys, pf, pr = _random_case(6, 2, 120_000, 3000, 4, 6, 7, rates=(0.0,), n_contigs=2, n_frac=0.0)
    res, o = _check_against_oracle(gens, keys, pf, pr, 10, 20)
    assert res.hits.size > 65536 and res.timing["n_reruns"] >= 1


def test_edge_inputs():
    from primerforge_b200 import synth
    from primerforge_b200.outgroup import OutgroupScreen
    from primerforge_b200.engine import PfbError
    e = np.zeros(0, dtype=np.uint8)
    key = ["ACGTACGTACGTACGTAC"]
    one = np.zeros(1, dtype=np.uint32)
    # no genome at all, a genome without contigs, empty and tiny contigs
    for gens in ([], [synth.Genome(name="o", contig_ids=[], contigs=[])],
                 [synth.Genome(name="o", contig_ids=["a", "b"], contigs=[e, np.frombuffer(b"ACG", dtype=np.uint8)])]):
        res = _screen(key, one, one, 1, 10, gens)
        assert res.hits.size == 0 and res.kept.all() and res.products.size == 0
    # no pairs, no keys
    g = synth.Genome(name="o", contig_ids=["a"], contigs=[np.frombuffer(b"ACGTACGTACGTACGTACGTACGT", dtype=np.uint8)])
    res = _screen([], np.zeros(0, np.uint32), np.zeros(0, np.uint32), 1, 10, [g])
    assert res.hits.size == 0 and res.removed_at.size == 0
    res = _screen(key, one, one, 1, 10, [g])
    assert sorted(res.hits["start"][res.hits["strand"] == 0].tolist()) == [0, 4]
    with OutgroupScreen(device=0) as scr:
        with pytest.raises(ValueError):
            scr.set_pairs(["ACGN"], ["ACGT"], 1, 2)
        with pytest.raises(PfbError):
            scr.set_keys_and_pairs(np.array([1, 1], np.uint64), np.array([4, 4], np.uint8), one, one, 1, 2)
            scr.run()
        with pytest.raises(PfbError):
            scr.set_keys_and_pairs(np.array([1], np.uint64), np.array([4], np.uint8), np.array([3], np.uint32), one, 1, 2)


class _Rec:
    def __init__(self, rid, seq):
        self.id, self.seq = rid, seq


class _P:
    """what the reference's pairs dict is keyed by: objects whose str() is the primer sequence"""
    def __init__(self, s):
        self.s = s

    def __str__(self):
        return self.s


@pytest.mark.parametrize("name", ["og_multi", "og_widek", "og_allgone"])
def test_drop_in_entry_point(name, capsys):
    """_removeOutgroupPrimers(outgroup, pairs, params): same in-place effect, Products, log lines and error as the reference"""
    from primerforge_b200.outgroup import ERR_ALL_IN_OUTGROUP, _removeOutgroupPrimers
    g = load_og(name)
    ingroup = SimpleNamespace(dimerTm=12.5)
    pair_keys = [(_P(f), _P(r)) for f, r in zip(g.fwd, g.rev)]
    pairs = {pk: {"ingroup_000.fasta": ingroup} for pk in pair_keys}
    outgroup = {gen.name: (_Rec(cid, c.tobytes().decode()) for cid, c in zip(gen.contig_ids, gen.contigs)) for gen in g.genomes}
    lines = []
    log = SimpleNamespace(rename=lambda *a: None, info=lambda *a: None, error=lambda *a: None, debug=lines.append)
    params = SimpleNamespace(disallowedLens=range(g.bad_lo, g.bad_hi + 1), log=log)
    if g.error:
        with pytest.raises(RuntimeError, match=ERR_ALL_IN_OUTGROUP):
            _removeOutgroupPrimers(outgroup, pairs, params)
        assert not pairs
        return
    _removeOutgroupPrimers(outgroup, pairs, params)
    out = capsys.readouterr().out
    assert "removing primer pairs present in the outgroup sequences" in out and "processing outgroup results" in out
    assert list(pairs.keys()) == [pk for pk, keep in zip(pair_keys, g.kept) if keep]          # order preserved
    assert lines == g.debug
    want = {}
    for pi, gi, c, s in g.products.tolist():
        want.setdefault((pi, gi), []).append((g.genomes[gi].contig_ids[c], s))
    for pi, pk in enumerate(pair_keys):
        if not g.kept[pi]:
            continue
        for gi, gen in enumerate(g.genomes):
            pr = pairs[pk][gen.name]
            assert pr.dimerTm == 12.5 and pr.fbin == -1 and pr.rbin == -1 and pr.fStrand == "" and pr.rEnd == -1
            exp = want.get((pi, gi))
            if exp is None:
                assert (pr.contig, pr.size) == ("NA", 0)
            elif len(exp) == 1:
                assert (pr.contig, pr.size) == exp[0]
            else:
                assert sorted(zip(pr.contig.split(","), map(int, pr.size.split(",")))) == sorted(exp)
