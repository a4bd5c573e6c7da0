"""GPU parity tests: libpfb200.so (through the C ABI) against the committed reference outputs and
against the CPU oracle on seeded inputs.  Integer / index / strand results are compared bit for
bit; Tm and GC % must agree within 0.01 (BASELINE.json north_star) -- they are in fact identical."""
import zlib

import numpy as np
import pytest

from tests.util import candidate_set_from_arrays, golden_candidate_set, golden_cases, load_golden

pytestmark = pytest.mark.gpu

TOL = 0.01


def _engine(gold_or_kw, prefilter):
    from primerforge_b200.engine import Engine
    kw = gold_or_kw if isinstance(gold_or_kw, dict) else dict(
        k_min=gold_or_kw.k_min, k_max=gold_or_kw.k_max, min_gc=gold_or_kw.min_gc, max_gc=gold_or_kw.max_gc,
        min_tm=gold_or_kw.min_tm, max_tm=gold_or_kw.max_tm, max_repeat=gold_or_kw.max_repeat,
        table_size=gold_or_kw.table_size)
    return Engine(device=0, prefilter=int(prefilter), **kw)


def _pseudo_thal(seq, rc):
    def one(s, salt):
        return ((zlib.crc32(s.encode("ascii")) ^ ((salt * 0x9E3779B1) & 0xFFFFFFFF)) % 6400) / 100.0
    return one(seq, 1), one(rc, 1), one(seq, 2), one(rc, 2)


def _run(genomes, kw, prefilter, scan_mode=0):
    eng = _engine(kw, prefilter)
    eng.set_scan_mode(scan_mode)
    for g in genomes:
        eng.add_genome(g.contigs)
    eng.run()
    res = eng.result()
    eng.close()
    if scan_mode:
        assert res.timing["used_filter"] == (1 if scan_mode == 2 else 0)
    return res, None


def _picked(res, thal, limit):
    from primerforge_b200.candidates import screen_groups
    if thal == "pseudo":
        picked, _ = screen_groups(res.records, res.positions, res.kmer_strings(), _pseudo_thal, limit)
        return picked
    return res.picked


@pytest.mark.parametrize("scan_mode", [1, 2], ids=["l2probe", "smemfilter"])
@pytest.mark.parametrize("name", golden_cases())
def test_full_mode_equals_reference_shared_dict(name, scan_mode):
    gold = load_golden(name)
    res, _ = _run(gold.genomes, gold, prefilter=False, scan_mode=scan_mode)
    kmers = res.kmer_strings()
    assert kmers == gold.shared                      # set AND dict insertion order
    for gi in range(len(gold.genomes)):
        pos = res.positions[gi]
        assert np.array_equal(pos["contig"], gold.z[f"g{gi}_sh_contig"])
        assert np.array_equal(pos["start"].astype(np.int64), gold.z[f"g{gi}_sh_start"])
        assert np.array_equal(pos["strand"].astype(np.uint8), gold.z[f"g{gi}_sh_strand"])
    picked = _picked(res, gold.thal, gold.min_tm - gold.temp_tolerance)
    for gi in range(len(gold.genomes)):
        pos = res.positions[gi]
        mine = candidate_set_from_arrays(kmers, res.records["k"], picked, pos["contig"], pos["start"], pos["strand"])
        assert mine == golden_candidate_set(gold, gi)
    # Tm / GC of the reference's candidates
    z = gold.z
    idx = z["g0_cand_kmer"]
    assert np.max(np.abs(res.records["tm"][idx] - z["g0_cand_tm"]), initial=0.0) <= TOL
    gc = res.records["gc_count"][idx] / res.records["k"][idx] * 100
    assert np.max(np.abs(gc - z["g0_cand_gc"]), initial=0.0) <= TOL


@pytest.mark.parametrize("scan_mode", [1, 2], ids=["l2probe", "smemfilter"])
@pytest.mark.parametrize("name", golden_cases())
def test_prefilter_mode_equals_reference_candidates(name, scan_mode):
    gold = load_golden(name)
    res, _ = _run(gold.genomes, gold, prefilter=True, scan_mode=scan_mode)
    assert np.all((res.records["flags"] & 7) == 7)
    kmers = res.kmer_strings()
    picked = _picked(res, gold.thal, gold.min_tm - gold.temp_tolerance)
    for gi in range(len(gold.genomes)):
        pos = res.positions[gi]
        mine = candidate_set_from_arrays(kmers, res.records["k"], picked, pos["contig"], pos["start"], pos["strand"])
        assert mine == golden_candidate_set(gold, gi)


CASES = [
    dict(n=3, length=200_000, seed=201, n_contigs=1, kw=dict()),
    dict(n=5, length=60_000, seed=202, n_contigs=7, kw=dict(table_size=1000003)),
    dict(n=4, length=50_000, seed=203, n_contigs=3, n_frac=0.0005, kw=dict(table_size=65521, k_min=12, k_max=30, max_repeat=3)),
    dict(n=2, length=30_000, seed=204, n_contigs=40, kw=dict(table_size=20011, k_min=10, k_max=32)),
    dict(n=1, length=50_000, seed=205, n_contigs=2, kw=dict()),
    dict(n=6, length=40_000, seed=206, n_contigs=1, kw=dict(k_min=20, k_max=20, min_gc=30.0, max_gc=70.0, min_tm=50.0, max_tm=75.0)),
]


@pytest.mark.parametrize("scan_mode", [1, 2], ids=["l2probe", "smemfilter"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"seed{c['seed']}")
@pytest.mark.parametrize("prefilter", [False, True])
def test_against_oracle(case, prefilter, scan_mode):
    from oracle import pf_oracle as po
    from primerforge_b200 import synth
    gens = synth.make_genomes(case["n"], case["length"], seed=case["seed"], snp_rate=0.003, n_inversions=3,
                              n_contigs=case["n_contigs"], n_frac=case.get("n_frac", 0.0))
    kw = dict(k_min=16, k_max=20, min_gc=40.0, max_gc=60.0, min_tm=55.0, max_tm=68.0, max_repeat=4, table_size=9999991)
    kw.update(case["kw"])
    o = po.run(gens, po.default_params(**kw))
    res, _ = _run(gens, kw, prefilter, scan_mode)
    rec = res.records
    if not prefilter:
        assert rec.size == o.n_shared
        assert np.array_equal(rec["word"], o.word) and np.array_equal(rec["k"], o.k)
        assert np.array_equal(rec["flags"] & 15, o.flags)
        assert np.array_equal(rec["gc_count"].astype(np.int32), o.gc_count)
        assert np.max(np.abs(rec["tm"] - o.tm), initial=0.0) <= TOL
        assert np.array_equal(rec["tm"], o.tm)                   # same IEEE operations -> identical doubles
        for gi in range(len(gens)):
            pos = res.positions[gi]
            assert np.array_equal(pos["contig"].astype(np.int32), o.contig[gi])
            assert np.array_equal(pos["start"].astype(np.int64), o.start[gi])
            assert np.array_equal(pos["strand"].astype(np.uint8), o.strand[gi])
    else:
        keep = (o.flags & 7) == 7
        assert rec.size == int(keep.sum())
        assert np.array_equal(rec["word"], o.word[keep]) and np.array_equal(rec["k"], o.k[keep])
        assert np.array_equal((rec["flags"] & 8) != 0, (o.flags[keep] & 8) != 0)
        for gi in range(len(gens)):
            pos = res.positions[gi]
            assert np.array_equal(pos["contig"].astype(np.int32), o.contig[gi][keep])
            assert np.array_equal(pos["start"].astype(np.int64), o.start[gi][keep])
            assert np.array_equal(pos["strand"].astype(np.uint8), o.strand[gi][keep])


def test_readme_tm_rows_on_device():
    from primerforge_b200.engine import Engine
    from tests.test_oracle_tm import README_ROWS
    from oracle import pf_oracle as po
    with Engine(device=0) as eng:
        for seq, tm, gc in README_ROWS + [("GAATTCGAATTC", None, None), ("AGAGTTTGATCCTGGCTCAG", None, None)]:
            dtm, dgc = eng.oligo_tm(seq)
            assert dtm == po.calc_tm(seq)
            if tm is not None:
                assert round(dtm, 1) == tm and round(dgc, 1) == gc


def test_no_shared_kmers_raises():
    from primerforge_b200 import synth
    from primerforge_b200.candidates import run_stage
    from primerforge_b200.engine import ERR_NO_SHARED
    a = synth.make_genomes(1, 5000, seed=301)[0]
    b = synth.make_genomes(1, 5000, seed=302)[0]
    with pytest.raises(RuntimeError, match=ERR_NO_SHARED):
        run_stage([(a.contig_ids, a.contigs), (b.contig_ids, b.contigs)])


def test_empty_and_ragged_inputs():
    from primerforge_b200 import synth
    from primerforge_b200.engine import Engine, PfbError
    from oracle import pf_oracle as po
    gens = synth.make_genomes(3, 20_000, seed=303, snp_rate=0.002, n_inversions=1)
    # ragged: an empty contig, a 1-base contig and a contig shorter than k in the middle of genome 2
    a = gens[1].contigs[0]
    gens[1].contigs = [a[:5000], a[5000:5000], a[5000:5001], a[5001:5012], a[5012:]]
    gens[1].contig_ids = [f"c{i}" for i in range(5)]
    kw = dict(k_min=16, k_max=20, min_gc=40.0, max_gc=60.0, min_tm=55.0, max_tm=68.0, max_repeat=4, table_size=9999991)
    o = po.run(gens, po.default_params(**kw))
    res, _ = _run(gens, kw, prefilter=False)
    assert np.array_equal(res.records["word"], o.word)
    for gi in range(3):
        assert np.array_equal(res.positions[gi]["contig"].astype(np.int32), o.contig[gi])
        assert np.array_equal(res.positions[gi]["start"].astype(np.int64), o.start[gi])
    with Engine(device=0) as eng:
        with pytest.raises(PfbError):
            eng.run()                                   # no genome added


def test_picked_output_path_matches_full_output():
    from primerforge_b200 import synth
    from primerforge_b200.engine import Engine
    gens = synth.make_genomes(4, 50_000, seed=401, snp_rate=0.003, n_inversions=2, n_contigs=3)
    for prefilter in (0, 1):
        with Engine(device=0, prefilter=prefilter) as eng:
            for g in gens:
                eng.add_genome(g.contigs)
            eng.run()
            res = eng.result()
            rec, ss, ct = eng.result_picked()
            sel = res.picked
            assert rec.size == int(sel.sum()) and ct is not None
            assert np.array_equal(rec["word"], res.records["word"][sel])
            assert np.all(rec["flags"] & 8)
            for gi in range(len(gens)):
                pos = res.positions[gi][sel]
                assert np.array_equal(ss[gi] & 0x7FFFFFFF, pos["start"])
                assert np.array_equal(ss[gi] >> 31, pos["strand"])
                assert np.array_equal(ct[gi], pos["contig"])
                pk = eng.positions(gi, picked_only=True)
                assert np.array_equal(pk["start"], pos["start"]) and np.array_equal(pk["contig"], pos["contig"])


@pytest.mark.parametrize("name,thal", [("basic", "off"), ("collide_multi", "pseudo")])
def test_drop_in_entry_point(name, thal, tmp_path, capsys, monkeypatch):
    """_getAllCandidateKmers(params, sharedExists) from FASTA files, as bin/main.py:86 calls it"""
    from types import SimpleNamespace
    from primerforge_b200 import synth
    from primerforge_b200.candidates import _getAllCandidateKmers
    gold = load_golden(name)
    if gold.table_size != 9999991:
        pytest.skip("the entry point always uses khmer's table size") if thal == "off" else None
    paths = synth.write_fasta(gold.genomes, str(tmp_path))
    dumped = {}
    log = SimpleNamespace(rename=lambda *a: None, info=lambda *a: None, debug=lambda *a: None, error=lambda *a: None)
    params = SimpleNamespace(ingroupFns=paths, format="fasta", minLen=gold.k_min, maxLen=gold.k_max, minGc=gold.min_gc,
                             maxGc=gold.max_gc, minTm=gold.min_tm, maxTm=gold.max_tm, maxRepeatLen=gold.max_repeat,
                             tempTolerance=gold.temp_tolerance, numThreads=1, p3_mvConc=50.0, p3_dvConc=1.5,
                             p3_dntpConc=0.6, p3_dnaConc=50.0, p3_tempC=37.0, p3_maxLoop=30, log=log,
                             pickles={1: "sharedKmers.p", 2: "candidates.p"},
                             dumpObj=lambda obj, fn, what, prefix="": dumped.__setitem__(fn, obj), loadObj=dumped.get)
    if gold.table_size != 9999991:
        monkeypatch.setenv("PFB200_TABLE_SIZE", str(gold.table_size))      # the golden case forces a small table
    out = _getAllCandidateKmers(params, False, thal=("off" if thal == "off" else _pseudo_thal))
    assert "candidate kmers" in capsys.readouterr().out
    assert dumped["candidates.p"] is out and list(out.keys()) == gold.names
    for gi, (nm, gen) in enumerate(zip(gold.names, gold.genomes)):
        cidx = {c: i for i, c in enumerate(gen.contig_ids)}
        mine = set()
        for contig, plist in out[nm].items():
            keys = [min(p.start, p.end) for p in plist]
            assert keys == sorted(keys)
            mine |= {(str(p), cidx[contig], p.start, p.end, int(p.strand == "-")) for p in plist}
        assert mine == golden_candidate_set(gold, gi)


@pytest.mark.parametrize("config", [2, 5])
def test_full_size_properties(config):
    """BASELINE configs 2 and 5 at full size: size-independent properties of the result (the oracle would
    take minutes here).  Every reported k-mer must read back from every genome at the reported
    place and strand, be unique there among the reported set, and the order must be the dict order.
    Config 5 adds 50 contigs per genome, 19 lengths (k 12-30, the 64-bit modular reduction) and max_repeats 3."""
    from primerforge_b200 import synth
    from primerforge_b200.engine import Engine
    gens, cfg = synth.make_config(config)
    with Engine(device=0, k_min=cfg["k_min"], k_max=cfg["k_max"], max_repeat=cfg["max_repeats"]) as eng:
        for g in gens:
            eng.add_genome(g.contigs)
        eng.run()
        rec, ss, ct = eng.result_picked()
        rec, ss = rec.copy(), ss.copy()
        ct = None if ct is None else ct.copy()
        tm = eng.timing()
    multi = cfg["n_contigs"] > 1
    assert (ct is not None) == multi and rec.size > 100_000 and tm["used_filter"] == 1
    k = rec["k"].astype(np.int64)
    assert np.all((rec["flags"] & 15) == 15)
    assert k.min() >= cfg["k_min"] and k.max() <= cfg["k_max"]
    # dict order: genome-1 contig, end position ascending, longest first at one end
    c0 = ct[0].astype(np.int64) if multi else np.zeros(rec.size, dtype=np.int64)
    end0 = (ss[0] & 0x7FFFFFFF).astype(np.int64) + k - 1
    order_key = (c0 << 40) + end0 * 64 + (63 - k)
    assert np.all(np.diff(order_key) > 0)
    assert np.all((ss[0] >> 31) == 0)                       # genome 1 holds every key on its (+) strand
    code = np.full(256, 4, dtype=np.uint8)
    for ch, v in zip(b"ATCG", range(4)):
        code[ch] = v
    rng = np.random.default_rng(5)
    sample = rng.choice(rec.size, size=3000, replace=False)
    # no homopolymer of max_repeats or more in any candidate (checked on the 2-bit words of the sample)
    for w, kk in zip(rec["word"][sample][:500], k[sample][:500]):
        bases = [(int(w) >> (2 * (int(kk) - 1 - i))) & 3 for i in range(int(kk))]
        run = best = 1
        for i in range(1, len(bases)):
            run = run + 1 if bases[i] == bases[i - 1] else 1
            best = max(best, run)
        assert best < cfg["max_repeats"]
    for gi, gen in enumerate(gens):
        seqs = [code[c] for c in gen.contigs]
        start = (ss[gi] & 0x7FFFFFFF).astype(np.int64)[sample]
        minus = (ss[gi] >> 31).astype(bool)[sample]
        contig = ct[gi][sample] if multi else np.zeros(sample.size, dtype=np.int64)
        for s0, mi, ci, w, kk in zip(start, minus, contig, rec["word"][sample], k[sample]):
            window = seqs[int(ci)][s0:s0 + kk]
            assert window.size == kk
            if mi:
                window = (window[::-1] ^ 1)
            val = 0
            for c in window:
                val = (val << 2) | int(c)
            assert val == int(w)
        # a position holds at most one picked k-mer per k
        cg = ct[gi].astype(np.int64) if multi else np.zeros(rec.size, dtype=np.int64)
        assert np.unique((cg << 40) | (ss[gi].astype(np.int64) << 6) | k).size == rec.size


def test_pipelined_run_equals_upload_then_compute():
    from primerforge_b200 import synth
    from primerforge_b200.engine import Engine
    gens = synth.make_genomes(9, 700_000, seed=402, snp_rate=0.003, n_inversions=2, n_contigs=2)   # several 16 MB batches
    with Engine(device=0) as eng:
        for g in gens:
            eng.add_genome(g.contigs)
        eng.run()
        a = eng.result()
        eng.upload()
        eng.compute()
        b = eng.result()
    assert a.records.size > 0 and np.array_equal(a.records, b.records)
    for pa, pb in zip(a.positions, b.positions):
        assert np.array_equal(pa, pb)


def _random_case(seed):
    from primerforge_b200 import synth
    rng = np.random.default_rng(seed)
    n_g = int(rng.integers(1, 6))
    length = int(rng.integers(2_000, 30_000))
    n_contigs = int(rng.choice([1, 1, 2, 5, 12]))
    k_min = int(rng.integers(8, 30))
    k_max = int(min(32, k_min + rng.integers(0, 8)))
    table = int(rng.choice([4093, 20011, 65521, 1000003, 9999991, 16777259]))     # the last one is > 2^24
    kw = dict(k_min=k_min, k_max=k_max, max_repeat=int(rng.integers(2, 6)), table_size=table,
              min_gc=float(rng.choice([30.0, 40.0])), max_gc=float(rng.choice([60.0, 70.0])),
              min_tm=float(rng.choice([45.0, 55.0])), max_tm=float(rng.choice([68.0, 80.0])))
    gens = synth.make_genomes(n_g, length, seed=1000 + seed, snp_rate=float(rng.choice([0.0, 0.002, 0.01])),
                              n_inversions=int(rng.integers(0, 4)), n_contigs=n_contigs,
                              n_frac=float(rng.choice([0.0, 0.0, 0.001])))
    if rng.random() < 0.3 and n_g > 1:            # a genome with ragged tiny contigs
        a = gens[-1].contigs[0]
        cut = [0, 3, 3, 10, int(a.size // 2), int(a.size // 2) + k_min - 1, a.size]
        pieces = [a[cut[i]:cut[i + 1]] for i in range(len(cut) - 1)] + gens[-1].contigs[1:]
        gens[-1].contigs = [np.ascontiguousarray(x) for x in pieces]
        gens[-1].contig_ids = [f"r{i}" for i in range(len(pieces))]
    return gens, kw


@pytest.mark.parametrize("seed", range(24))
def test_randomised_differential(seed):
    """random small configurations (lengths, contig counts, k ranges up to 32, table sizes on both sides
    of 2^24, N bases, ragged contigs, duplicated genomes) against the oracle, both scan kernels"""
    from oracle import pf_oracle as po
    gens, kw = _random_case(seed)
    o = po.run(gens, po.default_params(**kw))
    for scan_mode in (1, 2):
        for prefilter in (False, True):
            try:
                res, _ = _run(gens, kw, prefilter, scan_mode)
            except Exception as exc:      # no shared k-mer at all: both sides must agree
                assert o.n_shared == 0 or (prefilter and not ((o.flags & 7) == 7).any()), exc
                continue
            keep = np.ones(o.n_shared, bool) if not prefilter else (o.flags & 7) == 7
            rec = res.records
            assert rec.size == int(keep.sum())
            assert np.array_equal(rec["word"], o.word[keep]) and np.array_equal(rec["k"], o.k[keep])
            assert np.array_equal(rec["flags"] & 15, o.flags[keep])
            assert np.array_equal(rec["tm"], o.tm[keep])
            for gi in range(len(gens)):
                pos = res.positions[gi]
                assert np.array_equal(pos["contig"].astype(np.int32), o.contig[gi][keep])
                assert np.array_equal(pos["start"].astype(np.int64), o.start[gi][keep])
                assert np.array_equal(pos["strand"].astype(np.uint8), o.strand[gi][keep])


def test_genome_too_large_is_refused():
    from primerforge_b200 import _lib
    from primerforge_b200.engine import Engine, PfbError
    big = np.full((1 << 26) + 64, ord("A"), dtype=np.uint8)
    with Engine(device=0) as eng:
        eng.add_genome([big])
        with pytest.raises(PfbError) as exc:
            eng.upload()
        assert exc.value.code == _lib.PFB_ETOOLARGE


def test_identical_genomes_and_single_genome():
    from oracle import pf_oracle as po
    from primerforge_b200 import synth
    g = synth.make_genomes(1, 30_000, seed=555)[0]
    kw = dict(k_min=16, k_max=20, min_gc=40.0, max_gc=60.0, min_tm=55.0, max_tm=68.0, max_repeat=4, table_size=9999991)
    for gens in ([g], [g, g, g]):
        o = po.run(gens, po.default_params(**kw))
        res, _ = _run(gens, kw, False)
        assert np.array_equal(res.records["word"], o.word) and np.array_equal(res.records["flags"] & 15, o.flags)


def test_id_list_path_of_the_key_table(monkeypatch):
    """ids that do not fit the 24-bit inline format of the genomes-2..G key table go through the per-word id
    lists (normally only beyond 16.7 M keys); PFB200_AT_WIDE lowers the limit so that most keys take that path"""
    from primerforge_b200 import synth
    from primerforge_b200.engine import Engine
    gens = synth.make_genomes(4, 150_000, seed=23, snp_rate=0.003, n_inversions=2, n_contigs=2)

    def run():
        with Engine(device=0, prefilter=1) as eng:
            for g in gens:
                eng.add_genome(g.contigs)
            eng.run()
            res = eng.result()
            return res.records.copy(), [p.copy() for p in res.positions]
    rec_a, pos_a = run()
    monkeypatch.setenv("PFB200_AT_WIDE", "500")
    rec_b, pos_b = run()
    assert rec_a.size > 5000
    for f in ("word", "tm", "k", "flags", "gc_count"):
        assert np.array_equal(rec_a[f], rec_b[f])
    for a, b in zip(pos_a, pos_b):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("n_ranks", [2, 3, 8])
@pytest.mark.parametrize("n_contigs,n_frac", [(1, 0.0), (6, 0.0005)])
def test_shared_out_first_genome_pipeline_equals_oracle(n_ranks, n_contigs, n_frac):
    """the multi-GPU formulation of genome 1's own pipeline -- candidate windows listed per rank and applied by all,
    its scan in per-rank shares whose rows are packed to 32 bits and summed, polluter windows re-derived by every
    rank -- played by ONE context for n ranks in turn (PFB_OPT_EMULATE_RANKS: everything but the NCCL calls), so
    that a single-GPU box checks it against the oracle"""
    from oracle import pf_oracle as po
    from primerforge_b200 import _lib, synth
    from primerforge_b200.engine import Engine
    from tests.util import assert_equals_oracle
    gens = synth.make_genomes(4, 150_000, seed=520 + n_ranks, snp_rate=0.003, n_inversions=3, n_contigs=n_contigs, n_frac=n_frac)
    kw = dict(k_min=16, k_max=20, min_gc=40.0, max_gc=60.0, min_tm=55.0, max_tm=68.0, max_repeat=4,
              table_size=9999991 if n_contigs == 1 else 250007)       # the small table: collisions + junction pollution
    o = po.run(gens, po.default_params(**kw))
    with Engine(device=0, prefilter=1, **kw) as eng:
        eng.set_option(_lib.OPT_EMULATE_RANKS, n_ranks)
        for g in gens:
            eng.add_genome(g.contigs)
        eng.run()
        assert eng.timing()["sharded_first_genome"] == 1
        res = eng.result()
    assert_equals_oracle(res, o, True, len(gens))


def test_capacities_that_prove_too_small_repeat_the_run(monkeypatch):
    """no step of a run waits for a count: buffers have capacities, and a count that does not fit is found at the end,
    the capacity grown and the run repeated -- here forced with a tiny a-priori share, single and shared-out"""
    from oracle import pf_oracle as po
    from primerforge_b200 import _lib, synth
    from primerforge_b200.engine import Engine
    from tests.util import assert_equals_oracle
    gens = synth.make_genomes(3, 120_000, seed=530, snp_rate=0.003, n_inversions=2, n_contigs=2)
    kw = dict(k_min=16, k_max=20, min_gc=40.0, max_gc=60.0, min_tm=55.0, max_tm=68.0, max_repeat=4, table_size=9999991)
    o = po.run(gens, po.default_params(**kw))
    monkeypatch.setenv("PFB200_CAP_SHARE", "0.0005")
    for prefilter, emulate in ((0, 0), (1, 0), (1, 4)):
        with Engine(device=0, prefilter=prefilter, **kw) as eng:
            if emulate:
                eng.set_option(_lib.OPT_EMULATE_RANKS, emulate)
            for g in gens:
                eng.add_genome(g.contigs)
            eng.run()
            assert eng.timing()["n_reruns"] >= 1
            res = eng.result()
            assert_equals_oracle(res, o, bool(prefilter), len(gens))
            eng.upload()
            eng.compute()                              # the capacities have settled: no repeat the second time
            assert eng.timing()["n_reruns"] == res.timing["n_reruns"]
            assert_equals_oracle(eng.result(), o, bool(prefilter), len(gens))
