"""Host side of the drop-in (grouping + lazy secondary-structure screening + record building),
driven by the oracle's arrays in place of the GPU's so that it runs on a CPU-only box."""
import zlib
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import pf_oracle as po
from primerforge_b200 import _lib
from primerforge_b200.candidates import _word_starts, screen_groups
from primerforge_b200.lazy import Columns, build_lazy_output
from primerforge_b200.engine import words_to_strings
from tests.util import golden_candidate_set, load_golden


def _pseudo_thal(seq, rc):
    def one(s, salt):
        return ((zlib.crc32(s.encode("ascii")) ^ ((salt * 0x9E3779B1) & 0xFFFFFFFF)) % 6400) / 100.0
    return one(seq, 1), one(rc, 1), one(seq, 2), one(rc, 2)


def _as_stage_result(o, n_genomes):
    rec = np.zeros(o.n_shared, dtype=_lib.RECORD_DTYPE)
    rec["word"], rec["k"], rec["tm"], rec["gc_count"] = o.word, o.k, o.tm, o.gc_count
    rec["flags"] = o.flags & 7
    pos = []
    for g in range(n_genomes):
        p = np.zeros(o.n_shared, dtype=_lib.POS_DTYPE)
        p["contig"], p["start"], p["strand"] = o.contig[g], o.start[g], o.strand[g]
        pos.append(p)
    return rec, pos


@pytest.mark.parametrize("name", ["collide_multi", "widek", "tightfilters", "basic"])
def test_screening_and_output_equal_reference(name):
    gold = load_golden(name)
    par = po.default_params(k_min=gold.k_min, k_max=gold.k_max, min_gc=gold.min_gc, max_gc=gold.max_gc, min_tm=gold.min_tm,
                            max_tm=gold.max_tm, max_repeat=gold.max_repeat, table_size=gold.table_size)
    o = po.run(gold.genomes, par)
    rec, pos = _as_stage_result(o, len(gold.genomes))
    kmers = words_to_strings(rec["word"], rec["k"])
    assert kmers == gold.shared
    thal = _pseudo_thal if gold.thal == "pseudo" else (lambda s, r: (0.0, 0.0, 0.0, 0.0))
    picked, values = screen_groups(rec, pos, kmers, thal, gold.min_tm - gold.temp_tolerance)
    sel = np.flatnonzero(picked)
    # the picked records as the columns the drop-in builds from the GPU's per-id view
    wstarts = [_word_starts([c.size for c in g.contigs]) for g in gold.genomes]
    enc = []
    for g in range(len(gold.genomes)):
        lp = pos[g]["start"][sel].astype(np.int64) + (32 * wstarts[g][pos[g]["contig"][sel]] if wstarts[g] is not None else 0)
        enc.append(lp.astype(np.uint32) | (pos[g]["strand"][sel].astype(np.uint32) << np.uint32(31)))
    cols = Columns(gold.names, [g.contig_ids for g in gold.genomes], wstarts, rec["word"][sel], rec["k"][sel], rec["tm"][sel],
                   rec["gc_count"][sel], enc, thal={i: values[int(r)] for i, r in enumerate(sel)})
    out = build_lazy_output(cols)
    for gi, (name_g, gen) in enumerate(zip(gold.names, gold.genomes)):
        cidx = {c: i for i, c in enumerate(gen.contig_ids)}
        mine = set()
        for contig, plist in out[name_g].items():
            keys = [min(p.start, p.end) for p in plist]
            assert keys == sorted(keys)                                   # getCandidateKmers.py:602
            for p in plist:
                mine.add((str(p), cidx[contig], p.start, p.end, int(p.strand == "-")))
                assert p.hairpinTm is not None and len(p) == len(str(p))
        assert mine == golden_candidate_set(gold, gi)
    # Tm / gcPer carried by the records equal the reference's (0.01 contract)
    z = gold.z
    first = {str(p): p for plist in out[gold.names[0]].values() for p in plist}
    for i, km in enumerate(z["g0_cand_kmer"][:500]):
        p = first[gold.shared[int(km)]]
        assert abs(p.Tm - z["g0_cand_tm"][i]) <= 0.01 and abs(p.gcPer - z["g0_cand_gc"][i]) <= 0.01


def test_words_to_strings():
    assert words_to_strings(np.array([0b00011011], dtype=np.uint64), np.array([4])) == ["ATCG"]
    assert words_to_strings(np.array([], dtype=np.uint64), np.array([])) == []
    w = po.default_params()
    assert words_to_strings(np.array([3, 0], dtype=np.uint64), np.array([2, 3])) == ["AG", "AAA"]
