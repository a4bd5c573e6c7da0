"""oracle/pf_oracle.c (CPU restatement) against the committed reference outputs (tests/golden,
produced by oracle/make_golden.py from the reference's unmodified bin/getCandidateKmers.py)."""
import numpy as np
import pytest

from oracle import pf_oracle as po
from tests.util import candidate_set_from_arrays, golden_candidate_set, golden_cases, load_golden


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_matches_reference(name):
    gold = load_golden(name)
    p = po.default_params(k_min=gold.k_min, k_max=gold.k_max, min_gc=gold.min_gc, max_gc=gold.max_gc,
                          min_tm=gold.min_tm, max_tm=gold.max_tm, max_repeat=gold.max_repeat,
                          thal_mode=1 if gold.thal == "pseudo" else 0, temp_tolerance=gold.temp_tolerance,
                          table_size=gold.table_size)
    o = po.run(gold.genomes, p)
    kmers = o.kmer_strings()
    # (1) shared set and (2) dict insertion order  (SURVEY.md 8c parity definition)
    assert kmers == gold.shared
    # coordinates / strands in every genome
    for gi in range(len(gold.genomes)):
        assert np.array_equal(o.contig[gi], gold.z[f"g{gi}_sh_contig"])
        assert np.array_equal(o.start[gi], gold.z[f"g{gi}_sh_start"])
        assert np.array_equal(o.strand[gi], gold.z[f"g{gi}_sh_strand"])
    # (4) candidate set per genome
    picked = (o.flags & 8) != 0
    for gi in range(len(gold.genomes)):
        mine = candidate_set_from_arrays(kmers, o.k, picked, o.contig[gi], o.start[gi], o.strand[gi])
        assert mine == golden_candidate_set(gold, gi)
    # (3) Tm / GC of the candidates: bit-exact here (same IEEE operations), 0.01 is the contract
    z = gold.z
    idx = z["g0_cand_kmer"]
    assert np.max(np.abs(o.tm[idx] - z["g0_cand_tm"]), initial=0.0) <= 1e-9
    gc = o.gc_count[idx] / o.k[idx] * 100
    assert np.max(np.abs(gc - z["g0_cand_gc"]), initial=0.0) <= 1e-9
