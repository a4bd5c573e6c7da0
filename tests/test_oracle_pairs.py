"""oracle/pf_oracle_pairs.c (CPU restatement of the pair search) against the committed outputs of the reference's
unmodified bin/getPrimerPairs.py (tests/golden/pp_*.npz, oracle/make_golden_pairs.py)."""
import numpy as np
import pytest

from oracle import pf_oracle as po
from tests.util import expected_products, load_pp, pp_cases


@pytest.mark.parametrize("name", pp_cases())
def test_oracle_matches_reference(name):
    f = load_pp(name)
    o = po.run_pairs(f.word, f.k, f.tm, f.genomes, max_bin=f.max_bin, min_primer_len=f.min_len, min_pcr=f.min_pcr,
                     max_pcr=f.max_pcr, max_tm_diff=f.max_tm_diff)
    kept = o.kept
    # the pairs dict: keys in insertion order (getPrimerPairs.py:396-398, 459-462)
    assert np.array_equal(o.pairs[kept, 0], f.pair_fwd) and np.array_equal(o.pairs[kept, 1], f.pair_rev)
    # every field of every Product (contig, size, fbin, rbin, fStart, fEnd, fStrand, rStart, rEnd, rStrand)
    assert np.array_equal(expected_products(f, o.pairs, o.shared, kept), f.products)


def test_fixtures_cover_the_interesting_cases():
    f = load_pp("pp_contigs")
    o = po.run_pairs(f.word, f.k, f.tm, f.genomes, max_bin=f.max_bin, min_primer_len=f.min_len, min_pcr=f.min_pcr,
                     max_pcr=f.max_pcr, max_tm_diff=f.max_tm_diff)
    assert o.reduced.any()                                   # __reduceBinSize removed something
    assert not o.kept.all() and o.kept.any()                 # pairs of the first genome that some other genome rejects
    assert any(g[3].any() for g in f.genomes[1:])            # '-' strand candidates
    assert (f.products[:, 1:, 6] == 1).any()                 # ... that end up in kept pairs
