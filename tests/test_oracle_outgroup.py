"""oracle/pf_oracle_outgroup.c (CPU restatement of the outgroup screen) against the committed outputs of the
reference's unmodified bin/removeOutgroupPrimers.py (tests/golden/og_*.npz, oracle/make_golden_outgroup.py)."""
import numpy as np
import pytest

from oracle import pf_oracle as po
from tests.util import load_og, og_cases


@pytest.mark.parametrize("name", og_cases())
def test_oracle_matches_reference(name):
    g = load_og(name)
    o = po.run_outgroup(g.keys, g.pair_fwd, g.pair_rev, g.bad_lo, g.bad_hi, g.genomes)
    # __extractKmerData's start lists (removeOutgroupPrimers.py:34-62), every contig and strand
    mine = np.array(sorted(map(tuple, o.hits.tolist())), dtype=np.int64).reshape(-1, 5)
    assert np.array_equal(mine, g.hits)
    # which pairs survive the loop at :236-291
    kept = o.removed_at == 0xFFFFFFFF
    if g.error:
        assert not kept.any()          # the reference raised ERR_MSG (:301-303)
        return
    assert np.array_equal(kept, g.kept)
    # the (contig, size) sets behind the Product objects of the kept pairs (:125-178, 289-291)
    assert np.array_equal(o.products.astype(np.int64), g.products)
    # the debug lines (:240-247, 294): pairs removed per outgroup genome
    remaining = kept.size
    for gi, name_g in enumerate(g.names):
        gone = int(np.count_nonzero(o.removed_at == gi))
        remaining -= gone
        assert g.debug[gi] == f"removed {gone} pairs after processing {name_g} ({remaining} pairs remaining)"


def test_fixtures_cover_the_interesting_cases():
    multi = load_og("og_multi")
    pg = multi.products[:, :2]
    _, counts = np.unique(pg, axis=0, return_counts=True)
    assert (counts > 1).any()                                  # several products for one (pair, genome)
    assert multi.kept.any() and not multi.kept.all()
    assert any(c.size < 16 for gen in multi.genomes for c in gen.contigs)      # contigs shorter than a primer
    assert any((c == ord("N")).any() for gen in multi.genomes for c in gen.contigs)
    assert any(np.isin(c, np.frombuffer(b"acgt", dtype=np.uint8)).any() for gen in multi.genomes for c in gen.contigs)
    assert load_og("og_allgone").error
    assert {len(k) for k in load_og("og_widek").keys} - set(range(16, 21))    # lengths outside 16..20
