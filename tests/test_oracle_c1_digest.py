"""BASELINE config 1 at FULL size (3 x 1 Mbp, k 16-20, the real 9 999 991-bin table): the CPU oracle reproduces
the result of the reference's unmodified Python (run once in the build container against oracle/standins by
tools/time_reference_c1.py, committed as digests -- oracle/digest.py -- in tests/golden/c1_reference_digest.json):
the shared k-mers in dict insertion order, their coordinates and strands in every genome, and every genome's
candidate set with Primer.start / Primer.end / Tm."""
import json
import os

from oracle import digest, pf_oracle as po
from primerforge_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_reference_digest.json")


def test_oracle_reproduces_the_reference_at_full_config1_size():
    with open(GOLD) as fh:
        gold = json.load(fh)
    gens, cfg = synth.make_config(1)
    assert len(gens) == gold["genomes"] and gens[0].n_bases == gold["genome_bp"]
    o = po.run(gens, po.default_params(k_min=cfg["k_min"], k_max=cfg["k_max"], max_repeat=cfg["max_repeats"]))
    mine = digest.result_digests(o.word, o.k, o.flags, o.tm, o.contig, o.start, o.strand)
    assert mine["n_shared"] == gold["n_shared"] and mine["n_candidates"] == gold["n_candidates"]
    assert mine["shared"] == gold["shared"]
    assert mine["candidates"] == gold["candidates"]
