"""Known-answer tests the reference itself publishes for this path."""
import pytest

from oracle import pf_oracle as po

# /root/reference/README.md:375-377 -- six primers with Tm (0.1 degC) and GC (0.1 %)
README_ROWS = [
    ("AGAAGCAAAGGATATGGGAACAAC", 57.1, 41.7),
    ("AAATCAACACCCTCAATAAGCTCC", 57.1, 41.7),
    ("TCCATCTAATGAAGATCAACCAGG", 55.9, 41.7),
    ("CCCTAATTGTGATGAGTTGACAAC", 55.9, 41.7),
    ("CATCAGCTGTTGTAAATAACCCAC", 56.2, 41.7),
    ("GTGGAGCTATGAAACCAATATCAG", 55.3, 41.7),
]


@pytest.mark.parametrize("seq,tm,gc", README_ROWS)
def test_readme_tm_rows(seq, tm, gc):
    assert round(po.calc_tm(seq), 1) == tm
    assert round(sum(c in "GC" for c in seq) / len(seq) * 100, 1) == gc


def test_standin_tm_equals_oracle_tm():
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "standins"))
    import primer3
    for seq, _, _ in README_ROWS + [("AGAGTTTGATCCTGGCTCAG", 0, 0), ("GGTTACCTTGTTACGACTT", 0, 0), ("GAATTCGAATTC", 0, 0)]:
        assert primer3.calc_tm(seq) == po.calc_tm(seq)


def test_invalid_oligo():
    assert po.calc_tm("ACGTNACGTACGTACG") < -999999
