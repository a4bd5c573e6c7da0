"""Primer record mirror: the concrete coordinates of the reference's own unit test
(/root/reference/bin/unit_tests/primer_test.py:11-23, 96-112, 208-269)."""
import pytest

from primerforge_b200.primer import _LocalPrimer as Primer, make_primer, oligo_tm

FWD_SEQ = "agagtttgatcctggctcag"
REV_SEQ = "ggttaccttgttacgactt"


def test_start_end_flip_on_minus_strand():
    f1 = Primer(FWD_SEQ, "contig 1", 4, len(FWD_SEQ), "+")
    f2 = Primer(FWD_SEQ, "contig 2", 8, len(FWD_SEQ), "-")
    assert (f1.start, f1.end) == (4, 23)
    assert (f2.start, f2.end) == (27, 8)
    r1 = Primer(REV_SEQ, "contig 1", 300, len(REV_SEQ), "-")
    r2 = Primer(REV_SEQ, "contig 2", 666, len(REV_SEQ), "+")
    assert (r1.start, r1.end) == (318, 300)
    assert (r2.start, r2.end) == (666, 684)
    assert len(f1) == 20 and len(r1) == 19
    assert str(f1) == FWD_SEQ.upper() and f1 == FWD_SEQ.upper() and hash(f1) == hash(FWD_SEQ.upper())


def test_gc_and_tm():
    f1 = Primer(FWD_SEQ, "c", 0, len(FWD_SEQ), "+")
    assert f1.gcPer == sum(c in "GC" for c in FWD_SEQ.upper()) / len(FWD_SEQ) * 100
    assert f1.Tm == oligo_tm(FWD_SEQ)
    assert round(oligo_tm("AGAAGCAAAGGATATGGGAACAAC"), 1) == 57.1     # README.md:375


def test_reverse_complement_swaps_fields_and_keeps_tm_gc():
    f1 = Primer(FWD_SEQ, "contig 1", 4, len(FWD_SEQ), "+")
    f1.hairpinTm, f1.rcHairpin, f1.homodimerTm, f1.rcHomodimer = 1.0, 2.0, 3.0, 4.0
    rc = f1.reverseComplement()
    assert rc.strand == "-" and (rc.start, rc.end) == (23, 4)
    assert str(rc) == "CTGAGCCAGGATCAAACTCT"
    assert (rc.hairpinTm, rc.rcHairpin, rc.homodimerTm, rc.rcHomodimer) == (2.0, 1.0, 4.0, 3.0)
    assert rc.Tm == pytest.approx(f1.Tm, abs=1e-9) and rc.gcPer == f1.gcPer
    back = rc.reverseComplement()
    assert back.strand == "+" and (back.start, back.end) == (4, 23) and back == f1


def test_invalid_strand_and_compare():
    with pytest.raises(ValueError):
        Primer(FWD_SEQ, "c", 0, 20, "x")
    with pytest.raises(ValueError):
        Primer(FWD_SEQ, "c", 0, 20, "+") == 3


def test_make_primer_matches_constructor():
    a = Primer(FWD_SEQ, "c1", 8, 20, "-")
    b = make_primer(Primer, FWD_SEQ.upper(), "c1", 8, 20, True, a.Tm, a.gcPer)
    for attr in ("seq", "contig", "start", "end", "strand", "Tm", "gcPer", "hairpinTm", "_Primer__length"):
        assert getattr(a, attr) == getattr(b, attr), attr
    import pickle
    c = pickle.loads(pickle.dumps(b))
    assert c == b and c.start == b.start and len(c) == 20
