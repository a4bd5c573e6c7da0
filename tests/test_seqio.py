import os
import sys

import numpy as np

from primerforge_b200 import seqio, synth

STANDINS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "standins")


def test_fasta_reader_matches_biopython_rules(tmp_path):
    gens = synth.make_genomes(2, 5000, seed=9, n_contigs=4)
    paths = synth.write_fasta(gens, str(tmp_path))
    sys.path.insert(0, STANDINS)
    from Bio import SeqIO
    for gen, path in zip(gens, paths):
        ids, contigs = seqio.read_genome(path, "fasta")
        ref = list(SeqIO.parse(path, "fasta"))
        assert ids == [r.id for r in ref] == gen.contig_ids
        for c, r, g in zip(contigs, ref, gen.contigs):
            assert c.tobytes().decode() == str(r.seq) and np.array_equal(c, g)


def test_fasta_edge_cases(tmp_path):
    p = tmp_path / "x.fasta"
    p.write_bytes(b">c1 desc here\r\nACGT\r\nAC\r\n>c2\n\n>c3\tz\nGG>GT\nTT")
    ids, contigs = seqio.read_fasta(str(p))
    assert ids == ["c1", "c2", "c3"]
    assert [c.tobytes() for c in contigs] == [b"ACGTAC", b"", b"GG>GTTT"]


def test_genbank_reader(tmp_path):
    p = tmp_path / "x.gbk"
    p.write_text("LOCUS       AB000001  12 bp    DNA\nVERSION     AB000001.2\nORIGIN\n        1 acgtacgtac gt\n//\n"
                 "LOCUS       NOVERSION  4 bp\nORIGIN\n        1 ttga\n//\n")
    ids, contigs = seqio.read_genome(str(p), "genbank")
    assert ids == ["AB000001.2", "NOVERSION"]
    assert [c.tobytes() for c in contigs] == [b"ACGTACGTACGT", b"TTGA"]


def test_ids_from_header_offsets():
    """the host half of the device ingest: record ids from the offsets the GPU parser reports"""
    fa = np.frombuffer(b"junk\n>c1 desc here\r\nACGT\n>c2\n\n>c3\tz\nGG", dtype=np.uint8)
    assert seqio.fasta_ids(fa, [5, 25, 30]) == ["c1", "c2", "c3"]
    gb = ("LOCUS       AB000001  12 bp    DNA\nDEFINITION  x\nVERSION     AB000001.2\nORIGIN\n        1 acgtacgtac gt\n//\n"
          "LOCUS       NOVERSION  4 bp\nORIGIN\n        1 ttga\n//\n").encode()
    arr = np.frombuffer(gb, dtype=np.uint8)
    offs = [0, gb.index(b"LOCUS       NOVERSION")]
    assert seqio.genbank_ids(arr, offs) == ["AB000001.2", "NOVERSION"]
    assert seqio.file_ids(arr, offs, "genbank") == ["AB000001.2", "NOVERSION"]


import pytest  # noqa: E402


@pytest.mark.parametrize("variant", ["crlf", "junk", "ragged", "no_final_newline", "spaces", "tabs", "tabs_crlf", "tabs_no_final_newline"])
def test_host_reader_equals_the_oracle_side_reader_on_irregular_files(variant, tmp_path):
    """primerforge_b200.seqio.read_fasta against oracle/standins/Bio (the reader the reference runs on when the golden
    fixtures are made): ids and bytes of every record, for every irregularity the device parser is tested on"""
    from tests.test_gpu_fasta import _fasta_bytes, _oracle_read
    gens = synth.make_genomes(2, 20_000, seed=91, n_contigs=3)
    kw = {}
    if variant == "crlf":
        kw["crlf"] = True
    elif variant == "junk":
        kw["junk"] = b"; a comment\n\nsome text before the first record\n"
    elif variant == "ragged":
        kw["rng"] = np.random.default_rng(1)
    elif variant == "no_final_newline":
        kw["trailing_newline"] = False
    elif variant == "spaces":
        kw["spaces"] = True
    else:
        kw.update(tabs=True, crlf=variant == "tabs_crlf", trailing_newline=variant != "tabs_no_final_newline")
    for g in gens:
        fn = str(tmp_path / f"{g.name}")
        _fasta_bytes(g.contigs, g.contig_ids, 70, **kw).tofile(fn)
        ids_o, contigs_o = _oracle_read(fn, "fasta")
        ids_h, contigs_h = seqio.read_fasta(fn)
        assert ids_o == ids_h == g.contig_ids
        assert all(np.array_equal(a, b) for a, b in zip(contigs_o, contigs_h))
        if variant.startswith("tabs"):
            assert any((c == 9).any() for c in contigs_h)      # a tab inside a line is a byte of the sequence
        else:
            assert all(np.array_equal(a, b) for a, b in zip(contigs_h, g.contigs))


def test_genbank_reader_equals_the_oracle_side_reader(tmp_path):
    from tests.test_gpu_fasta import _genbank_bytes, _oracle_read
    gens = synth.make_genomes(2, 20_000, seed=92, n_contigs=3)
    for kw in ({}, {"version": False}, {"crlf": True}, {"upper": True}, {"long_feature": True}):
        for g in gens:
            fn = str(tmp_path / "x.gbk")
            _genbank_bytes(g.contigs, g.contig_ids, **kw).tofile(fn)
            ids_o, contigs_o = _oracle_read(fn, "genbank")
            ids_h, contigs_h = seqio.read_genbank(fn)
            assert ids_o == ids_h
            assert all(np.array_equal(a, b) for a, b in zip(contigs_o, contigs_h))
