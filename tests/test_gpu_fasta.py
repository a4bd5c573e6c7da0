"""Device FASTA / GenBank ingest (pfb_add_fasta, pfb_add_genbank): the records the GPU parser finds and the stage
results computed from them must equal what the ORACLE side's reader gives -- `SeqIO.parse` of oracle/standins/Bio,
the module the reference's unmodified getCandidateKmers.py:93,212 runs against when the golden fixtures are made --
through pfb_add_genome, for well-formed files and for every irregularity the rules cover.  The product's own host
reader (primerforge_b200.seqio) is held to the same stand-in in tests/test_seqio.py."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

_STANDINS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "standins")


def _oracle_read(fn, fmt):
    """(ids, contigs) as the reference sees them: SeqRecord.id and str(rec.seq) of every record"""
    if _STANDINS not in sys.path:
        sys.path.insert(0, _STANDINS)
    from Bio import SeqIO
    recs = list(SeqIO.parse(fn, fmt))
    return [r.id for r in recs], [np.frombuffer(str(r.seq).encode("latin-1"), dtype=np.uint8).copy() for r in recs]


def _fasta_bytes(contigs, ids, width, rng=None, crlf=False, junk=b"", trailing_newline=True, spaces=False, tabs=False):
    out = bytearray(junk)
    nl = b"\r\n" if crlf else b"\n"
    for cid, arr in zip(ids, contigs):
        out += b">" + cid.encode() + b" some description > with a bracket" + nl
        seq = arr.tobytes()
        i = 0
        while i < len(seq):
            w = width if rng is None else int(rng.integers(1, 2 * width))
            line = seq[i:i + w]
            if spaces and len(line) > 4:
                line = line[:3] + b" " + line[3:]
            if tabs and len(line) > 4:
                # white space that line.rstrip() removes (runs of any length, also across a 32-byte step of the
                # parser) and a tab INSIDE a line, which stays a byte of the sequence
                r = (i // max(w, 1)) % 7
                if r == 0:
                    line = line + b"\t"
                elif r == 1:
                    line = line + b" \t \x0b\x0c" * 9
                elif r == 2:
                    line = line[:2] + b"\t" + line[2:]
                elif r == 3:
                    line = line[:2] + b"\t" + line[2:] + b"\t\t"
                elif r == 4:
                    line = b"\t" + line + b"\x1c\x1f "
            out += line + nl
            i += w
        if rng is not None and rng.random() < 0.3:
            out += nl                                     # blank line between records
    if not trailing_newline:
        while out and out[-1:] in (b"\n", b"\r"):
            out = out[:-1]
    return np.frombuffer(bytes(out), dtype=np.uint8).copy()


def _run(add, mode):
    from primerforge_b200.engine import Engine
    with Engine(device=0, prefilter=0) as eng:
        add(eng)
        if mode == "pipelined":
            eng.run()
        else:
            eng.upload()
            eng.compute()
        res = eng.result()
        tables = [eng.contig_table(g) for g in range(eng.n_genomes)]
        info = [eng.genome_info(g) for g in range(eng.n_genomes)]
    return res, tables, info


def _compare(images, mode):
    from primerforge_b200 import seqio
    import tempfile
    host = []
    with tempfile.TemporaryDirectory() as d:
        for i, img in enumerate(images):
            fn = os.path.join(d, f"g{i}.fasta")
            img.tofile(fn)
            host.append(_oracle_read(fn, "fasta"))
    res_f, tables, info = _run(lambda e: [e.add_fasta(img) for img in images], mode)
    res_h, _, _ = _run(lambda e: [e.add_genome(c) for _ids, c in host], mode)
    for g, (ids, contigs) in enumerate(host):
        hdr, ln = tables[g]
        assert info[g] == (len(contigs), sum(c.size for c in contigs))
        assert ln.tolist() == [c.size for c in contigs]
        assert seqio.fasta_ids(images[g], hdr) == ids
        assert all(images[g][int(h)] == ord(">") for h in hdr)
    assert res_f.records.size == res_h.records.size and res_f.records.size > 0
    for f in ("word", "tm", "gc_count", "k", "flags"):
        assert np.array_equal(res_f.records[f], res_h.records[f]), f
    for pf, ph in zip(res_f.positions, res_h.positions):
        assert np.array_equal(pf, ph)


@pytest.mark.parametrize("mode", ["sync", "pipelined"])
def test_regular_fasta(mode):
    from primerforge_b200 import synth
    gens = synth.make_genomes(4, 120_000, seed=77, snp_rate=0.003, n_inversions=2, n_contigs=4)
    images = [_fasta_bytes(g.contigs, g.contig_ids, 80) for g in gens]
    _compare(images, mode)


@pytest.mark.parametrize("mode", ["sync", "pipelined"])
@pytest.mark.parametrize("variant", ["crlf", "junk", "unwrapped", "ragged", "no_final_newline", "spaces", "long_title", "tiny_records",
                                     "tabs", "tabs_crlf", "tabs_no_final_newline"])
def test_irregular_fasta(variant, mode):
    from primerforge_b200 import synth
    import zlib
    rng = np.random.default_rng(zlib.crc32(variant.encode()) % 1000)
    gens = synth.make_genomes(3, 60_000, seed=91, snp_rate=0.003, n_inversions=1, n_contigs=3)
    images = []
    for g in gens:
        kw = {}
        width = 70
        ids = list(g.contig_ids)
        contigs = list(g.contigs)
        if variant == "crlf":
            kw["crlf"] = True
        elif variant == "junk":
            kw["junk"] = b"; a comment line\n\nsome text before the first record\n"
        elif variant == "unwrapped":
            width = 10_000_000                            # one line per record: lines far longer than a tile
        elif variant == "ragged":
            kw["rng"] = rng
        elif variant == "no_final_newline":
            kw["trailing_newline"] = False
        elif variant == "spaces":
            kw["spaces"] = True
        elif variant.startswith("tabs"):
            kw["tabs"] = True
            kw["crlf"] = variant == "tabs_crlf"
            kw["trailing_newline"] = variant != "tabs_no_final_newline"
        elif variant == "long_title":
            ids = [cid + "_" + "x" * 9000 for cid in ids]  # a title line longer than two tiles
        elif variant == "tiny_records":
            extra = [np.frombuffer(b"ACGT", dtype=np.uint8), np.zeros(0, dtype=np.uint8), np.frombuffer(b"GGCCNNacgt", dtype=np.uint8)]
            contigs = contigs[:1] + extra + contigs[1:]
            ids = ids[:1] + ["t1", "empty", "t3"] + ids[1:]
        images.append(_fasta_bytes(contigs, ids, width, **kw))
    _compare(images, mode)


def test_fasta_errors():
    from primerforge_b200.engine import Engine, PfbError
    from primerforge_b200 import synth
    gens = synth.make_genomes(2, 20_000, seed=5)
    with Engine(device=0) as eng:
        eng.add_fasta(np.frombuffer(b"ACGTACGTACGT\nACGT\n", dtype=np.uint8).copy())     # no record at all
        with pytest.raises(PfbError):
            eng.run()
    with Engine(device=0) as eng:                                                       # the two input kinds do not mix
        eng.add_genome(gens[0].contigs)
        eng.add_fasta(_fasta_bytes(gens[1].contigs, gens[1].contig_ids, 60))
        with pytest.raises(PfbError):
            eng.run()
    with Engine(device=0) as eng:
        with pytest.raises(PfbError):
            eng.add_fasta(np.zeros(0, dtype=np.uint8))


def test_many_records_grow_the_table():
    """more records than the initial record table holds: the gather pass is repeated with a larger one"""
    from primerforge_b200.engine import Engine
    rng = np.random.default_rng(3)
    n_rec = 70_000                                       # initial capacity is 65536 + bytes / 256
    body = bytearray()
    for i in range(n_rec):
        body += b">r%d\n" % i + bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=2)) + b"\n"
    img = np.frombuffer(bytes(body), dtype=np.uint8).copy()
    with Engine(device=0) as eng:
        eng.add_fasta(img)
        eng.upload()
        nc, nb = eng.genome_info(0)
        hdr, ln = eng.contig_table(0)
    assert nc == n_rec and nb == 2 * n_rec and np.all(ln == 2)
    assert np.all(img[hdr.astype(np.int64)] == ord(">"))


# ---------------------------------------------------------------------------------------------- GenBank
def _genbank_bytes(contigs, ids, version=True, crlf=False, junk=b"", upper=False, long_feature=False):
    nl = "\r\n" if crlf else "\n"
    out = [junk.decode()]
    for cid, arr in zip(ids, contigs):
        seq = arr.tobytes().decode()
        seq = seq if upper else seq.lower()
        out.append(f"LOCUS       {cid}  {len(seq)} bp    DNA     linear   BCT 01-JAN-2000{nl}")
        out.append(f"DEFINITION  synthetic record, the word ORIGIN and // inside a line must not matter.{nl}")
        out.append(f"ACCESSION   {cid}{nl}")
        if version:
            out.append(f"VERSION     {cid}.1{nl}")
        out.append(f"FEATURES             Location/Qualifiers{nl}     source          1..{len(seq)}{nl}")
        out.append(f"                     /organism=\"Synthetic acgt\"{nl}")
        if long_feature:
            out.append("                     /translation=\"" + "MKV" * 4000 + f"\"{nl}")     # a line longer than two tiles
        out.append(f"ORIGIN      {nl}")
        for i in range(0, len(seq), 60):
            line = seq[i:i + 60]
            groups = " ".join(line[j:j + 10] for j in range(0, len(line), 10))
            out.append(f"{i + 1:9d} {groups}{nl}")
        out.append(f"//{nl}")
    return np.frombuffer("".join(out).encode(), dtype=np.uint8).copy()


def _compare_genbank(images, mode):
    from primerforge_b200 import seqio
    import tempfile
    host = []
    with tempfile.TemporaryDirectory() as d:
        for i, img in enumerate(images):
            fn = os.path.join(d, f"g{i}.gbk")
            img.tofile(fn)
            host.append(_oracle_read(fn, "genbank"))
    res_f, tables, info = _run(lambda e: [e.add_genbank(img) for img in images], mode)
    res_h, _, _ = _run(lambda e: [e.add_genome(c) for _ids, c in host], mode)
    for g, (ids, contigs) in enumerate(host):
        hdr, ln = tables[g]
        assert info[g] == (len(contigs), sum(c.size for c in contigs))
        assert ln.tolist() == [c.size for c in contigs]
        assert seqio.genbank_ids(images[g], hdr) == ids
        assert all(images[g][int(h):int(h) + 5].tobytes() == b"LOCUS" for h in hdr)
    assert res_f.records.size == res_h.records.size and res_f.records.size > 0
    for f in ("word", "tm", "gc_count", "k", "flags"):
        assert np.array_equal(res_f.records[f], res_h.records[f]), f
    for pf, ph in zip(res_f.positions, res_h.positions):
        assert np.array_equal(pf, ph)


@pytest.mark.parametrize("mode", ["sync", "pipelined"])
@pytest.mark.parametrize("variant", ["plain", "no_version", "crlf", "junk", "upper", "long_feature", "empty_records"])
def test_genbank_device_parser(variant, mode):
    from primerforge_b200 import synth
    gens = synth.make_genomes(3, 60_000, seed=93, snp_rate=0.003, n_inversions=1, n_contigs=3)
    images = []
    for g in gens:
        kw = {}
        ids = list(g.contig_ids)
        contigs = list(g.contigs)
        if variant == "no_version":
            kw["version"] = False
        elif variant == "crlf":
            kw["crlf"] = True
        elif variant == "junk":
            kw["junk"] = b"some text\nORIGIN\n        1 acgtacgt\n//\nmore text\n"     # sequence-looking text before any LOCUS
        elif variant == "upper":
            kw["upper"] = True
        elif variant == "long_feature":
            kw["long_feature"] = True
        elif variant == "empty_records":
            contigs = contigs[:1] + [np.zeros(0, dtype=np.uint8), np.frombuffer(b"ACGTNNACGT", dtype=np.uint8)] + contigs[1:]
            ids = ids[:1] + ["empty", "tiny"] + ids[1:]
        images.append(_genbank_bytes(contigs, ids, **kw))
    _compare_genbank(images, mode)


def test_drop_in_entry_point_genbank(tmp_path, capsys):
    """_getAllCandidateKmers with params.format = 'genbank' (the reference's default): device ingest and host
    reader give the same candidates"""
    from types import SimpleNamespace
    from primerforge_b200 import synth
    from primerforge_b200.candidates import _getAllCandidateKmers
    gens = synth.make_genomes(3, 80_000, seed=17, snp_rate=0.003, n_inversions=2, n_contigs=2)
    paths = []
    for g in gens:
        fn = str(tmp_path / g.name.replace(".fasta", ".gbk"))
        _genbank_bytes(g.contigs, g.contig_ids).tofile(fn)
        paths.append(fn)
    log = SimpleNamespace(rename=lambda *a: None, info=lambda *a: None, debug=lambda *a: None, error=lambda *a: None)

    def params():
        return SimpleNamespace(ingroupFns=paths, format="genbank", minLen=16, maxLen=20, minGc=40.0, maxGc=60.0, minTm=55.0,
                               maxTm=68.0, maxRepeatLen=4, tempTolerance=5.0, numThreads=1, p3_mvConc=50.0, p3_dvConc=1.5,
                               p3_dntpConc=0.6, p3_dnaConc=50.0, p3_tempC=37.0, p3_maxLoop=30, log=log,
                               pickles={1: "sharedKmers.p", 2: "candidates.p"}, dumpObj=lambda *a, **k: None, loadObj=lambda *a: None)
    out_dev = _getAllCandidateKmers(params(), False, thal="off")
    os.environ["PFB200_HOST_PARSE"] = "1"
    try:
        out_host = _getAllCandidateKmers(params(), False, thal="off")
    finally:
        del os.environ["PFB200_HOST_PARSE"]
    capsys.readouterr()
    assert list(out_dev.keys()) == list(out_host.keys()) == [os.path.basename(p) for p in paths]
    for name in out_dev:
        assert list(out_dev[name].keys()) == list(out_host[name].keys())
        for contig in out_dev[name]:
            a = [(str(p), p.start, p.end, p.strand, p.Tm, p.gcPer) for p in out_dev[name][contig]]
            b = [(str(p), p.start, p.end, p.strand, p.Tm, p.gcPer) for p in out_host[name][contig]]
            assert a == b and len(a) > 0


def test_many_tiny_records_outgrow_the_word_estimate():
    """pipelined run: the word arrays are sized from the file sizes before genomes 2..G are parsed; thousands
    of records shorter than a word break that estimate and the arrays must grow without losing genome 1"""
    from primerforge_b200 import synth
    gens = synth.make_genomes(3, 60_000, seed=29, snp_rate=0.003, n_inversions=1, n_contigs=2)
    rng = np.random.default_rng(8)
    images = []
    for gi, g in enumerate(gens):
        contigs, ids = list(g.contigs), list(g.contig_ids)
        if gi:
            for i in range(20_000):
                contigs.append(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=8))
                ids.append(f"t{i}")
        images.append(_fasta_bytes(contigs, ids, 70))
    _compare(images, "pipelined")
    _compare(images, "sync")
