"""``SeqIO.parse(fn, 'fasta'|'genbank')`` -- record id rules as in biopython:
FASTA id = header up to the first whitespace, sequence case preserved;
GenBank id = ACCESSION.VERSION from the VERSION line (else LOCUS name),
sequence upper-cased."""
from .Seq import Seq
from .SeqRecord import SeqRecord


def _parse_fasta(handle):
    title = None
    chunks = []
    for line in handle:
        if line.startswith(">"):
            if title is not None:
                yield _fasta_record(title, chunks)
            title = line[1:].rstrip()
            chunks = []
        elif title is not None:
            chunks.append(line.rstrip())        # SimpleFastaParser: trailing white space only
    if title is not None:
        yield _fasta_record(title, chunks)


def _fasta_record(title, chunks):
    parts = title.split(None, 1)
    rid = parts[0] if parts else ""
    return SeqRecord(Seq("".join(chunks).replace(" ", "").replace("\r", "")), id=rid, name=rid, description=title)


def _parse_genbank(handle):
    locus = None
    version = None
    in_origin = False
    chunks = []
    for line in handle:
        if line.startswith("LOCUS"):
            f = line.split()
            locus = f[1] if len(f) > 1 else "<unknown id>"
            version = None
            in_origin = False
            chunks = []
        elif line.startswith("VERSION"):
            f = line.split()
            if len(f) > 1:
                version = f[1]
        elif line.startswith("ORIGIN"):
            in_origin = True
        elif line.startswith("//"):
            if locus is not None:
                rid = version if version else locus
                yield SeqRecord(Seq("".join(chunks).upper()), id=rid, name=locus)
            locus = None
            in_origin = False
            chunks = []
        elif in_origin:
            chunks.append("".join(ch for ch in line if ch.isalpha()))


def parse(source, fmt):
    fmt = fmt.lower()
    if fmt not in ("fasta", "genbank", "gb"):
        raise ValueError(f"unsupported format {fmt!r}")
    own = isinstance(source, (str, bytes)) or hasattr(source, "__fspath__")
    handle = open(source, "r") if own else source
    try:
        gen = _parse_fasta(handle) if fmt == "fasta" else _parse_genbank(handle)
        for rec in gen:
            yield rec
    finally:
        if own:
            handle.close()
