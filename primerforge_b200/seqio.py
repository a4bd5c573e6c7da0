"""Ingroup file readers: FASTA / GenBank -> contig ids + ASCII base arrays.

Takes the place of Bio.SeqIO.parse at /root/reference/bin/getCandidateKmers.py:93,212 (biopython
is not a dependency of this package).  Record id rules follow biopython: FASTA id = header up to
the first whitespace (sequence case preserved); GenBank id = VERSION accession if present, else the
LOCUS name (sequence upper-cased).
"""
from __future__ import annotations

import numpy as np

_NL, _CR, _SP, _GT = 10, 13, 32, 62


def read_fasta(path: str):
    data = np.fromfile(path, dtype=np.uint8)
    if data.size == 0:
        return [], []
    gt = np.flatnonzero(data == _GT)
    if gt.size:
        at_line_start = np.ones(gt.size, dtype=bool)
        nz = gt > 0
        at_line_start[nz] = data[gt[nz] - 1] == _NL
        gt = gt[at_line_start]
    nl = np.flatnonzero(data == _NL)
    ids, contigs = [], []
    for i, st in enumerate(gt):
        end = int(gt[i + 1]) if i + 1 < gt.size else data.size
        j = np.searchsorted(nl, st)
        hdr_end = int(nl[j]) if j < nl.size and nl[j] < end else end
        header = data[st + 1:hdr_end].tobytes().decode("latin-1").strip("\r")
        parts = header.split(None, 1)
        ids.append(parts[0] if parts else "")
        body = data[min(hdr_end + 1, end):end]
        keep = (body != _NL) & (body != _CR) & (body != _SP)
        # line.rstrip(): other white space goes only where nothing but white space follows on its line
        tws = (body == 9) | (body == 11) | (body == 12) | ((body >= 0x1c) & (body <= 0x1f))
        if tws.any():
            solid = np.flatnonzero(keep & ~tws)                      # bytes that are neither white space nor newlines
            nls = np.flatnonzero(body == _NL)
            at = np.flatnonzero(tws)
            nxt_solid = np.append(solid, body.size + 1)[np.searchsorted(solid, at)]
            nxt_nl = np.append(nls, body.size)[np.searchsorted(nls, at)]
            keep[at[nxt_nl < nxt_solid]] = False
        contigs.append(np.ascontiguousarray(body[keep]))
    return ids, contigs


def read_genbank(path: str):
    ids, contigs = [], []
    locus = version = None
    chunks = None
    with open(path, "r") as fh:
        for line in fh:
            if line.startswith("LOCUS"):
                f = line.split()
                locus, version, chunks = (f[1] if len(f) > 1 else "<unknown id>"), None, None
            elif line.startswith("VERSION"):
                f = line.split()
                version = f[1] if len(f) > 1 else None
            elif line.startswith("ORIGIN"):
                chunks = []
            elif line.startswith("//"):
                if locus is not None:
                    seq = "".join(chunks or []).upper()
                    ids.append(version or locus)
                    contigs.append(np.frombuffer(seq.encode("latin-1"), dtype=np.uint8).copy())
                locus = version = chunks = None
            elif chunks is not None:
                chunks.append("".join(ch for ch in line if ch.isalpha()))
    return ids, contigs


def read_genome(path: str, fmt: str):
    fmt = fmt.lower()
    if fmt == "fasta":
        return read_fasta(path)
    if fmt in ("genbank", "gb"):
        return read_genbank(path)
    raise ValueError(f"unsupported ingroup format {fmt!r}")


def fasta_ids(data: np.ndarray, header_off) -> list:
    """record ids of a FASTA file image given the offsets of the '>' of its records (as found by the
    device parser, pfb_contig_table): the title up to the first whitespace, like SeqRecord.id"""
    ids = []
    n = data.size
    for st in header_off:
        st = int(st)
        chunk = data[st + 1:min(n, st + 1 + 4096)]
        nl = np.flatnonzero(chunk == _NL)
        if nl.size == 0 and st + 1 + 4096 < n:           # a very long title line
            chunk = data[st + 1:]
            nl = np.flatnonzero(chunk == _NL)
        end = int(nl[0]) if nl.size else chunk.size
        parts = chunk[:end].tobytes().decode("latin-1").strip("\r").split(None, 1)
        ids.append(parts[0] if parts else "")
    return ids


def genbank_ids(data: np.ndarray, header_off) -> list:
    """record ids of a GenBank file image given the offsets of its LOCUS lines (device parser): the VERSION
    accession if the header has one, else the LOCUS name -- like SeqRecord.id"""
    ids = []
    n = data.size
    for st in header_off:
        st = int(st)
        locus = version = None
        pos = st
        while pos < n:
            chunk = data[pos:min(n, pos + 65536)]
            nl = np.flatnonzero(chunk == _NL)
            end = pos + (int(nl[0]) if nl.size else chunk.size)
            line = data[pos:end].tobytes().decode("latin-1")
            if line.startswith("LOCUS") and pos == st:
                f = line.split()
                locus = f[1] if len(f) > 1 else "<unknown id>"
            elif line.startswith("VERSION"):
                f = line.split()
                version = f[1] if len(f) > 1 else None
                break
            elif line.startswith("ORIGIN") or line.startswith("//") or (line.startswith("LOCUS") and pos != st) \
                    or line.startswith("FEATURES"):
                break
            pos = end + 1
        ids.append(version or locus or "<unknown id>")
    return ids


def file_ids(data: np.ndarray, header_off, fmt: str) -> list:
    return fasta_ids(data, header_off) if fmt.lower() == "fasta" else genbank_ids(data, header_off)
