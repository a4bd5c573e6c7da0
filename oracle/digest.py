"""Digests of a stage result (TEST INFRASTRUCTURE ONLY): a full-size result of BASELINE config 1 is megabytes, its
sha256 is a line.  The same canonical byte layout is produced from the reference's Python objects
(tools/time_reference_c1.py, build container only), from the CPU oracle and from the GPU result, so that equality of
digests is equality of: the shared k-mers in dict insertion order, their (contig, leftmost start, strand) in every
genome, and the candidate set of every genome with Primer.start / Primer.end / Tm.
"""
from __future__ import annotations

import hashlib

import numpy as np

_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(b"ATCG"):          # khmer's 2-bit code: A0 T1 C2 G3
    _CODE[_c] = _i


def strings_to_words(kmers) -> tuple:
    """list of ACGT strings -> (uint64 words, uint8 lengths), first base most significant"""
    n = len(kmers)
    ks = np.fromiter((len(s) for s in kmers), dtype=np.uint8, count=n)
    words = np.zeros(n, dtype=np.uint64)
    if n == 0:
        return words, ks
    kmax = int(ks.max())
    buf = np.zeros((n, kmax), dtype=np.uint8)
    raw = np.frombuffer("".join(s.ljust(kmax, "A") for s in kmers).encode("ascii"), dtype=np.uint8).reshape(n, kmax)
    buf[:] = _CODE[raw]
    for j in range(kmax):
        live = j < ks
        words[live] = (words[live] << np.uint64(2)) | buf[live, j].astype(np.uint64)
    return words, ks


def shared_digest(word, k, contig, start, strand) -> str:
    """word [N] u64, k [N] u8 in dict order; contig / start / strand [G, N]"""
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(word, dtype="<u8").tobytes())
    h.update(np.ascontiguousarray(k, dtype="u1").tobytes())
    for g in range(len(contig)):
        h.update(np.ascontiguousarray(contig[g], dtype="<i4").tobytes())
        h.update(np.ascontiguousarray(start[g], dtype="<i8").tobytes())
        h.update(np.ascontiguousarray(strand[g], dtype="u1").tobytes())
    return h.hexdigest()


def candidate_digest(word, k, contig, pstart, pend, strand, tm=None) -> str:
    """one genome's candidates as a SET: rows (contig, min(start, end), k, strand, word) sorted, so the reference's
    unspecified tie order does not matter; Primer.start / Primer.end (flipped on '-', Primer.py:42-45) and,
    if given, the Tm doubles go into the digest in that order"""
    word = np.asarray(word, dtype=np.uint64); k = np.asarray(k, dtype=np.int64)
    contig = np.asarray(contig, dtype=np.int64); pstart = np.asarray(pstart, dtype=np.int64)
    pend = np.asarray(pend, dtype=np.int64); strand = np.asarray(strand, dtype=np.int64)
    order = np.lexsort((word, strand, k, np.minimum(pstart, pend), contig))
    h = hashlib.sha256()
    for a, dt in ((word, "<u8"), (k, "u1"), (contig, "<i4"), (pstart, "<i8"), (pend, "<i8"), (strand, "u1")):
        h.update(np.ascontiguousarray(a[order], dtype=dt).tobytes())
    if tm is not None:
        h.update(np.ascontiguousarray(np.asarray(tm, dtype=np.float64)[order], dtype="<f8").tobytes())
    return h.hexdigest()


def result_digests(word, k, flags, tm, contig, start, strand) -> dict:
    """digests of a columnar result (oracle or GPU, full mode): `shared` + per-genome `candidates`"""
    G = len(contig)
    picked = (np.asarray(flags) & 8) != 0
    out = {"n_shared": int(len(word)), "n_candidates": int(picked.sum()),
           "shared": shared_digest(word, k, contig, start, strand), "candidates": []}
    kk = np.asarray(k, dtype=np.int64)[picked]
    for g in range(G):
        left = np.asarray(start[g], dtype=np.int64)[picked]
        minus = np.asarray(strand[g])[picked] != 0
        right = left + kk - 1
        out["candidates"].append(candidate_digest(np.asarray(word)[picked], kk, np.asarray(contig[g])[picked],
                                                  np.where(minus, right, left), np.where(minus, left, right),
                                                  minus.astype(np.int64), np.asarray(tm)[picked]))
    return out
