"""Deterministic synthetic ingroup genomes.

Ancestor: iid uniform ACGT of length L.  Variant sites: Bernoulli(snp_rate) per
position, alt allele uniform over the other three bases; every genome takes the
alt allele at each site with p = 0.5.  Genomes g >= 1 get ``n_inversions``
non-overlapping in-place reverse-complemented segments of ``inv_frac * L``
bases (exercises '-' strand placement).  Multi-contig mode cuts every genome
into ``n_contigs`` pieces at genome-specific random points, reverse-complements
the odd pieces and shuffles the piece order.  Upper-case ACGT only.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTNacgtn~", b"TGCANtgcan~"):
    _COMP[_a] = _b

BASE_SEED = 20260101


@dataclass
class Genome:
    name: str
    contig_ids: list
    contigs: list  # list[np.ndarray[uint8]] ASCII bases
    path: str | None = None

    @property
    def n_bases(self) -> int:
        return int(sum(c.size for c in self.contigs))


def revcomp_ascii(a: np.ndarray) -> np.ndarray:
    return _COMP[a[::-1]]


def make_genomes(n_genomes: int, length: int, seed: int = BASE_SEED, snp_rate: float = 0.01,
                 n_inversions: int = 4, inv_frac: float = 0.05, n_contigs: int = 1,
                 n_frac: float = 0.0, prefix: str = "ingroup", only=None) -> list:
    """returns ``n_genomes`` Genome objects (in memory).  ``only``: an iterable of genome indices -- the others are
    returned as ``None`` (their random draws are still made, so genome g is the same sequence whatever else is
    generated; this is what lets a rank of a sharded run build just its own genomes)"""
    only = None if only is None else set(int(i) for i in only)
    rng = np.random.Generator(np.random.PCG64(seed))
    anc = rng.integers(0, 4, size=length, dtype=np.uint8)
    sites = np.flatnonzero(rng.random(length) < snp_rate)
    alt = ((anc[sites] + rng.integers(1, 4, size=sites.size, dtype=np.uint8)) & 3).astype(np.uint8)
    out = []
    seg = int(inv_frac * length)
    for g in range(n_genomes):
        want = only is None or g in only
        take = rng.random(sites.size) < 0.5
        seq = None
        if want:
            codes = anc.copy()
            codes[sites[take]] = alt[take]
            seq = _ASCII[codes]
        if n_frac > 0.0:
            nmask = rng.random(length) < n_frac
            if want:
                seq = seq.copy()
                seq[nmask] = ord("N")
        if g >= 1 and n_inversions > 0 and seg > 0:
            q = length // n_inversions
            for i in range(n_inversions):
                if q - seg <= 0:
                    break
                st = i * q + int(rng.integers(0, q - seg))
                if want:
                    seq[st:st + seg] = revcomp_ascii(seq[st:st + seg].copy())
        name = f"{prefix}_{g:03d}.fasta"
        if n_contigs <= 1:
            ids = [f"g{g:03d}_c000"]
            contigs = [np.ascontiguousarray(seq)] if want else None
        else:
            cuts = np.sort(rng.choice(np.arange(1, length), size=n_contigs - 1, replace=False))
            order = rng.permutation(n_contigs)
            contigs = None
            if want:
                pieces = np.split(seq, cuts)
                pieces = [revcomp_ascii(p) if (i & 1) else p for i, p in enumerate(pieces)]
                contigs = [np.ascontiguousarray(pieces[i]) for i in order]
            ids = [f"g{g:03d}_c{c:03d}" for c in range(n_contigs)]
        out.append(Genome(name=name, contig_ids=ids, contigs=contigs) if want else None)
    return out


def write_fasta(genomes: list, directory: str, width: int = 80) -> list:
    """writes one FASTA per genome; returns the ordered list of paths"""
    os.makedirs(directory, exist_ok=True)
    paths = []
    for gen in genomes:
        path = os.path.join(directory, gen.name)
        with open(path, "wb") as fh:
            for cid, arr in zip(gen.contig_ids, gen.contigs):
                fh.write(b">" + cid.encode() + b" synthetic\n")
                n = arr.size
                full = (n // width) * width
                if full:
                    body = np.empty((n // width, width + 1), dtype=np.uint8)
                    body[:, :width] = arr[:full].reshape(-1, width)
                    body[:, width] = 10
                    fh.write(body.tobytes())
                if n > full:
                    fh.write(arr[full:].tobytes() + b"\n")
        gen.path = path
        paths.append(path)
    return paths


# the five BASELINE.json configurations (index -> generator arguments)
CONFIGS = {
    1: dict(n_genomes=3, length=1_000_000, n_contigs=1, k_min=16, k_max=20, max_repeats=4),
    2: dict(n_genomes=10, length=5_000_000, n_contigs=1, k_min=16, k_max=20, max_repeats=4),
    3: dict(n_genomes=50, length=5_000_000, n_contigs=1, k_min=16, k_max=20, max_repeats=4),
    4: dict(n_genomes=200, length=5_000_000, n_contigs=1, k_min=16, k_max=20, max_repeats=4),
    5: dict(n_genomes=20, length=10_000_000, n_contigs=50, k_min=12, k_max=30, max_repeats=3),
}


def make_config(index: int, n_genomes: int | None = None, length: int | None = None) -> tuple:
    """genomes + stage parameters for BASELINE.json config ``index`` (1-based)"""
    cfg = dict(CONFIGS[index])
    if n_genomes is not None:
        cfg["n_genomes"] = n_genomes
    if length is not None:
        cfg["length"] = length
    gens = make_genomes(cfg["n_genomes"], cfg["length"], seed=BASE_SEED + index, n_contigs=cfg["n_contigs"])
    return gens, cfg


def make_outgroup(first_genome: Genome, n_genomes: int = 10, rate: float = 0.05, seed: int = BASE_SEED + 1000) -> list:
    """BASELINE config 3's outgroup genomes: the ingroup's first genome with ``rate`` iid substitutions
    each, same contig structure"""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for g in range(n_genomes):
        contigs = []
        for c in first_genome.contigs:
            seq = c.copy()
            sites = np.flatnonzero(rng.random(seq.size) < rate)
            cur = ((seq[sites] >> 1) & 3).astype(np.uint8)          # A0 C1 T2 G3
            new = (cur + rng.integers(1, 4, size=sites.size, dtype=np.uint8)) & 3
            seq[sites] = np.frombuffer(b"ACTG", dtype=np.uint8)[new]
            contigs.append(seq)
        out.append(Genome(name=f"outgroup_{g:03d}.fasta", contig_ids=[f"o{g:03d}_c{i:03d}" for i in range(len(contigs))],
                          contigs=contigs))
    return out


def make_pairs(first_genome: Genome, n_pairs: int, k_min: int = 16, k_max: int = 20, min_pcr: int = 120, max_pcr: int = 2400,
               seed: int = BASE_SEED + 2000) -> tuple:
    """synthetic primer pairs on the first ingroup genome: the forward primer is a window of its (+) strand, the reverse
    primer the reverse complement of a window min_pcr..max_pcr downstream -- what bin/getPrimerPairs.py hands to
    the outgroup screen (str(fwd), str(rev)).  Returns (key words uint64, key lengths uint8, pair_fwd, pair_rev):
    distinct keys, pairs as key indices."""
    rng = np.random.Generator(np.random.PCG64(seed))
    seq = first_genome.contigs[0]
    n = seq.size
    code = np.zeros(256, dtype=np.uint64)
    for i, ch in enumerate(b"ATCG"):
        code[ch] = i
    c = code[seq]
    kf = rng.integers(k_min, k_max + 1, size=n_pairs)
    kr = rng.integers(k_min, k_max + 1, size=n_pairs)
    size = rng.integers(min_pcr, max_pcr + 1, size=n_pairs)
    a = rng.integers(0, n - max_pcr - 1, size=n_pairs)
    b = a + size - kr                                    # the reverse primer's window starts here, ends at a + size - 1
    words = np.zeros(2 * n_pairs, dtype=np.uint64)
    ks = np.concatenate([kf, kr]).astype(np.uint8)
    for j in range(int(k_max)):
        live = j < kf
        words[:n_pairs][live] = (words[:n_pairs][live] << np.uint64(2)) | c[a[live] + j]
        live = j < kr                                    # reverse complement: last base first, complemented (code ^ 1)
        words[n_pairs:][live] = (words[n_pairs:][live] << np.uint64(2)) | (c[b[live] + kr[live] - 1 - j] ^ np.uint64(1))
    both = np.stack([words, ks.astype(np.uint64)], axis=1)
    uniq, inverse = np.unique(both, axis=0, return_inverse=True)
    inverse = inverse.reshape(-1)
    return (np.ascontiguousarray(uniq[:, 0]), uniq[:, 1].astype(np.uint8), inverse[:n_pairs].astype(np.uint32),
            inverse[n_pairs:].astype(np.uint32))
