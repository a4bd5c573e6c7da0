"""Primer record of the candidate-k-mer stage.

When primerForge itself is importable (``bin.Primer``) the stage returns real ``bin.Primer.Primer``
instances, so everything downstream (``getPrimerPairs`` ... pickles) keeps working unchanged.
Otherwise it returns the structurally identical record defined here (same attribute names, the
``-`` strand start/end flip of /root/reference/bin/Primer.py:42-45, equality and hashing by
sequence, ``reverseComplement`` swapping the secondary-structure fields).

Records coming out of the GPU stage are created without running the constructor: GC % and Tm are
the values computed on the device (bit-identical to the formula below), which is what makes
G x |S| records affordable.
"""
from __future__ import annotations

import math

try:  # biopython present (a real primerForge install)
    from Bio.Seq import Seq  # type: ignore
except Exception:  # pragma: no cover - exercised in this image
    _RC = str.maketrans("ACGTacgtNn", "TGCAtgcaNn")

    class Seq(str):  # minimal stand-in: hashes / compares like the plain string
        __slots__ = ()

        def reverse_complement(self):
            return Seq(str.__str__(self)[::-1].translate(_RC))

        def upper(self):
            return Seq(str.upper(self))

        def __str__(self):
            return str.__str__(self)

        def __reduce__(self):
            return (Seq, (str.__str__(self),))

# SantaLucia (1998) nearest-neighbour values in primer3's oligotm units (100 cal/mol, 0.1 cal/K/mol)
_NN = {"AA": (79, 222), "AC": (84, 224), "AG": (78, 210), "AT": (72, 204), "CA": (85, 227), "CC": (80, 199),
       "CG": (106, 272), "CT": (78, 210), "GA": (82, 222), "GC": (98, 244), "GG": (80, 199), "GT": (84, 224),
       "TA": (72, 213), "TC": (82, 222), "TG": (85, 227), "TT": (79, 222)}
_PAIR = {"A": "T", "T": "A", "C": "G", "G": "C"}


def oligo_tm(seq: str, mv=50.0, dv=1.5, dntp=0.6, dna=50.0, dmso=0.0, dmso_fact=0.6, formamide=0.8) -> float:
    """primer3 ``calc_tm`` with its defaults (what bin/Primer.py:104 calls); used only when a Primer
    is built by hand and primer3 is not installed -- the stage itself takes Tm from the GPU"""
    s = str(seq).upper()
    n = len(s)
    if n < 2 or any(c not in _PAIR for c in s):
        return -999999.9999
    sym = n % 2 == 0 and all(_PAIR[s[i]] == s[n - 1 - i] for i in range(n // 2))
    dh, ds = 0, (14 if sym else 0)
    for end in (s[0], s[-1]):
        if end in "AT":
            dh, ds = dh - 23, ds - 41
        else:
            dh, ds = dh - 1, ds + 28
    for i in range(n - 1):
        h, e = _NN[s[i:i + 2]]
        dh += h
        ds += e
    if dv == 0:
        dntp = 0
    dv = max(dv, dntp)
    salt = mv + 120.0 * math.sqrt(dv - dntp)
    delta_s = ds * -0.1 + 0.368 * (n - 1) * math.log(salt / 1000.0)
    tm = dh * -100.0 / (delta_s + 1.987 * math.log(dna / (1000000000.0 if sym else 4000000000.0))) - 273.15
    tm -= dmso * dmso_fact
    gc = sum(c in "GC" for c in s)
    return tm + (0.453 * gc / n - 2.88) * formamide      # oligotm.c: 0.453 * GC_count / len, left to right


def _calc_tm(seq: str) -> float:
    try:
        import primer3  # type: ignore
        return primer3.calc_tm(seq)
    except ImportError:
        return oligo_tm(seq)


class Primer:
    """mirror of bin.Primer.Primer (used when primerForge is not importable)"""
    PLUS = "+"
    MINUS = "-"

    def __init__(self, seq, contig: str, start: int, length: int, strand: str):
        if strand not in (self.PLUS, self.MINUS):
            raise ValueError("invalid strand specified")
        first, last = start, start + length - 1
        if strand == self.MINUS:
            first, last = last, first
        self.seq = Seq(str(seq).upper())
        self.start, self.end = first, last
        self.contig, self.strand = contig, strand
        self.hairpinTm = self.rcHairpin = self.homodimerTm = self.rcHomodimer = None
        self._Primer__length = length
        text = str(self.seq)
        self.gcPer = sum(b in "GC" for b in text) / length * 100
        self.Tm = _calc_tm(text)

    def __hash__(self):
        return hash(self.seq)

    def __len__(self):
        return self._Primer__length

    def __str__(self):
        return str(self.seq)

    __repr__ = __str__

    def __eq__(self, other):
        if isinstance(other, Primer):
            return self.seq == other.seq
        if isinstance(other, str):
            return self.seq == other
        raise ValueError(f"cannot compare Primer to type '{type(other)}'")

    def __ne__(self, other):
        return not self.seq == other.seq

    def __format__(self, spec):
        return format(str(self.seq), spec)

    def reverseComplement(self):
        anchor = self.start if self.strand == self.PLUS else self.end
        flipped = self.MINUS if self.strand == self.PLUS else self.PLUS
        new = type(self)(self.seq.reverse_complement(), self.contig, anchor, len(self), flipped)
        new.hairpinTm, new.rcHairpin = self.rcHairpin, self.hairpinTm
        new.homodimerTm, new.rcHomodimer = self.rcHomodimer, self.homodimerTm
        return new

    def getMinimizer(self, lmerSize: int, strand: str):
        if len(self.seq) < lmerSize:
            raise ValueError("Window size should be less than or equal to the k-mer length.")
        if strand not in (self.PLUS, self.MINUS):
            raise ValueError("Invalid strand")
        text = self.seq if strand == self.strand else self.seq.reverse_complement()
        return min((text[i:i + lmerSize] for i in range(len(text) - lmerSize + 1)), key=str)


_LocalPrimer = Primer


class Product:
    """mirror of /root/reference/bin/Product.py:4-44 (PCR product of a primer pair in one genome); used only when
    bin.Product is not importable"""

    def __init__(self, contig, size, fbin, rbin, dimerTm, fStart, fEnd, fStrand, rStart, rEnd, rStrand):
        self.contig = contig
        self.size = size
        self.dimerTm = dimerTm
        self.fbin = fbin
        self.rbin = rbin
        self.fStart = fStart
        self.fEnd = fEnd
        self.fStrand = fStrand
        self.rStart = rStart
        self.rEnd = rEnd
        self.rStrand = rStrand

    def __str__(self):
        return f"{self.contig}: {self.size} bp"

    def __repr__(self):
        return str(self)


def primer_class():
    """bin.Primer.Primer when primerForge is importable, else the local mirror"""
    try:
        from bin.Primer import Primer as RefPrimer  # type: ignore
        return RefPrimer
    except Exception:
        return _LocalPrimer


def make_primer(cls, text: str, contig: str, leftmost: int, k: int, minus: bool, tm: float, gc_per: float, thal=None):
    """record without running the constructor (no per-object Tm call); same fields as
    /root/reference/bin/Primer.py:29-50"""
    p = cls.__new__(cls)
    p.seq = Seq(text)
    if minus:
        p.start, p.end, p.strand = leftmost + k - 1, leftmost, "-"
    else:
        p.start, p.end, p.strand = leftmost, leftmost + k - 1, "+"
    p.contig = contig
    p.Tm = tm
    p.gcPer = gc_per
    if thal is None:
        p.hairpinTm = p.rcHairpin = p.homodimerTm = p.rcHomodimer = None
    else:
        p.hairpinTm, p.rcHairpin, p.homodimerTm, p.rcHomodimer = thal
    p._Primer__length = k
    return p
