_TR = str.maketrans("ACGTacgtNnRYKMSWBDHVrykmswbdhv", "TGCAtgcaNnYRMKSWVHDByrmkswvhdb")


class Seq(str):
    """str subclass: hashes / compares equal to the plain string (the reference uses
    ``Primer.seq`` as a key into a dict keyed by str, ``getCandidateKmers.py:459``)."""
    __slots__ = ()

    def __new__(cls, data=""):
        return super().__new__(cls, str(data))

    def reverse_complement(self):
        return Seq(str.__str__(self)[::-1].translate(_TR))

    def complement(self):
        return Seq(str.__str__(self).translate(_TR))

    def upper(self):
        return Seq(str.upper(self))

    def lower(self):
        return Seq(str.lower(self))

    def __getitem__(self, idx):
        out = str.__getitem__(self, idx)
        return Seq(out) if isinstance(idx, slice) else out

    def __str__(self):
        return str.__str__(self)

    def __repr__(self):
        return f"Seq({str.__str__(self)!r})"

    def __reduce__(self):
        return (Seq, (str.__str__(self),))
