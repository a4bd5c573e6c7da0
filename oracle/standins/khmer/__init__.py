"""Stand-in for the third-party ``khmer`` package (>= 2.1, not installable here).

TEST INFRASTRUCTURE ONLY.  Used to run the reference's *unmodified*
``bin/getCandidateKmers.py`` in this container so golden vectors can be made
(see oracle/make_golden.py).  Only the three methods the reference calls are
provided (``/root/reference/bin/getCandidateKmers.py:46-47,50-51,54-55,38``).

Semantics are *recalled* from khmer 2.1.1 (``khmer/_oxli/graphs.pyx``,
``include/oxli/kmer_hash.hh``) and are NOT verifiable against a real install
in this image ("parity unpinned" for these internals):

* ``Countgraph(k, starting_size, n_tables)`` sizes each table with the
  ``n_tables`` largest primes below ``starting_size`` (1e7 -> 9 999 991).
* two-bit code: A=0, T=1, C=2, everything else 3; complement code: A->1,
  T->0, C->3, everything else 2 (``twobit_repr`` / ``twobit_comp``); no
  upper-casing, no validation in ``consume`` / ``get``.
* hash = min(forward word, reverse-complement word); bin = hash % prime;
  one-byte saturating counter (255).
* ``get_kmers(s)`` returns the raw substrings.
"""
import os

import numpy as np

_FWD = np.full(256, 3, dtype=np.uint64)
_FWD[ord("A")] = 0
_FWD[ord("T")] = 1
_FWD[ord("C")] = 2
_CMP = np.full(256, 2, dtype=np.uint64)
_CMP[ord("A")] = 1
_CMP[ord("T")] = 0
_CMP[ord("C")] = 3

_FWD_TR = bytes(int(_FWD[i]) + 48 for i in range(256))
_CMP_TR = bytes(int(_CMP[i]) + 48 for i in range(256))


def _is_prime(n: int) -> bool:
    if n < 2:
        return False
    if n % 2 == 0:
        return n == 2
    i = 3
    while i * i <= n:
        if n % i == 0:
            return False
        i += 2
    return True


def get_n_primes_near_x(n_primes: int, max_size) -> list:
    """the n largest primes strictly below max_size (khmer/utils semantics)"""
    primes = []
    i = int(max_size) - 1
    if i % 2 == 0:
        i -= 1
    while len(primes) != n_primes and i > 0:
        if _is_prime(i):
            primes.append(i)
        i -= 2
    return primes


def _hash_str(kmer: str) -> int:
    b = kmer.encode("latin-1")
    h = int(b.translate(_FWD_TR), 4)
    r = int(b[::-1].translate(_CMP_TR), 4)
    return h if h < r else r


def _hash_all(seq: str, k: int) -> np.ndarray:
    """canonical hash of every window of seq (vectorised)"""
    b = np.frombuffer(seq.encode("latin-1"), dtype=np.uint8)
    n = b.size - k + 1
    if n <= 0:
        return np.zeros(0, dtype=np.uint64)
    f = _FWD[b]
    c = _CMP[b]
    h = np.zeros(n, dtype=np.uint64)
    r = np.zeros(n, dtype=np.uint64)
    for j in range(k):
        h = (h << np.uint64(2)) | f[j:j + n]
        r = r | (c[j:j + n] << np.uint64(2 * j))
    return np.minimum(h, r)


class Countgraph:
    def __init__(self, k, starting_size, n_tables, primes=None):
        self._k = int(k)
        if self._k > 32:
            raise ValueError("k must be <= 32")
        # PF_STANDIN_TABLE (test hook): force a small table so that sketch collisions and junction
        # pollution actually occur on kilobase-sized fixtures
        forced = os.environ.get("PF_STANDIN_TABLE")
        if forced:
            starting_size = int(forced) + 1
        self._primes = list(primes) if primes else get_n_primes_near_x(int(n_tables), starting_size)
        self._tables = [np.zeros(p, dtype=np.uint8) for p in self._primes]

    def ksize(self):
        return self._k

    def hashsizes(self):
        return list(self._primes)

    def consume(self, sequence: str) -> int:
        if len(sequence) < self._k:
            raise ValueError("sequence length must be >= the k-mer size")
        hv = _hash_all(sequence, self._k)
        for p, t in zip(self._primes, self._tables):
            cnt = np.bincount((hv % np.uint64(p)).astype(np.int64), minlength=p)
            tot = t.astype(np.int64) + cnt
            np.minimum(tot, 255, out=tot)
            t[:] = tot.astype(np.uint8)
        return int(hv.size)

    def get(self, kmer) -> int:
        if isinstance(kmer, int):
            hv = kmer
        else:
            hv = _hash_str(kmer[: self._k])
        return int(min(int(t[hv % p]) for p, t in zip(self._primes, self._tables)))

    def get_kmers(self, sequence: str) -> list:
        k = self._k
        return [sequence[i:i + k] for i in range(len(sequence) - k + 1)]
