"""ctypes wrapper of oracle/pf_oracle.c (the CPU restatement; TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "pf_oracle.c")
_SRCS = [_SRC, os.path.join(_HERE, "pf_oracle_outgroup.c"), os.path.join(_HERE, "pf_oracle_pairs.c")]
_LIB = os.path.join(_HERE, "libpf_oracle.so")


class PfoParams(C.Structure):
    _fields_ = [("k_min", C.c_int32), ("k_max", C.c_int32), ("table_size", C.c_uint64),
                ("min_gc", C.c_double), ("max_gc", C.c_double), ("min_tm", C.c_double), ("max_tm", C.c_double),
                ("max_repeat", C.c_int32), ("thal_mode", C.c_int32), ("temp_tolerance", C.c_double),
                ("mv", C.c_double), ("dv", C.c_double), ("dntp", C.c_double), ("dna", C.c_double),
                ("dmso", C.c_double), ("dmso_fact", C.c_double), ("formamide", C.c_double),
                ("n_threads", C.c_int32), ("_pad", C.c_int32)]


class PfoResult(C.Structure):
    _fields_ = [("n_shared", C.c_uint64), ("n_genomes", C.c_int32),
                ("k", C.POINTER(C.c_uint8)), ("word", C.POINTER(C.c_uint64)), ("flags", C.POINTER(C.c_uint8)),
                ("gc_count", C.POINTER(C.c_int32)), ("tm", C.POINTER(C.c_double)),
                ("contig", C.POINTER(C.c_int32)), ("start", C.POINTER(C.c_int64)), ("strand", C.POINTER(C.c_uint8)),
                ("n_allowed", C.POINTER(C.c_uint64)), ("err", C.c_char * 256)]


class PfoOgResult(C.Structure):
    _fields_ = [("n_hits", C.c_uint64), ("n_products", C.c_uint64), ("hits", C.POINTER(C.c_uint32)),
                ("products", C.POINTER(C.c_uint32)), ("removed_at", C.POINTER(C.c_uint32)), ("err", C.c_char * 256)]


class PfoPpParams(C.Structure):
    _fields_ = [("max_bin", C.c_int32), ("min_primer_len", C.c_int32), ("min_pcr", C.c_int32), ("max_pcr", C.c_int32),
                ("max_tm_diff", C.c_double)]


class PfoPpResult(C.Structure):
    _fields_ = [("n_pairs", C.c_uint64), ("pairs", C.POINTER(C.c_uint32)), ("shared", C.POINTER(C.c_uint32)),
                ("bin_of", C.POINTER(C.c_uint32)), ("reduced", C.POINTER(C.c_uint8)), ("err", C.c_char * 256)]


def build(force: bool = False) -> str:
    """gcc -O2 -fopenmp -shared; rebuilt when the source is newer"""
    have = [f for f in _SRCS if os.path.exists(f)]      # (a box without the sources uses the prebuilt library)
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(f) > os.path.getmtime(_LIB) for f in have):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-fno-fast-math", "-ffp-contract=off", "-shared", "-fPIC",
                               "-o", _LIB] + _SRCS + ["-lm"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build()
        _lib = C.CDLL(path)
        _lib.pfo_run.restype = C.POINTER(PfoResult)
        _lib.pfo_run.argtypes = [C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int32),
                                 C.POINTER(PfoParams)]
        _lib.pfo_free.argtypes = [C.POINTER(PfoResult)]
        _lib.pfo_free.restype = None
        _lib.pfo_calc_tm.restype = C.c_double
        _lib.pfo_calc_tm.argtypes = [C.c_char_p, C.POINTER(PfoParams)]
        _lib.pfo_max_threads.restype = C.c_int
        _lib.pfo_outgroup.restype = C.POINTER(PfoOgResult)
        _lib.pfo_outgroup.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                      C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_int32)]
        _lib.pfo_og_free.argtypes = [C.POINTER(PfoOgResult)]
        _lib.pfo_og_free.restype = None
        _lib.pfo_pairs.restype = C.POINTER(PfoPpResult)
        _lib.pfo_pairs.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(PfoPpParams)]
        _lib.pfo_pp_free.argtypes = [C.POINTER(PfoPpResult)]
        _lib.pfo_pp_free.restype = None
    return _lib


def default_params(k_min=16, k_max=20, min_gc=40.0, max_gc=60.0, min_tm=55.0, max_tm=68.0, max_repeat=4,
                   thal_mode=0, temp_tolerance=5.0, table_size=9999991, n_threads=0) -> PfoParams:
    return PfoParams(k_min=k_min, k_max=k_max, table_size=table_size, min_gc=min_gc, max_gc=max_gc, min_tm=min_tm,
                     max_tm=max_tm, max_repeat=max_repeat, thal_mode=thal_mode, temp_tolerance=temp_tolerance,
                     mv=50.0, dv=1.5, dntp=0.6, dna=50.0, dmso=0.0, dmso_fact=0.6, formamide=0.8,
                     n_threads=n_threads)


@dataclass
class OracleResult:
    n_shared: int
    k: np.ndarray          # [N] uint8
    word: np.ndarray       # [N] uint64 forward 2-bit word of the key string
    flags: np.ndarray      # [N] uint8: 1 gc ok, 2 tm ok, 4 no repeat, 8 picked
    gc_count: np.ndarray   # [N] int32
    tm: np.ndarray         # [N] float64
    contig: np.ndarray     # [G, N] int32
    start: np.ndarray      # [G, N] int64 leftmost (+) coordinate
    strand: np.ndarray     # [G, N] uint8 (0 '+', 1 '-')
    n_allowed: np.ndarray  # [G, nk]

    def kmer_strings(self) -> list:
        return [word_to_str(int(w), int(k)) for w, k in zip(self.word, self.k)]


_LET = "ATCG"


def word_to_str(w: int, k: int) -> str:
    return "".join(_LET[(w >> (2 * (k - 1 - i))) & 3] for i in range(k))


def calc_tm(seq: str, params: PfoParams | None = None) -> float:
    p = params or default_params()
    return float(lib().pfo_calc_tm(seq.encode("ascii"), C.byref(p)))


def max_threads() -> int:
    return int(lib().pfo_max_threads())


def run(genomes, params: PfoParams) -> OracleResult:
    """genomes: list of objects with ``.contigs`` (list of uint8 ASCII arrays); genome 0 is the first genome"""
    L = lib()
    G = len(genomes)
    bases_keep, offs_keep = [], []
    for gen in genomes:
        contigs = [np.ascontiguousarray(c, dtype=np.uint8) for c in gen.contigs]
        cat = np.concatenate(contigs) if contigs else np.zeros(0, np.uint8)
        off = np.zeros(len(contigs) + 1, dtype=np.int64)
        off[1:] = np.cumsum([c.size for c in contigs])
        bases_keep.append(np.ascontiguousarray(cat))
        offs_keep.append(off)
    bptr = (C.c_void_p * G)(*[b.ctypes.data for b in bases_keep])
    optr = (C.c_void_p * G)(*[o.ctypes.data for o in offs_keep])
    nct = (C.c_int32 * G)(*[len(g.contigs) for g in genomes])
    rp = L.pfo_run(G, bptr, optr, nct, C.byref(params))
    try:
        r = rp.contents
        err = r.err.decode()
        if err:
            raise RuntimeError(f"oracle: {err}")
        N = int(r.n_shared)
        nk = params.k_max - params.k_min + 1

        def arr(ptr, n, dt):
            if n == 0:
                return np.zeros(0, dtype=dt)
            return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True)

        out = OracleResult(
            n_shared=N, k=arr(r.k, N, np.uint8), word=arr(r.word, N, np.uint64), flags=arr(r.flags, N, np.uint8),
            gc_count=arr(r.gc_count, N, np.int32), tm=arr(r.tm, N, np.float64),
            contig=arr(r.contig, N * G, np.int32).reshape(G, N), start=arr(r.start, N * G, np.int64).reshape(G, N),
            strand=arr(r.strand, N * G, np.uint8).reshape(G, N), n_allowed=arr(r.n_allowed, G * nk, np.uint64).reshape(G, nk))
    finally:
        L.pfo_free(rp)
    return out


@dataclass
class OutgroupOracleResult:
    hits: np.ndarray        # [n, 5] uint32: key, genome, contig, start, strand -- __extractKmerData's lists in scan order
    products: np.ndarray    # [n, 4] uint32: pair, genome, contig, size -- distinct, sorted, kept pairs only
    removed_at: np.ndarray  # [n_pairs] uint32, 0xFFFFFFFF = kept


def run_outgroup(keys, pair_fwd, pair_rev, bad_lo: int, bad_hi: int, genomes) -> OutgroupOracleResult:
    """oracle/pf_oracle_outgroup.c: keys = distinct primer strings, pair_fwd / pair_rev = key indices of str(fwd) /
    str(rev) per pair (dict order), genomes = objects with ``.contigs`` (outgroup genomes in order)"""
    L = lib()
    kb = [str(k).encode("ascii") for k in keys]
    key_bytes = np.frombuffer(b"".join(kb) + b"\0", dtype=np.uint8).copy()
    key_len = np.array([len(k) for k in kb], dtype=np.uint8)
    key_off = np.zeros(len(kb) + 1, dtype=np.uint64)
    key_off[1:] = np.cumsum([len(k) for k in kb])
    pf = np.ascontiguousarray(pair_fwd, dtype=np.uint32)
    pr = np.ascontiguousarray(pair_rev, dtype=np.uint32)
    G = len(genomes)
    bases_keep, offs_keep = [], []
    for gen in genomes:
        contigs = [np.ascontiguousarray(c, dtype=np.uint8) for c in gen.contigs]
        cat = np.concatenate(contigs) if contigs else np.zeros(0, np.uint8)
        off = np.zeros(len(contigs) + 1, dtype=np.int64)
        off[1:] = np.cumsum([c.size for c in contigs])
        bases_keep.append(np.ascontiguousarray(np.r_[cat, np.zeros(1, np.uint8)]))
        offs_keep.append(off)
    bptr = (C.c_void_p * max(G, 1))(*[b.ctypes.data for b in bases_keep])
    optr = (C.c_void_p * max(G, 1))(*[o.ctypes.data for o in offs_keep])
    nct = (C.c_int32 * max(G, 1))(*[len(g.contigs) for g in genomes])
    rp = L.pfo_outgroup(len(kb), key_bytes.ctypes.data, key_off.ctypes.data, key_len.ctypes.data, pf.size, pf.ctypes.data,
                        pr.ctypes.data, int(bad_lo), int(bad_hi), G, bptr, optr, nct)
    try:
        r = rp.contents
        if r.err:
            raise RuntimeError(f"oracle: {r.err.decode()}")

        def arr(ptr, n, w):
            if n == 0:
                return np.zeros((0, w) if w > 1 else 0, dtype=np.uint32)
            a = np.ctypeslib.as_array(ptr, shape=(n * w,)).astype(np.uint32, copy=True)
            return a.reshape(n, w) if w > 1 else a

        return OutgroupOracleResult(hits=arr(r.hits, int(r.n_hits), 5), products=arr(r.products, int(r.n_products), 4),
                                    removed_at=arr(r.removed_at, pf.size, 1))
    finally:
        L.pfo_og_free(rp)


@dataclass
class PairOracleResult:
    pairs: np.ndarray       # [n, 6] uint32: fwd candidate, rev candidate, contig, fbin, rbin, pcrLen -- the generator's order
    shared: np.ndarray      # [n, G, 4] uint32: valid, length, fbin, rbin per genome (genome 0 = the first genome)
    bin_of: np.ndarray      # [G, S] uint32
    reduced: np.ndarray     # [S] uint8

    @property
    def kept(self) -> np.ndarray:
        """pairs with a product in every genome (getPrimerPairs.py:459-462)"""
        return self.shared[:, :, 0].all(axis=1) if self.pairs.shape[0] else np.zeros(0, dtype=bool)


def run_pairs(word, k, tm, genomes, max_bin=64, min_primer_len=16, min_pcr=120, max_pcr=2400, max_tm_diff=5.0) -> PairOracleResult:
    """oracle/pf_oracle_pairs.c.  word / k / tm: the S candidates; genomes: per genome a tuple (id, contig, left, strand)
    of arrays of length S in the genome's list order (first genome first)"""
    L = lib()
    word = np.ascontiguousarray(word, dtype=np.uint64)
    k = np.ascontiguousarray(k, dtype=np.uint8)
    tm = np.ascontiguousarray(tm, dtype=np.float64)
    S, G = word.size, len(genomes)
    keep = [[np.ascontiguousarray(g[0], dtype=np.uint32), np.ascontiguousarray(g[1], dtype=np.uint32),
             np.ascontiguousarray(g[2], dtype=np.uint32), np.ascontiguousarray(g[3], dtype=np.uint8)] for g in genomes]
    ptrs = [(C.c_void_p * G)(*[kk[j].ctypes.data for kk in keep]) for j in range(4)]
    prm = PfoPpParams(max_bin=max_bin, min_primer_len=min_primer_len, min_pcr=min_pcr, max_pcr=max_pcr, max_tm_diff=max_tm_diff)
    rp = L.pfo_pairs(S, word.ctypes.data, k.ctypes.data, tm.ctypes.data, G, ptrs[0], ptrs[1], ptrs[2], ptrs[3], C.byref(prm))
    try:
        r = rp.contents
        if r.err:
            raise RuntimeError(f"oracle: {r.err.decode()}")
        n = int(r.n_pairs)
        pairs = np.ctypeslib.as_array(r.pairs, shape=(max(n, 1) * 6,))[:n * 6].astype(np.uint32).reshape(n, 6)
        shared = np.ctypeslib.as_array(r.shared, shape=(max(n * G, 1) * 4,))[:n * G * 4].astype(np.uint32).reshape(n, G, 4)
        bin_of = np.ctypeslib.as_array(r.bin_of, shape=(max(G * S, 1),))[:G * S].astype(np.uint32).reshape(G, S)
        reduced = np.ctypeslib.as_array(r.reduced, shape=(max(S, 1),))[:S].astype(np.uint8)
        return PairOracleResult(pairs=pairs, shared=shared, bin_of=bin_of, reduced=reduced)
    finally:
        L.pfo_pp_free(rp)
