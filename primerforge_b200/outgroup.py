"""Drop-in for primerForge's outgroup screen.

    from primerforge_b200.outgroup import _removeOutgroupPrimers
    import bin.main; bin.main._removeOutgroupPrimers = _removeOutgroupPrimers

Same entry point, arguments, in-place effect on ``pairs``, messages and error as
/root/reference/bin/removeOutgroupPrimers.py:181-313.  The scan of every outgroup contig on both strands
(:34-62, 215-226), the product sizes of every pair (:65-122) and the decision which pairs to delete (:281-287)
run in libpfb200.so on the GPU (pfb_og_*, include/pfb200.h); the host builds the ``Product`` objects the
interface asks for (:125-178).  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import time
from dataclasses import dataclass

import numpy as np

from . import _lib
from .engine import PfbError

ERR_ALL_IN_OUTGROUP = "failed to find primer pairs that are absent in the outgroup"     # removeOutgroupPrimers.py:202
NULL_PRODUCT = ("NA", 0)                                                                 # :10
NOT_REMOVED = 0xFFFFFFFF

_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _ch in enumerate(b"ATCG"):
    _CODE[_ch] = _i


def strings_to_words(seqs) -> tuple:
    """primer strings -> (uint64 words, uint8 lengths): A0 T1 C2 G3, first base most significant"""
    n = len(seqs)
    word = np.zeros(n, dtype=np.uint64)
    k = np.zeros(n, dtype=np.uint8)
    for i, s in enumerate(seqs):
        b = np.frombuffer(str(s).encode("ascii"), dtype=np.uint8)
        c = _CODE[b]
        if b.size < 1 or b.size > 32 or (c == 255).any():
            raise ValueError(f"primer {s!r}: only upper-case A/C/G/T strings of 1..32 bases can be keys")
        w = 0
        for v in c.tolist():
            w = (w << 2) | v
        word[i] = w
        k[i] = b.size
    return word, k


@dataclass
class OutgroupResult:
    removed_at: np.ndarray   # [n_pairs] uint32: first outgroup genome with a disallowed product size, 0xFFFFFFFF = kept
    products: np.ndarray     # OG_PRODUCT_DTYPE, distinct, sorted by (pair, genome, contig, size)
    hits: np.ndarray         # OG_HIT_DTYPE, sorted by (genome, contig, strand, key, start)
    timing: dict

    @property
    def kept(self) -> np.ndarray:
        return self.removed_at == NOT_REMOVED


class OutgroupScreen:
    """one pfb_og handle: keys + pairs, outgroup genomes, run, results"""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        rc = self._lib.pfb_og_create(C.byref(self._h), int(device))
        if rc != _lib.PFB_OK:
            raise PfbError(rc, (self._lib.pfb_og_last_error(None) or b"").decode())
        self._keep = []
        self.n_pairs = 0
        self.n_keys = 0

    def close(self):
        if self._h:
            self._lib.pfb_og_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != _lib.PFB_OK:
            raise PfbError(rc, (self._lib.pfb_og_last_error(self._h) or b"").decode())

    def set_pairs(self, fwd, rev, disallowed_lo: int, disallowed_hi: int):
        """fwd / rev: the str(fwd), str(rev) of every pair in dict order (:248, 254); the key set is the distinct
        strings among them (:27)"""
        fwd = [str(s) for s in fwd]
        rev = [str(s) for s in rev]
        index = {}
        for s in fwd:
            index.setdefault(s, len(index))
        for s in rev:
            index.setdefault(s, len(index))
        keys = list(index)
        word, k = strings_to_words(keys)
        self.set_keys_and_pairs(word, k, np.fromiter((index[s] for s in fwd), dtype=np.uint32, count=len(fwd)),
                                np.fromiter((index[s] for s in rev), dtype=np.uint32, count=len(rev)),
                                disallowed_lo, disallowed_hi)
        return keys

    def set_keys_and_pairs(self, word, k, fwd_key, rev_key, disallowed_lo: int, disallowed_hi: int):
        word = np.ascontiguousarray(word, dtype=np.uint64)
        k = np.ascontiguousarray(k, dtype=np.uint8)
        fwd_key = np.ascontiguousarray(fwd_key, dtype=np.uint32)
        rev_key = np.ascontiguousarray(rev_key, dtype=np.uint32)
        if word.size != k.size or fwd_key.size != rev_key.size:
            raise ValueError("keys / pairs: arrays of different lengths")
        self._ck(self._lib.pfb_og_set_keys(self._h, word.ctypes.data, k.ctypes.data, word.size))
        self._ck(self._lib.pfb_og_set_pairs(self._h, fwd_key.ctypes.data, rev_key.ctypes.data, fwd_key.size,
                                            int(disallowed_lo), int(disallowed_hi)))
        self.n_keys, self.n_pairs = int(word.size), int(fwd_key.size)

    def add_genome(self, contigs):
        """contigs: list of uint8 ASCII arrays (str(contig.seq) of every record, in file order)"""
        arrs = [np.ascontiguousarray(c, dtype=np.uint8) for c in contigs]
        cat = np.ascontiguousarray(np.concatenate(arrs)) if arrs else np.zeros(0, dtype=np.uint8)
        off = np.zeros(len(arrs) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([a.size for a in arrs])
        self._keep += [cat, off]           # borrowed by the library until the run
        self._ck(self._lib.pfb_og_add_genome(self._h, cat.ctypes.data, cat.size, off.ctypes.data, len(arrs)))

    def set_scan_mode(self, mode: int):
        """0 auto, 1 the window filter stays in global memory (results identical)"""
        self._ck(self._lib.pfb_og_set_scan_mode(self._h, int(mode)))

    def clear_genomes(self):
        self._ck(self._lib.pfb_og_clear_genomes(self._h))
        self._keep = []

    def run(self):
        self._ck(self._lib.pfb_og_run(self._h))
        self._keep = []

    def compute(self):
        self._ck(self._lib.pfb_og_compute(self._h))

    def timing(self) -> dict:
        t = _lib.PfbOgTiming()
        self._ck(self._lib.pfb_og_timings(self._h, C.byref(t)))
        return {name: getattr(t, name) for name, _ in t._fields_}

    def result(self, with_hits: bool = True) -> OutgroupResult:
        nh, npr, nr = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._ck(self._lib.pfb_og_counts(self._h, C.byref(nh), C.byref(npr), C.byref(nr)))
        removed = np.empty(self.n_pairs, dtype=np.uint32)
        self._ck(self._lib.pfb_og_removed_at(self._h, removed.ctypes.data, removed.size))
        prod = np.empty(npr.value, dtype=_lib.OG_PRODUCT_DTYPE)
        self._ck(self._lib.pfb_og_products(self._h, prod.ctypes.data, prod.size))
        prod = np.unique(prod)             # the reference keeps SETS of (contig, size) (:289-291); structured sort = field order
        prod = prod[removed[prod["pair"]] == NOT_REMOVED]      # a deleted pair has no products any more (:284)
        hits = np.empty(nh.value if with_hits else 0, dtype=_lib.OG_HIT_DTYPE)
        if with_hits:
            self._ck(self._lib.pfb_og_hits(self._h, hits.ctypes.data, hits.size))
            hits = hits[np.lexsort((hits["start"], hits["key"], hits["strand"], hits["contig"], hits["genome"]))]
        return OutgroupResult(removed_at=removed, products=prod, hits=hits, timing=self.timing())


class _Clock:
    """the two calls of bin/Clock.py the stage makes, without its spinner process"""

    def __init__(self):
        self._t0 = time.time()
        self._dt = 0.0

    def printStart(self, msg):
        self._t0 = time.time()
        print(msg, end=" ... ", flush=True)

    def printDone(self):
        self._dt = time.time() - self._t0
        print(f"done {self._dt:.2f}s")

    def getTimeString(self):
        return f"{self._dt:.2f}s"


def _product_class():
    try:
        from bin.Product import Product  # type: ignore  (the reference's own class when it is importable)
        return Product
    except Exception:
        from .primer import Product
        return Product


def _removeOutgroupPrimers(outgroup, pairs, params, device: int = 0) -> None:
    """drop-in for bin/removeOutgroupPrimers.py:181 (same arguments; ``pairs`` is modified in place)"""
    gap = " " * 4
    msg_1 = "removing primer pairs present in the outgroup sequences"
    msg_2 = f"{gap}getting outgroup PCR products"
    clock = _Clock()
    log = params.log
    log.rename(_removeOutgroupPrimers.__name__)
    log.info(msg_1)
    print(msg_1)
    clock.printStart(msg_2)
    log.info(msg_2)

    pair_list = list(pairs.keys())
    names = list(outgroup.keys())
    contig_ids = []
    bad = params.disallowedLens
    with OutgroupScreen(device=device) as scr:
        scr.set_pairs([str(f) for f, _ in pair_list], [str(r) for _, r in pair_list], min(bad), max(bad))
        for name in names:
            ids, seqs = [], []
            for rec in outgroup[name]:
                ids.append(rec.id)
                seqs.append(np.frombuffer(str(rec.seq).encode("latin-1"), dtype=np.uint8))
            contig_ids.append(ids)
            scr.add_genome(seqs)
        scr.run()
        res = scr.result(with_hits=False)
    clock.printDone()
    log.info(f"done {clock.getTimeString()}")
    apply_outgroup_result(res, pairs, pair_list, names, contig_ids, params, clock)


def apply_outgroup_result(res: OutgroupResult, pairs, pair_list, names, contig_ids, params, clock=None) -> None:
    """the host half of the drop-in (no GPU involved: the CPU tests run it on the oracle's result): the deletions of
    :281-287 with the debug lines of :240-247, 294, the error of :299-303, and __processOutgroupResults (:125-178)"""
    Product = _product_class()
    clock = clock or _Clock()
    log = params.log
    gap = " " * 4
    msg_3 = f"{gap}filtering primer pairs"
    msg_4 = f"{gap}processing outgroup results"
    clock.printStart(msg_3)
    log.info(msg_3)

    # delete the pairs with a disallowed product size (:281-287) and report like :240-247, 294
    remaining = len(pair_list)
    for gi, name in enumerate(names):
        gone = int(np.count_nonzero(res.removed_at == gi))
        remaining -= gone
        log.debug(f"removed {gone} pairs after processing {name} ({remaining} pairs remaining)")
    for pi in np.flatnonzero(~res.kept).tolist():
        del pairs[pair_list[pi]]
    clock.printDone()
    log.info(f"done {clock.getTimeString()}")
    if not pairs:
        log.error(ERR_ALL_IN_OUTGROUP)
        raise RuntimeError(ERR_ALL_IN_OUTGROUP)

    log.info(msg_4)
    clock.printStart(msg_4)
    # __processOutgroupResults (:125-178): one Product per kept pair and outgroup genome
    ingroup_name = next(n for v in pairs.values() for n in v.keys() if n not in names)
    prod = res.products
    by_pair = {}
    if prod.size:
        cut = np.flatnonzero(np.r_[True, (prod["pair"][1:] != prod["pair"][:-1]) | (prod["genome"][1:] != prod["genome"][:-1])])
        end = np.r_[cut[1:], prod.size]
        for a, b in zip(cut.tolist(), end.tolist()):
            by_pair[(int(prod["pair"][a]), int(prod["genome"][a]))] = (a, b)
    for pi in np.flatnonzero(res.kept).tolist():
        pair = pair_list[pi]
        dimer = pairs[pair][ingroup_name].dimerTm
        for gi, name in enumerate(names):
            span = by_pair.get((pi, gi))
            if span is None:
                contig, size = NULL_PRODUCT
            elif span[1] - span[0] == 1:
                contig, size = contig_ids[gi][int(prod["contig"][span[0]])], int(prod["size"][span[0]])
            else:   # several products: comma-separated strings (:163-174; the reference's order is set order, here sorted)
                rows = prod[span[0]:span[1]]
                contig = ",".join(contig_ids[gi][int(c)] for c in rows["contig"])
                size = ",".join(str(int(s)) for s in rows["size"])
            pairs[pair][name] = Product(contig, size, -1, -1, dimer, -1, -1, "", -1, -1, "")
    clock.printDone()
    log.info(f"done {clock.getTimeString()}")
