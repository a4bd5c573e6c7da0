"""Host-side driver of the CUDA candidate-k-mer engine (thin layer over the C ABI)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib

ERR_NO_SHARED = "failed to identify a set of kmers shared between the ingroup genomes"     # getCandidateKmers.py:538
ERR_NO_CANDIDATES = "none of the ingroup kmers are suitable for use as a primer"            # getCandidateKmers.py:539

_LETTERS = np.frombuffer(b"ATCG", dtype=np.uint8)


class PfbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libpfb200 error {code}: {msg}")
        self.code = code


@dataclass
class StageResult:
    records: np.ndarray      # RECORD_DTYPE [N] in the reference's dict order
    positions: list          # per local genome: POS_DTYPE [N]
    timing: dict

    @property
    def picked(self) -> np.ndarray:
        return (self.records["flags"] & _lib.F_PICK) != 0

    def kmer_strings(self, idx=None) -> list:
        rec = self.records if idx is None else self.records[idx]
        return words_to_strings(rec["word"], rec["k"])


@dataclass
class IdView:
    """the per-id result columns (views of pinned host buffers owned by the engine, valid until its next run).
    ids = genome-1 k-mers unique in genome 1, in the reference's dict order; alive = shared by every ingroup genome;
    pick = PFB_F_PICK.  pos[g, id] = genome-local padded coordinate | strand << 31 (see include/pfb200.h)."""
    n_ids: int
    pitch: int             # entries each column / row occupies in the pinned buffers (>= n_ids: the copies move whole rows)
    word: np.ndarray       # [n_ids] uint64 (None when the records were left on the device)
    tm: np.ndarray         # [n_ids] float64
    info: np.ndarray       # [n_ids] uint16: (k - 1) | gc_count << 5 | flags << 11
    pos: np.ndarray        # [G, n_ids] uint32
    alive: np.ndarray      # [n_ids] uint8
    pick: np.ndarray       # [n_ids] uint8

    @property
    def k(self) -> np.ndarray:
        return (self.info & 31).astype(np.uint8) + 1

    @property
    def gc_count(self) -> np.ndarray:
        return ((self.info >> 5) & 63).astype(np.uint16)

    @property
    def flags(self) -> np.ndarray:
        return (self.info >> 11).astype(np.uint8)


def words_to_strings(words: np.ndarray, ks: np.ndarray) -> list:
    """2-bit words (A0 T1 C2 G3, first base most significant) -> str"""
    words = np.asarray(words, dtype=np.uint64)
    ks = np.asarray(ks, dtype=np.int64)
    n = words.size
    if n == 0:
        return []
    kmax = int(ks.max())
    shifts = (2 * (ks[:, None] - 1 - np.arange(kmax)[None, :])).clip(min=0).astype(np.uint64)
    codes = ((words[:, None] >> shifts) & np.uint64(3)).astype(np.intp)
    chars = _LETTERS[codes]
    raw = chars.tobytes()
    return [raw[i * kmax:i * kmax + int(ks[i])].decode("ascii") for i in range(n)]


class Engine:
    """one CUDA context (= one GPU).  Usage: add_genome()* -> run() -> result()"""

    def __init__(self, device: int = 0, _borrowed_ctx=None, **params):
        self._lib = _lib.load()
        p = make_params(**params)
        self.params = p
        self._owned = _borrowed_ctx is None
        if _borrowed_ctx is not None:          # a context that belongs to a MultiEngine
            self._ctx = C.c_void_p(_borrowed_ctx)
        else:
            self._ctx = C.c_void_p()
            rc = self._lib.pfb_create(C.byref(self._ctx), int(device), C.byref(p))
            if rc != _lib.PFB_OK:
                raise PfbError(rc, self._lib.pfb_last_error(None).decode())
        self._keep = []
        self.n_genomes = 0
        self._multi_contig = False
        self._pinned = {}

    # -- plumbing
    def _check(self, rc: int):
        if rc != _lib.PFB_OK:
            raise PfbError(rc, self._lib.pfb_last_error(self._ctx).decode())

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            if self._owned:
                self._lib.pfb_destroy(self._ctx)
            self._ctx = C.c_void_p()
            for ptr, _cap in self._pinned.values():
                self._lib.pfb_host_free(ptr)
            self._pinned = {}

    def _pinned_array(self, name: str, dtype, shape):
        """grow-only pinned host buffer viewed as a numpy array (full PCIe bandwidth for the D2H copy)"""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr, cap = self._pinned.get(name, (None, 0))
        if cap < nbytes:
            if ptr:
                self._lib.pfb_host_free(ptr)
            want = nbytes + nbytes // 4 + 4096
            new = C.c_void_p()
            rc = self._lib.pfb_host_alloc(C.byref(new), want)
            if rc != _lib.PFB_OK:
                raise PfbError(rc, "cudaHostAlloc failed")
            ptr, cap = new, want
            self._pinned[name] = (ptr, cap)
        buf = (C.c_uint8 * nbytes).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- input
    def add_genome(self, contigs=None, bases: np.ndarray | None = None, offsets: np.ndarray | None = None):
        """contigs: list of uint8 ASCII arrays; or a pre-concatenated (bases, offsets) pair (zero copy)"""
        if bases is None:
            contigs = [np.ascontiguousarray(c, dtype=np.uint8) for c in contigs]
            bases = np.concatenate(contigs) if len(contigs) > 1 else (contigs[0] if contigs else np.zeros(0, np.uint8))
            offsets = np.zeros(len(contigs) + 1, dtype=np.uint64)
            offsets[1:] = np.cumsum([c.size for c in contigs])
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._keep.append((bases, offsets))
        self._check(self._lib.pfb_add_genome(self._ctx, bases.ctypes.data, bases.size, offsets.ctypes.data, offsets.size - 1))
        self.n_genomes += 1
        self._multi_contig |= offsets.size > 2

    def add_genome_ptr(self, ptr: int, n_bases: int, offsets: np.ndarray):
        """bases already in (pinned) host memory at address `ptr`"""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._keep.append((None, offsets))
        self._check(self._lib.pfb_add_genome(self._ctx, ptr, n_bases, offsets.ctypes.data, offsets.size - 1))
        self.n_genomes += 1
        self._multi_contig |= offsets.size > 2

    def add_fasta(self, data: np.ndarray | None = None, ptr: int | None = None, n_bytes: int | None = None, fmt: str = "fasta"):
        """a genome given as its FASTA (or, fmt="genbank", GenBank) file image (uint8 array, or a raw host
        pointer e.g. into pinned memory); the records are found and compacted on the device -- see
        contig_table() after run()"""
        if data is not None:
            data = np.ascontiguousarray(data, dtype=np.uint8)
            ptr, n_bytes = data.ctypes.data, data.size
        self._keep.append((data, None))
        fn = self._lib.pfb_add_fasta if fmt == "fasta" else self._lib.pfb_add_genbank
        self._check(fn(self._ctx, ptr, int(n_bytes)))
        self.n_genomes += 1
        self._multi_contig = None       # known after the device has parsed the files

    def add_genbank(self, data: np.ndarray | None = None, ptr: int | None = None, n_bytes: int | None = None):
        self.add_fasta(data, ptr, n_bytes, fmt="genbank")

    def genome_info(self, genome_idx: int):
        nc, nb = C.c_uint32(), C.c_uint64()
        self._check(self._lib.pfb_genome_info(self._ctx, genome_idx, C.byref(nc), C.byref(nb)))
        return int(nc.value), int(nb.value)

    def contig_table(self, genome_idx: int):
        """(offset of each record's '>' in the FASTA file, length of each record)"""
        nc, _ = self.genome_info(genome_idx)
        hdr = np.zeros(nc, dtype=np.uint64)
        ln = np.zeros(nc, dtype=np.uint64)
        self._check(self._lib.pfb_contig_table(self._ctx, genome_idx, hdr.ctypes.data, ln.ctypes.data, nc))
        return hdr, ln

    def clear(self):
        self._check(self._lib.pfb_clear_genomes(self._ctx))
        self._keep = []
        self.n_genomes = 0
        self._multi_contig = False

    # -- run
    def upload(self):
        self._check(self._lib.pfb_upload(self._ctx))

    def upload_async(self):
        self._check(self._lib.pfb_upload_async(self._ctx))

    def upload_wait(self):
        self._check(self._lib.pfb_upload_wait(self._ctx))

    def compute(self):
        self._check(self._lib.pfb_compute(self._ctx))

    def run(self):
        self._check(self._lib.pfb_run(self._ctx))

    def sync(self):
        self._check(self._lib.pfb_sync(self._ctx))

    def set_scan_mode(self, mode: int):
        """0 auto, 1 plain L2 probe, 2 shared-memory filter (same results)"""
        self._check(self._lib.pfb_set_scan_mode(self._ctx, int(mode)))

    def set_option(self, option: int, value: int):
        self._check(self._lib.pfb_set_option(self._ctx, int(option), int(value)))

    # -- more than one GPU, one process per GPU: rank 0 makes the id, every rank joins; run() / compute() are then
    # collective calls (the exchanges happen inside the library, NCCL on the context's stream)
    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        lib = _lib.load()
        rc = lib.pfb_comm_unique_id(buf)
        if rc != _lib.PFB_OK:
            raise PfbError(rc, lib.pfb_last_error(None).decode())
        return bytes(buf)

    def comm_init(self, n_ranks: int, rank: int, unique_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self._lib.pfb_comm_init_rank(self._ctx, int(n_ranks), int(rank), buf))

    def stream_handle(self) -> int:
        """cudaStream_t of the context (for timing with the caller's events)"""
        ptr = C.c_void_p()
        self._check(self._lib.pfb_stream(self._ctx, C.byref(ptr)))
        return ptr.value or 0

    # -- output
    def counts(self):
        n, m = C.c_uint64(), C.c_uint64()
        rc = self._lib.pfb_result_counts(self._ctx, C.byref(n), C.byref(m))
        if rc not in (_lib.PFB_OK, _lib.PFB_ENOSHARED):
            self._check(rc)
        return int(n.value), int(m.value)

    def timing(self) -> dict:
        t = _lib.PfbTiming()
        self._check(self._lib.pfb_timings(self._ctx, C.byref(t)))
        return {name: getattr(t, name) for name, _ in t._fields_}

    def records(self) -> np.ndarray:
        n, _ = self.counts()
        out = np.empty(n, dtype=_lib.RECORD_DTYPE)
        self._check(self._lib.pfb_result_records(self._ctx, out.ctypes.data, n))
        return out

    def positions(self, genome_idx: int, picked_only: bool = False) -> np.ndarray:
        n, m = self.counts()
        cnt = m if picked_only else n
        out = np.empty(cnt, dtype=_lib.POS_DTYPE)
        self._check(self._lib.pfb_result_positions(self._ctx, genome_idx, int(picked_only), out.ctypes.data, cnt))
        return out

    def result(self, picked_only: bool = False) -> StageResult:
        rec = self.records()
        if picked_only:
            rec = rec[(rec["flags"] & _lib.F_PICK) != 0]
        pos = [self.positions(g, picked_only) for g in range(self.n_genomes)]
        return StageResult(records=rec, positions=pos, timing=self.timing())

    def result_by_id(self) -> IdView:
        """the hot output path: the per-id columns as zero-copy views of the context's pinned host buffers"""
        v = _lib.PfbIdView()
        self._check(self._lib.pfb_result_by_id(self._ctx, C.byref(v)))
        n, pitch, G = int(v.n_ids), int(v.pitch), int(v.n_genomes)

        def view(ptr, dtype, count):
            if not ptr or count == 0:
                return np.zeros(0, dtype=dtype)
            nbytes = count * np.dtype(dtype).itemsize
            return np.frombuffer((C.c_uint8 * nbytes).from_address(ptr), dtype=dtype, count=count)
        have = bool(v.records_valid)
        pos = view(v.pos, np.uint32, G * pitch).reshape(G, pitch)[:, :n] if pitch else np.zeros((G, 0), np.uint32)
        return IdView(n_ids=n, pitch=pitch, word=view(v.word, np.uint64, n) if have else None, tm=view(v.tm, np.float64, n) if have else None,
                      info=view(v.info, np.uint16, n) if have else None, pos=pos,
                      alive=view(v.alive, np.uint8, n), pick=view(v.pick, np.uint8, n))

    def result_picked(self):
        """the hot output path: (records of the picked k-mers, start|strand<<31 per genome [G, n],
        contig index per genome [G, n] or None when every genome is a single contig); the arrays are
        views of pinned buffers owned by the engine (valid until the next call / close)"""
        _n, m = self.counts()
        G = self.n_genomes
        cap = max(m, 1)
        if self._multi_contig is None:
            self._multi_contig = any(self.genome_info(g)[0] > 1 for g in range(G))
        rec = self._pinned_array("rec", _lib.RECORD_DTYPE, (cap,))
        ss = self._pinned_array("ss", np.uint32, (G, cap))
        ct = self._pinned_array("ct", np.uint32, (G, cap)) if self._multi_contig else None
        self._check(self._lib.pfb_result_picked(self._ctx, rec.ctypes.data, ss.ctypes.data,
                                                ct.ctypes.data if ct is not None else None, cap))
        return rec[:m], ss[:, :m], (ct[:, :m] if ct is not None else None)

    def oligo_tm(self, seq: str):
        tm, gc = C.c_double(), C.c_double()
        self._check(self._lib.pfb_oligo_tm(self._ctx, seq.encode("ascii"), C.byref(tm), C.byref(gc)))
        return tm.value, gc.value


def make_params(**params) -> "_lib.PfbParams":
    p = _lib.default_params()
    for key, val in params.items():
        if not hasattr(p, key):
            raise TypeError(f"unknown parameter {key!r}")
        setattr(p, key, val)
    return p


class MultiEngine:
    """N GPUs driven by ONE host thread of one process (pfb_multi_*): genome 0 goes to every GPU, the others
    round-robin; run() is the whole stage.  engines[i] exposes GPU i's results (genome indices there are local:
    where(g) tells which engine / local index holds global genome g)."""

    def __init__(self, devices, **params):
        self._lib = _lib.load()
        self.params = make_params(**params)
        self.devices = [int(d) for d in devices]
        arr = (C.c_int * len(self.devices))(*self.devices)
        self._m = C.c_void_p()
        rc = self._lib.pfb_multi_create(C.byref(self._m), arr, len(self.devices), C.byref(self.params))
        if rc != _lib.PFB_OK:
            raise PfbError(rc, self._lib.pfb_multi_last_error(None).decode())
        self.engines = []
        for i in range(len(self.devices)):
            ctx = C.c_void_p()
            self._check(self._lib.pfb_multi_context(self._m, i, C.byref(ctx)))
            self.engines.append(Engine(device=self.devices[i], _borrowed_ctx=ctx.value, **params))
        self._keep = []
        self.n_genomes = 0

    def _check(self, rc: int):
        if rc != _lib.PFB_OK:
            raise PfbError(rc, self._lib.pfb_multi_last_error(self._m).decode())

    def _placed(self, multi_contig):
        ci, li = self.where(self.n_genomes)
        targets = self.engines if self.n_genomes == 0 else [self.engines[ci]]
        for e in targets:
            e.n_genomes += 1
            if multi_contig is None:
                e._multi_contig = None
            elif e._multi_contig is not None:
                e._multi_contig |= multi_contig
        self.n_genomes += 1

    def add_genome(self, contigs=None, bases=None, offsets=None, ptr=None, n_bases=None):
        if ptr is None:
            if bases is None:
                contigs = [np.ascontiguousarray(c, dtype=np.uint8) for c in contigs]
                bases = np.concatenate(contigs) if len(contigs) > 1 else (contigs[0] if contigs else np.zeros(0, np.uint8))
                offsets = np.zeros(len(contigs) + 1, dtype=np.uint64)
                offsets[1:] = np.cumsum([c.size for c in contigs])
            bases = np.ascontiguousarray(bases, dtype=np.uint8)
            ptr, n_bases = bases.ctypes.data, bases.size
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._keep.append((bases, offsets))
        self._check(self._lib.pfb_multi_add_genome(self._m, ptr, int(n_bases), offsets.ctypes.data, offsets.size - 1))
        self._placed(offsets.size > 2)

    def add_fasta(self, data=None, ptr=None, n_bytes=None, fmt: str = "fasta"):
        if data is not None:
            data = np.ascontiguousarray(data, dtype=np.uint8)
            ptr, n_bytes = data.ctypes.data, data.size
        self._keep.append((data, None))
        fn = self._lib.pfb_multi_add_fasta if fmt == "fasta" else self._lib.pfb_multi_add_genbank
        self._check(fn(self._m, ptr, int(n_bytes)))
        self._placed(None)

    def where(self, genome_idx: int):
        """(engine index, local genome index) of global genome `genome_idx` (for genome 0: engine 0, it is on all)"""
        if genome_idx == 0:
            return 0, 0
        n = len(self.engines)
        return (genome_idx - 1) % n, 1 + (genome_idx - 1) // n

    def run(self):
        self._check(self._lib.pfb_multi_run(self._m))

    def close(self):
        if getattr(self, "_m", None) and self._m.value:
            for e in self.engines:
                e.close()
            self._lib.pfb_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
