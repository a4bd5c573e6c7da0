"""ctypes binding of libpfb200.so (include/pfb200.h).

The CUDA library is the product: there is no CPU path.  Importing this module never touches the
GPU; ``load()`` raises ``RuntimeError`` when the shared library has not been built and every
call that needs a device raises when CUDA is unavailable.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PFB200_LIB") or os.path.join(_HERE, "libpfb200.so")   # PFB200_LIB: A/B builds
SRC_PATH = os.path.join(_HERE, "csrc", "pfb200.cu")
SRC_PATHS = [SRC_PATH, os.path.join(_HERE, "csrc", "pfb_outgroup.cu"), os.path.join(_HERE, "csrc", "pfb_pairs.cu")]      # one translation unit per stage
INCLUDE_DIR = os.path.join(os.path.dirname(_HERE), "include")

PFB_OK, PFB_EINVAL, PFB_ECUDA, PFB_ESTATE, PFB_ENOSHARED, PFB_ETOOLARGE = 0, -1, -2, -3, -4, -5
OPT_SCAN_MODE, OPT_FETCH_RECORDS, OPT_SHARD_FIRST, OPT_EMULATE_RANKS = 0, 1, 2, 3
F_GC_OK, F_TM_OK, F_REP_OK, F_PICK, F_PALIN = 1, 2, 4, 8, 16

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]
NVCC_LIBS = ["-ldl"]      # NCCL is bound at run time (dlopen), only when more than one GPU takes part

# every symbol include/pfb200.h declares
EXPORTS = ["pfb_create", "pfb_destroy", "pfb_last_error", "pfb_default_params", "pfb_add_genome",
           "pfb_add_fasta", "pfb_add_genbank", "pfb_genome_info", "pfb_contig_table",
           "pfb_clear_genomes", "pfb_upload", "pfb_upload_async", "pfb_upload_wait", "pfb_compute", "pfb_run",
           "pfb_comm_unique_id", "pfb_comm_init_rank", "pfb_set_option",
           "pfb_multi_create", "pfb_multi_destroy", "pfb_multi_last_error", "pfb_multi_add_genome", "pfb_multi_add_fasta",
           "pfb_multi_add_genbank", "pfb_multi_run", "pfb_multi_size", "pfb_multi_context", "pfb_multi_locate",
           "pfb_sync", "pfb_stream", "pfb_result_counts", "pfb_result_by_id", "pfb_result_records",
           "pfb_result_positions", "pfb_result_picked", "pfb_host_alloc", "pfb_host_free", "pfb_timings", "pfb_set_scan_mode", "pfb_oligo_tm",
           "pfb_og_create", "pfb_og_destroy", "pfb_og_last_error", "pfb_og_set_keys", "pfb_og_set_pairs", "pfb_og_add_genome",
           "pfb_og_clear_genomes", "pfb_og_set_scan_mode", "pfb_og_run", "pfb_og_compute", "pfb_og_counts", "pfb_og_hits", "pfb_og_products",
           "pfb_og_removed_at", "pfb_og_timings",
           "pfb_pp_create", "pfb_pp_destroy", "pfb_pp_last_error", "pfb_pp_set_candidates", "pfb_pp_add_genome",
           "pfb_pp_clear_genomes", "pfb_pp_run", "pfb_pp_counts", "pfb_pp_pairs", "pfb_pp_shared", "pfb_pp_kept", "pfb_pp_bins", "pfb_pp_timings"]


class PfbParams(C.Structure):
    _fields_ = [("k_min", C.c_int32), ("k_max", C.c_int32), ("table_size", C.c_uint32), ("max_repeat", C.c_int32),
                ("min_gc", C.c_double), ("max_gc", C.c_double), ("min_tm", C.c_double), ("max_tm", C.c_double),
                ("mv_conc", C.c_double), ("dv_conc", C.c_double), ("dntp_conc", C.c_double), ("dna_conc", C.c_double),
                ("dmso_conc", C.c_double), ("dmso_fact", C.c_double), ("formamide_conc", C.c_double),
                ("prefilter", C.c_int32), ("reserved", C.c_int32)]


class PfbTiming(C.Structure):
    _fields_ = [("ms_upload", C.c_float), ("ms_encode", C.c_float), ("ms_first", C.c_float), ("ms_scan", C.c_float),
                ("ms_finalize", C.c_float), ("ms_total", C.c_float), ("n_bases", C.c_uint64), ("n_windows", C.c_uint64),
                ("n_slots", C.c_uint64), ("n_records", C.c_uint64), ("n_picked", C.c_uint64), ("n_launches", C.c_uint64),
                ("used_filter", C.c_uint64), ("scan_main_bases", C.c_uint64), ("ms_scan_main", C.c_float),
                ("ms_exchange", C.c_float), ("n_ids", C.c_uint64), ("sharded_first_genome", C.c_uint64),
                ("n_reruns", C.c_uint64), ("exchange_mode", C.c_uint64)]


class PfbIdView(C.Structure):
    _fields_ = [("n_ids", C.c_uint64), ("pitch", C.c_uint64), ("n_genomes", C.c_uint64), ("records_valid", C.c_uint64),
                ("word", C.c_void_p), ("tm", C.c_void_p), ("info", C.c_void_p), ("pos", C.c_void_p),
                ("alive", C.c_void_p), ("pick", C.c_void_p)]


class PfbOgTiming(C.Structure):
    _fields_ = [("ms_upload", C.c_float), ("ms_encode", C.c_float), ("ms_scan", C.c_float), ("ms_bucket", C.c_float),
                ("ms_products", C.c_float), ("ms_total", C.c_float), ("n_bases", C.c_uint64), ("n_windows", C.c_uint64),
                ("n_hits", C.c_uint64), ("n_products", C.c_uint64), ("n_probe_occupied", C.c_uint64),
                ("n_launches", C.c_uint64), ("n_reruns", C.c_uint64), ("table_bytes", C.c_uint64),
                ("filter_in_smem", C.c_uint64), ("filter_bits_per_record", C.c_uint64)]


class PfbPpParams(C.Structure):
    _fields_ = [("max_bin", C.c_int32), ("min_primer_len", C.c_int32), ("min_pcr", C.c_int32), ("max_pcr", C.c_int32),
                ("max_tm_diff", C.c_double)]


class PfbPpTiming(C.Structure):
    _fields_ = [("ms_bin", C.c_float), ("ms_pairs", C.c_float), ("ms_shared", C.c_float), ("ms_total", C.c_float),
                ("n_candidates", C.c_uint64), ("n_genomes", C.c_uint64), ("n_bins", C.c_uint64), ("n_pairs", C.c_uint64),
                ("n_kept", C.c_uint64), ("n_launches", C.c_uint64)]


PP_PAIR_DTYPE = np.dtype([("fwd", "<u4"), ("rev", "<u4"), ("contig", "<u4"), ("fbin", "<u4"), ("rbin", "<u4"), ("pcr_len", "<u4")])
PP_VALID = 0x80000000
OG_HIT_DTYPE = np.dtype([("key", "<u4"), ("genome", "<u4"), ("contig", "<u4"), ("start", "<u4"), ("strand", "<u4")])
OG_PRODUCT_DTYPE = np.dtype([("pair", "<u4"), ("genome", "<u4"), ("contig", "<u4"), ("size", "<i4")])
RECORD_DTYPE = np.dtype([("word", "<u8"), ("tm", "<f8"), ("slot", "<u4"), ("gc_count", "<u2"), ("k", "u1"), ("flags", "u1")])
POS_DTYPE = np.dtype([("contig", "<u4"), ("start", "<u4"), ("strand", "<u4")])
assert RECORD_DTYPE.itemsize == 24 and POS_DTYPE.itemsize == 12


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> primerforge_b200/libpfb200.so (in tree)"""
    # every file the library is compiled from: the sources under csrc/ (pfb200.cu includes the .cuh files) and the header
    srcs = [os.path.join(os.path.dirname(SRC_PATH), f) for f in sorted(os.listdir(os.path.dirname(SRC_PATH)))]
    srcs.append(os.path.join(INCLUDE_DIR, "pfb200.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(f) > os.path.getmtime(LIB_PATH) for f in srcs)
    if force or stale:
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SRC_PATHS + NVCC_LIBS
        subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def load():
    """dlopen libpfb200.so; fails loudly when it has not been built (no fallback exists)"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a); primerforge_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    lib.pfb_create.argtypes = [C.POINTER(vp), i32, C.POINTER(PfbParams)]
    lib.pfb_destroy.argtypes = [vp]
    lib.pfb_destroy.restype = None
    lib.pfb_last_error.argtypes = [vp]
    lib.pfb_last_error.restype = C.c_char_p
    lib.pfb_default_params.argtypes = [C.POINTER(PfbParams)]
    lib.pfb_add_genome.argtypes = [vp, vp, u64, vp, u32]
    lib.pfb_add_fasta.argtypes = [vp, vp, u64]
    lib.pfb_add_genbank.argtypes = [vp, vp, u64]
    lib.pfb_genome_info.argtypes = [vp, u32, C.POINTER(u32), C.POINTER(u64)]
    lib.pfb_contig_table.argtypes = [vp, u32, vp, vp, u32]
    lib.pfb_clear_genomes.argtypes = [vp]
    for fn in ("pfb_upload", "pfb_upload_async", "pfb_upload_wait", "pfb_compute", "pfb_run", "pfb_sync"):
        getattr(lib, fn).argtypes = [vp]
    lib.pfb_comm_unique_id.argtypes = [vp]
    lib.pfb_comm_init_rank.argtypes = [vp, i32, i32, vp]
    lib.pfb_set_option.argtypes = [vp, i32, i32]
    lib.pfb_multi_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), i32, C.POINTER(PfbParams)]
    lib.pfb_multi_destroy.argtypes = [vp]
    lib.pfb_multi_destroy.restype = None
    lib.pfb_multi_last_error.argtypes = [vp]
    lib.pfb_multi_last_error.restype = C.c_char_p
    lib.pfb_multi_add_genome.argtypes = [vp, vp, u64, vp, u32]
    lib.pfb_multi_add_fasta.argtypes = [vp, vp, u64]
    lib.pfb_multi_add_genbank.argtypes = [vp, vp, u64]
    lib.pfb_multi_run.argtypes = [vp]
    lib.pfb_multi_size.argtypes = [vp]
    lib.pfb_multi_context.argtypes = [vp, i32, C.POINTER(vp)]
    lib.pfb_multi_locate.argtypes = [vp, u32, C.POINTER(C.c_int), C.POINTER(u32)]
    lib.pfb_result_by_id.argtypes = [vp, C.POINTER(PfbIdView)]
    lib.pfb_stream.argtypes = [vp, C.POINTER(vp)]
    lib.pfb_result_counts.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    lib.pfb_result_records.argtypes = [vp, vp, u64]
    lib.pfb_result_positions.argtypes = [vp, u32, i32, vp, u64]
    lib.pfb_result_picked.argtypes = [vp, vp, vp, vp, u64]
    lib.pfb_host_alloc.argtypes = [C.POINTER(vp), u64]
    lib.pfb_host_free.argtypes = [vp]
    lib.pfb_timings.argtypes = [vp, C.POINTER(PfbTiming)]
    lib.pfb_set_scan_mode.argtypes = [vp, i32]
    lib.pfb_oligo_tm.argtypes = [vp, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.pfb_og_create.argtypes = [C.POINTER(vp), i32]
    lib.pfb_og_destroy.argtypes = [vp]
    lib.pfb_og_destroy.restype = None
    lib.pfb_og_last_error.argtypes = [vp]
    lib.pfb_og_last_error.restype = C.c_char_p
    lib.pfb_og_set_keys.argtypes = [vp, vp, vp, u32]
    lib.pfb_og_set_pairs.argtypes = [vp, vp, vp, u32, C.c_int32, C.c_int32]
    lib.pfb_og_add_genome.argtypes = [vp, vp, u64, vp, u32]
    for fn in ("pfb_og_clear_genomes", "pfb_og_set_scan_mode", "pfb_og_run", "pfb_og_compute"):
        getattr(lib, fn).argtypes = [vp]
    lib.pfb_og_set_scan_mode.argtypes = [vp, i32]
    lib.pfb_og_counts.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
    lib.pfb_og_hits.argtypes = [vp, vp, u64]
    lib.pfb_og_products.argtypes = [vp, vp, u64]
    lib.pfb_og_removed_at.argtypes = [vp, vp, u64]
    lib.pfb_og_timings.argtypes = [vp, C.POINTER(PfbOgTiming)]
    lib.pfb_pp_create.argtypes = [C.POINTER(vp), i32]
    lib.pfb_pp_destroy.argtypes = [vp]
    lib.pfb_pp_destroy.restype = None
    lib.pfb_pp_last_error.argtypes = [vp]
    lib.pfb_pp_last_error.restype = C.c_char_p
    lib.pfb_pp_set_candidates.argtypes = [vp, vp, vp, vp, u64]
    lib.pfb_pp_add_genome.argtypes = [vp, vp, vp, vp, vp, u64]
    lib.pfb_pp_clear_genomes.argtypes = [vp]
    lib.pfb_pp_run.argtypes = [vp, C.POINTER(PfbPpParams)]
    lib.pfb_pp_counts.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
    lib.pfb_pp_pairs.argtypes = [vp, vp, u64]
    lib.pfb_pp_shared.argtypes = [vp, vp, u64]
    lib.pfb_pp_kept.argtypes = [vp, vp, u64]
    lib.pfb_pp_bins.argtypes = [vp, u32, vp, vp, u64]
    lib.pfb_pp_timings.argtypes = [vp, C.POINTER(PfbPpTiming)]
    for name in EXPORTS:
        if name not in ("pfb_destroy", "pfb_last_error", "pfb_multi_destroy", "pfb_multi_last_error", "pfb_og_destroy",
                        "pfb_og_last_error", "pfb_pp_destroy", "pfb_pp_last_error"):
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def default_params() -> PfbParams:
    p = PfbParams()
    rc = load().pfb_default_params(C.byref(p))
    if rc != PFB_OK:
        raise RuntimeError("pfb_default_params failed")
    return p
