"""The C-ABI library loads on a CPU-only box and exports every symbol include/pfb200.h declares."""
import os
import re

import pytest

from primerforge_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "pfb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pfb_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in _declared():
        assert hasattr(lib, name), name


def test_struct_sizes_match_header():
    import ctypes as C
    assert C.sizeof(_lib.PfbParams) == 112
    assert _lib.RECORD_DTYPE.itemsize == 24 and _lib.POS_DTYPE.itemsize == 12
    assert C.sizeof(_lib.PfbTiming) == 6 * 4 + 8 * 8 + 2 * 4 + 4 * 8
    assert C.sizeof(_lib.PfbIdView) == 4 * 8 + 6 * 8


def test_defaults_are_the_reference_defaults():
    p = _lib.default_params()       # bin/Parameters.py:30-52, primer3 calc_tm defaults
    assert (p.k_min, p.k_max, p.table_size, p.max_repeat) == (16, 20, 9999991, 4)
    assert (p.min_gc, p.max_gc, p.min_tm, p.max_tm) == (40.0, 60.0, 55.0, 68.0)
    assert (p.mv_conc, p.dv_conc, p.dntp_conc, p.dna_conc, p.formamide_conc) == (50.0, 1.5, 0.6, 50.0, 0.8)


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from primerforge_b200.engine import Engine, PfbError
    with pytest.raises(PfbError):
        Engine(device=0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "primerforge_b200")
    for dirpath, _d, files in os.walk(pkg):
        for fn in files:
            text = open(os.path.join(dirpath, fn), errors="ignore").read() if fn.endswith((".py", ".cu", ".h")) else ""
            if fn.endswith(".py"):
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), fn
                assert "libpf_oracle" not in text, fn
            elif text:
                assert not re.search(r"#include.*oracle", text), fn
                assert "libpf_oracle" not in text, fn
                # the only library bound at run time is NCCL (more than one GPU)
                assert all("nccl" in lit for lit in re.findall(r'"([^"\n]*\.so\b[^"\n]*)"', text)), fn
