"""CPU tests of the C-ABI boundary: the library builds, loads, exports every
symbol include/wiggletools_b200.h declares, and refuses to run without a GPU."""
import ctypes
import os
import re

import pytest

from wiggletools_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wb200_[A-Za-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = declared_symbols("wiggletools_b200.h")
    assert syms, "no declarations parsed"
    assert set(syms) == set(_lib.EXPORTS)
    for s in syms:
        assert hasattr(lib, s), "missing export " + s


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _lib.load()
    assert not lib.wb200_create(0)
    msg = lib.wb200_last_error(None).decode()
    assert "no CPU fallback" in msg


def test_product_never_touches_the_oracle():
    for d, _, files in os.walk(os.path.join(ROOT, "wiggletools_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(d, f)).read()
                assert "pyoracle" not in text and "liboracle" not in text and "libwtref" not in text, f


def test_dropin_symbols_exported():
    """Every function include/wiggletools_dropin.h declares is exported."""
    text = open(os.path.join(ROOT, "include", "wiggletools_dropin.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    body = text[text.index("-- multiplexer.c") if "-- multiplexer.c" in text else 0:]
    names = set(re.findall(r"\b([A-Za-z][A-Za-z0-9_]*)\s*\(", text)) - {"void", "defined", "pop", "seek"}
    names = {n for n in names if n[0].isupper() or n.startswith(("new", "pop", "seek", "run", "destroy", "wb200_"))}
    names |= {"pop", "seek"}
    lib = ctypes.CDLL(_lib.LIB_PATH)
    assert len(names) >= 30 and {"wb200_sink_accepts", "wb200_sink_run", "wb200_mapped_iterator"} <= names
    for n in sorted(names):
        assert hasattr(lib, n), "missing drop-in export " + n


def test_selection_kernel_shared_memory_budgets():
    """the bucketed selection (min / max / median over 25..512 tracks, 1024-bp windows) keeps two CTAs
    resident per SM up to 300 tracks and always fits one; host arithmetic only (no device)."""
    from wiggletools_b200.api import OPS
    lib = _lib.load()
    sm_total, per_cta_reserved, cta_max = 228 * 1024, 1024, 227 * 1024
    for op in ("median", "min", "max"):
        for n in (25, 33, 64, 128, 200, 256, 300):
            b = lib.wb200_debug_window_smem_bytes(OPS[op], n, 1024)
            assert 0 < b and 2 * (b + per_cta_reserved) <= sm_total, (op, n, b)
        for n in (301, 400, 512):
            b = lib.wb200_debug_window_smem_bytes(OPS[op], n, 1024)
            assert 0 < b <= cta_max, (op, n, b)
    # up to 24 tracks: the per-warp sorted arrays, three CTAs per SM at 2048-bp windows
    b = lib.wb200_debug_window_smem_bytes(OPS["median"], 24, 2048)
    assert 0 < b and 3 * (b + per_cta_reserved) <= sm_total, b
