"""CPU tests: the C-ABI library loads and exports every symbol include/pvb.h
declares; the Python binding table lists exactly those symbols."""
import ctypes
import os
import re

from pyroved_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pvb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pvb_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exists_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "build with pyroved_b200/csrc/build.sh"
    lib = _lib.lib()
    assert lib.pvb_version() >= 100


def test_every_header_symbol_is_exported():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s


def test_binding_table_matches_header():
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_bad_arguments_are_reported_not_crashed():
    lib = _lib.lib()
    rc = lib.pvb_linear_fwd(None, None, None, None, None, 4, 4, 4, 0, None)
    assert rc == -1
    assert b"pvb_linear_fwd" in lib.pvb_last_error_string()
    cfg = _lib.FoldCfg(1, 1, 2, 0, 128, 0.1, 0.1, 0.1)  # 1-D with rotation: invalid
    rc = lib.pvb_fold_fwd(ctypes.byref(cfg), 8, None, 8, 8, 8, 8, 1, None)
    assert rc == -1
    assert b"1D" in lib.pvb_last_error_string()
