"""CPU: the C-ABI shared library loads and exports every symbol include/hb200.h declares, and the
ctypes prototype table covers the header one to one (no compute calls here: no GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "hb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge

    ge.build()
    lib = ctypes.CDLL(ge.LIB)
    syms = _header_symbols()
    assert len(syms) > 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_ctypes_table_matches_header():
    import habitat_lab_b200 as hb

    assert sorted(hb._lib.SIGNATURES) == _header_symbols()
    hb.load()  # declares argtypes for every entry; raises on a missing symbol


def test_product_path_never_imports_oracle():
    pkg = os.path.join(ROOT, "habitat-lab_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("test oracle", ""), f"{f} references oracle/"


def test_missing_library_fails_loudly(monkeypatch):
    import habitat_lab_b200 as hb
    import pytest

    monkeypatch.setattr(hb._lib, "_lib", None)
    monkeypatch.setattr(hb._lib, "LIB_PATH", "/nonexistent/libhb200.so")
    with pytest.raises(hb.Hb200Error):
        hb._lib.load()
