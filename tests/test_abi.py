"""The C-ABI library loads and exports exactly what include/vtb.h declares (no GPU needed: no compute calls)."""
import ctypes as C
import re
from pathlib import Path

import pytest

from vision_toolbox_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "vtb.h").read_text()


def header_symbols() -> list[str]:
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(vtb_[a-z0-9_]+)\s*\(", body)))


def test_header_and_binding_table_agree():
    assert header_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    h = _lib.lib()
    for name in header_symbols():
        assert hasattr(h, name), f"{name} declared in include/vtb.h but not exported"


def test_argument_counts_match_header():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", body, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), (name, n, len(args))


def test_host_only_queries_and_error_codes():
    h = _lib.lib()
    assert h.vtb_version() >= 100
    g = _lib.VtbConv(2, 11, 11, 32, 32, 3, 2, 1)
    ho, wo = C.c_int(), C.c_int()
    assert h.vtb_conv_out_hw(C.byref(g), C.byref(ho), C.byref(wo)) == 0
    assert (ho.value, wo.value) == (6, 6)          # 11 -> 6 with k3 s2 p1 (SURVEY.md odd-size case)
    bad = _lib.VtbConv(2, 11, 11, 30, 32, 3, 2, 1)  # channels not a multiple of 16
    assert h.vtb_conv_out_hw(C.byref(bad), C.byref(ho), C.byref(wo)) == -1
    assert b"bad geometry" in h.vtb_last_error()
    with pytest.raises(_lib.VtbError):
        _lib.check(-1, "probe")
    assert h.vtb_bn_bwd_rows(1000, 64) >= 1
    assert h.vtb_bn_bwd_rows(1000, 63) < 0
