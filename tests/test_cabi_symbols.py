"""CPU: the C-ABI shared library loads and exports every symbol include/gtn_b200.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "gtn_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gtn_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    lib_path = os.path.join(ROOT, "grassmanntn_b200", "libgtn_b200.so")
    assert os.path.exists(lib_path), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in include/gtn_b200.h is not exported" % n


def test_binding_covers_header():
    from grassmanntn_b200 import _cabi
    assert sorted(_cabi.EXPORTED) == _declared()
    assert ctypes.sizeof(_cabi.AxisEntry) == 24
    assert _cabi.lib.gtn_version() >= 100
    assert _cabi.lib.gtn_build_arch() == b"sm_100a"


def test_no_cpu_fallback():
    import torch
    import pytest
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import grassmanntn_b200 as gtn
    with pytest.raises(RuntimeError):
        gtn.dense([[1.0, 0.0], [0.0, 1.0]], statistics=(1, -1))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "grassmanntn_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            for line in src.splitlines():
                if line.strip().startswith(("import ", "from ")):
                    assert "gtn_oracle" not in line and "ref_harness" not in line and "oracle" not in line, (fn, line)
