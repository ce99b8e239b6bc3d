"""CPU: the C-ABI library loads and exports every symbol include/byol_b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "byol_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(byol_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from byol_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in syms:
        assert hasattr(raw, name), "libbyol_b200.so does not export %s" % name
    # and the Python binding table covers exactly the header
    assert sorted(_lib.EXPORTED_SYMBOLS) == syms
    assert _lib.lib.byol_abi_version() == 2
    assert _lib.last_error() == ""


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under byol_b200/ may import or reference it."""
    pkg = os.path.join(ROOT, "byol_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn
