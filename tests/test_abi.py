"""The C-ABI library loads and exports every symbol include/spamtree_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

from common import ROOT


def _declared():
    txt = open(os.path.join(ROOT, "include", "spamtree_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(st_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported():
    from spamtree_b200 import _lib
    names = _declared()
    assert len(names) >= 30
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in spamtree_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)


def test_version_and_error_paths_without_gpu():
    from spamtree_b200 import _lib
    assert b"sm_100a" in _lib.lib.st_version()
    assert _lib.lib.st_create(None, None) == 1  # ST_ERR_INVALID, no crash
    assert _lib.lib.st_get_w(None, None) == 1


def test_product_never_imports_the_oracle():
    """the product package must not reference oracle/ (a product path through the oracle voids parity claims)"""
    pkg = os.path.join(ROOT, "spamtree_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".hpp", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                code = "\n".join(l for l in src.splitlines() if not l.strip().startswith(("//", "#", "*", '"""')))
                assert "import oracle" not in code and "from oracle" not in code and "spamtree_oracle" not in code, f
