"""CPU tests: the C-ABI library builds, loads and exports every symbol include/dexdeform_mpm.h declares; the ctypes
tables cover the reference's binding (mpm/types.py:103-290).  No compute calls (no GPU here)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dexdeform_mpm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:[A-Za-z_][\w\s\*]*?[\s\*])([a-z_][a-z0-9_]*)\s*\(", text, flags=re.M)
    return sorted({n for n in names if n not in ("defined",)})


def test_header_declares_reference_abi():
    from dexdeform_b200.types import ABI1
    syms = declared_symbols()
    for name in ABI1:
        assert name in syms, f"{name} bound by ctypes but not declared in include/dexdeform_mpm.h"
    assert len(syms) >= 30


def test_library_exports_every_declared_symbol(product_lib):
    from dexdeform_b200.types import LIB_PATH
    raw = ctypes.cdll.LoadLibrary(LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(raw, s)]
    assert not missing, missing


def test_missing_library_fails_loudly(tmp_path):
    import pytest
    from dexdeform_b200.types import load_library
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        load_library(str(tmp_path / "nope.so"))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dexdeform_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} references oracle/"
