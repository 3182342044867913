"""CPU: the C-ABI shared library loads and exports every symbol include/pyascore_b200.h declares;
the product fails loudly without a GPU and never touches oracle/."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pyascore_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pa_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    from pyascore_b200 import _lib
    assert os.path.exists(_lib.SO_PATH), "build with python pyascore_b200/csrc/build.py"
    dll = ctypes.CDLL(_lib.SO_PATH)
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(dll, n), "library does not export %s" % n
    assert sorted(_lib.EXPORTS) == names          # the ctypes binding covers the whole header


def test_version_and_error_text_without_compute():
    from pyascore_b200 import _lib
    L = _lib.load()
    assert L.pa_version() >= 100
    assert isinstance(_lib.last_error(None), str)


def test_no_cpu_fallback():
    """without a CUDA device construction must raise; with one it must succeed (never silently fall back)"""
    import torch
    from pyascore_b200 import PyAscore
    if torch.cuda.is_available():
        PyAscore(100., 10, "STY", 79.966331)
    else:
        with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
            PyAscore(100., 10, "STY", 79.966331)


def test_bad_arguments_rejected_before_cuda():
    from pyascore_b200 import PyAscore
    with pytest.raises(ValueError):
        PyAscore(100., 9, "STY", 79.966331)            # n_top != 10: reference reads out of bounds
    with pytest.raises(ValueError):
        PyAscore(100., 10, "STY", 79.966331, 0.5, "bx")  # fragment type outside "bcyzZ"


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pyascore_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("the oracle", "").replace("oracle/", "ORACLE_DOC/") or f in ("synth.py",), (dirpath, f)
