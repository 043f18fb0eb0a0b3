"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol that
include/plda_b200.h declares, the ctypes table matches the header, the product never imports the
oracle, and without a GPU the product fails loudly instead of falling back."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "plda_b200.h")


@pytest.fixture(scope="module")
def lib():
    from plda_b200 import build, _ffi
    build.build()
    return _ffi.lib()


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"^\s*(?:int|const char\*)\s+((?:plda|lda)_[a-z0-9_]+)\s*\(", src, flags=re.M)))


def test_library_exports_every_declared_symbol(lib):
    names = header_functions()
    assert len(names) > 30
    for n in names:
        assert hasattr(lib, n), "libplda_b200.so does not export %s" % n


def test_ctypes_table_matches_header(lib):
    from plda_b200 import _ffi
    declared = set(header_functions())
    bound = set(_ffi.SIGNATURES) | set(_ffi.STRING_FUNCS)
    assert declared == bound, (declared - bound, bound - declared)
    # argument counts agree with the header prototypes
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, args in _ffi.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, src, flags=re.S)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), name


def test_version_string(lib):
    assert b"sm_100a" in lib.plda_version()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "plda_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert "oracle/" not in text or f == "_ffi.py", f


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from plda_b200 import PLDA, LDA, _ffi
    with pytest.raises(_ffi.PldaB200Error) as e:
        PLDA()
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    with pytest.raises(_ffi.PldaB200Error):
        LDA()


def test_input_validation_matches_reference_errors():
    from plda_b200 import _ffi
    with pytest.raises(ValueError):
        _ffi.as_matrix(np.zeros((3, 2), dtype=np.int32))            # src/pldamodule.cpp:59-62
    with pytest.raises(ValueError):
        _ffi.as_labels(np.array(["a", "b"]))                         # :128-131
    with pytest.raises(ValueError):
        _ffi.as_labels(np.array([-1, 2]))                            # :55-58 (not unsigned)
    with pytest.raises(ValueError):
        _ffi.as_labels(np.array([0.5, 1.0]))
    assert _ffi.as_labels(np.array([3, 1], dtype="uint")).dtype == np.uint64
    assert _ffi.as_labels(np.array([3, 1], dtype=np.int64)).dtype == np.uint64   # superset
    with pytest.raises(ValueError):
        _ffi.as_labels(np.array([3, 1], dtype=np.int64), require_unsigned=True)


def test_liblda_alias_exports_reference_names():
    import liblda
    assert liblda.__all__ == ["PLDA", "LDA"]          # python/liblda/__init__.py:3
    for name in ("fit", "transform", "norm", "score"):               # python/liblda/plda.py:9-51
        assert callable(getattr(liblda.PLDA, name))
    for name in ("fit", "decision_function", "predict_proba", "predict_log_proba"):   # lda.py:106-325
        assert callable(getattr(liblda.LDA, name))
