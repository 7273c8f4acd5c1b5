"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports every
function include/bacs_b200.h declares; the host side refuses to run without CUDA tensors."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "bacs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bacs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from bacs_b200 import _cabi
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), "missing export: %s" % name
    # and the Python binding covers the same set
    assert set(names) == set(_cabi.EXPORTED_SYMBOLS), set(names) ^ set(_cabi.EXPORTED_SYMBOLS)


def test_version_and_error_string_without_gpu():
    from bacs_b200 import _cabi
    lib = _cabi.load()
    assert lib.bacs_version() >= 100
    assert isinstance(_cabi.last_error(), str)
    # argument validation happens before any CUDA call
    assert lib.bacs_label_hist(None, 5, None, None) == -1
    assert "null pointer" in _cabi.last_error()
    assert lib.bacs_pixel_loss(None, None, 0, None) == -1


def test_no_cpu_fallback():
    from bacs_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.label_hist(torch.zeros(8, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.confmat_accumulate(torch.zeros(4, dtype=torch.int64), torch.zeros(4, dtype=torch.int64), 2,
                               torch.zeros(2, 2, dtype=torch.int64))


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from bacs_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.BacsError, match="no CPU / PyTorch fallback"):
        _cabi.load()
