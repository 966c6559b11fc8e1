"""The C-ABI library loads on a CPU-only box and exports every symbol include/slu.h declares;
argument validation (which happens before any CUDA call) reports through slu_last_error()."""
import ctypes
import os
import re

import pytest

from semanticlidarunc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "slu.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(slu_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 10
    h = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/slu.h but not exported by libslu.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.SIGNATURES"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_string():
    L = _lib.lib()
    assert L.slu_version() == 100
    rc = L.slu_confusion_ece(None, None, None, 5, 20, 0, 0, 0, None, None, None, None)
    assert rc == -1 and b"NULL" in L.slu_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, "slu_confusion_ece")
    rc = L.slu_reduce_metrics(ctypes.c_void_p(16), None, 2, 1, 40, 64, 0, 0, 1e-12, 1, 0, 0, 0, None,
                              None, None, None, None, None, None, None, None)
    assert rc == -2 and b"C=40" in L.slu_last_error()
    assert L.slu_project_workspace_bytes(120000, 1, 64 * 2048) > 120000 * 24


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from semanticlidarunc_b200 import ops
    with pytest.raises(_lib.SluError):
        ops.reduce_metrics(torch.zeros((1, 1, 20, 2, 8)), None)
    from semanticlidarunc_b200.dataset.utils import spherical_projection
    import numpy as np
    with pytest.raises(_lib.SluError):
        spherical_projection(np.zeros((4, 5)))
    n = ctypes.c_int()
    assert _lib.lib().slu_device_info(0, ctypes.byref(n), None, None) == -4
