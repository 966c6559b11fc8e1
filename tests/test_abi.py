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
    h = ctypes.c_void_p()
    assert _lib.lib().slu_stager_create(2, 1000, 1, 0, ctypes.byref(h)) != 0 and not h.value      # pinned slots need a device


def test_stager_reader_threads_without_a_gpu(tmp_path):
    """SLU_STAGER_HOST_DEST: the native reader threads, the in-order slot hand-out (more tickets than slots, more
    threads than slots) and the error paths, with host destinations -- no CUDA call involved."""
    import numpy as np
    L = _lib.lib()
    rng = np.random.default_rng(0)
    scans = []
    for i in range(23):
        n = 100 + 37 * i
        xyzi = rng.standard_normal((n, 4)).astype(np.float32)
        lab = rng.integers(0, 2 ** 32, n, dtype=np.uint32)
        b, l = tmp_path / f"{i}.bin", tmp_path / f"{i}.label"
        xyzi.tofile(b); lab.tofile(l)
        scans.append((str(b).encode(), str(l).encode(), xyzi, lab))
    (tmp_path / "bad.bin").write_bytes(b"x" * 30)
    for n_slots, n_threads in ((1, 1), (2, 5), (4, 2)):
        h = ctypes.c_void_p()
        assert L.slu_stager_create(n_slots, 2000, n_threads, 1, ctypes.byref(h)) == 0
        tickets = []
        for b, l, _, _ in scans:
            t = ctypes.c_int64()
            assert L.slu_stager_submit(h, b, l, ctypes.byref(t)) == 0
            tickets.append(t.value)
        t = ctypes.c_int64()
        L.slu_stager_submit(h, str(tmp_path / "bad.bin").encode(), None, ctypes.byref(t)); bad = t.value
        L.slu_stager_submit(h, str(tmp_path / "missing.bin").encode(), None, ctypes.byref(t)); missing = t.value
        L.slu_stager_submit(h, scans[0][0], scans[1][1], ctypes.byref(t)); mismatch = t.value
        L.slu_stager_submit(h, scans[3][0], None, ctypes.byref(t)); nolabel = t.value
        assert tickets == list(range(len(scans)))
        dx, dl = np.empty((2000, 4), np.float32), np.empty(2000, np.uint32)
        n, has = ctypes.c_int64(), ctypes.c_int()
        for tk, (_, _, xyzi, lab) in zip(tickets, scans):
            rc = L.slu_stager_fetch(h, tk, dx.ctypes.data_as(ctypes.c_void_p), dl.ctypes.data_as(ctypes.c_void_p), 2000,
                                    ctypes.byref(n), ctypes.byref(has), None)
            assert rc == 0 and n.value == xyzi.shape[0] and has.value == 1
            assert np.array_equal(dx[:n.value].view(np.uint32), xyzi.view(np.uint32)) and np.array_equal(dl[:n.value], lab)
        for tk in (bad, missing, mismatch):
            rc = L.slu_stager_fetch(h, tk, dx.ctypes.data_as(ctypes.c_void_p), None, 2000, ctypes.byref(n), ctypes.byref(has), None)
            assert rc == -5, L.slu_last_error()
            with pytest.raises(OSError):
                _lib.check(rc, "slu_stager_fetch")
        rc = L.slu_stager_fetch(h, nolabel, dx.ctypes.data_as(ctypes.c_void_p), None, 2000, ctypes.byref(n), ctypes.byref(has), None)
        assert rc == 0 and has.value == 0 and n.value == scans[3][2].shape[0]
        assert L.slu_stager_fetch(h, 10 ** 6, dx.ctypes.data_as(ctypes.c_void_p), None, 2000, ctypes.byref(n), ctypes.byref(has), None) == -1
        assert L.slu_stager_destroy(h) == 0
    # destroy with unfetched tickets must not hang
    h = ctypes.c_void_p()
    assert L.slu_stager_create(1, 2000, 3, 1, ctypes.byref(h)) == 0
    for b, l, _, _ in scans[:5]:
        L.slu_stager_submit(h, b, l, ctypes.byref(ctypes.c_int64()))
    assert L.slu_stager_destroy(h) == 0
