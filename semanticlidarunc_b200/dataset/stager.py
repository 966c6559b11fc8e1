"""`ScanStager`: KITTI-style `.bin` / `.label` files -> device tensors through libslu's native I/O threads
(csrc/slu_stager.cu).

The reference reads each scan with two `np.fromfile` calls inside DataLoader worker processes
(src/dataset/dataloader_semantic_KITTI.py:35-39); the arrays then travel pageable -> pinned -> device when the batch
is moved to the GPU.  Here reader threads fill pinned slots ahead of the consumer and `fetch` enqueues the H2D copies on
torch's current stream, so file reads, copies and the projection kernels of earlier scans overlap:

    st = ScanStager(n_slots=8, max_points=150_000)
    tickets = [st.submit(b, l) for b, l in paths]          # returns at once
    for t in tickets:
        xyzi, raw = st.fetch(t)                            # CUDA tensors [n,4] float32 / [n] int32 (uint32 bits)
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from .. import _lib


class ScanStager:
    def __init__(self, n_slots: int = 8, max_points: int = 300_000, n_io_threads: int = 4, device=None):
        self.device = _lib.require_cuda(device)
        self.max_points = int(max_points)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().slu_stager_create(int(n_slots), self.max_points, int(n_io_threads), 0, C.byref(h)), "slu_stager_create")
        self._h = h

    def submit(self, bin_path: str, label_path: Optional[str] = None) -> int:
        t = C.c_int64()
        _lib.check(_lib.lib().slu_stager_submit(self._h, os.fsencode(bin_path), None if label_path is None else os.fsencode(label_path),
                                                C.byref(t)), "slu_stager_submit")
        return int(t.value)

    def fetch_into(self, ticket: int, xyzi: torch.Tensor, raw_label: Optional[torch.Tensor]):
        """Copy the scan into caller-owned device buffers (xyzi [cap,4] float32, raw_label [cap] int32); returns
        (n_points, has_label).  The copies are enqueued on torch's current stream."""
        n, has = C.c_int64(), C.c_int()
        cap = xyzi.size(0) if raw_label is None else min(xyzi.size(0), raw_label.size(0))
        rc = _lib.lib().slu_stager_fetch(self._h, int(ticket), _lib.ptr(xyzi), _lib.ptr(raw_label), int(cap), C.byref(n), C.byref(has),
                                         _lib.stream_ptr())
        _lib.check(rc, "slu_stager_fetch")
        return int(n.value), bool(has.value)

    def fetch(self, ticket: int):
        """-> (xyzi [n,4] float32 CUDA, raw_label [n] int32 CUDA or None)"""
        xyzi = torch.empty((self.max_points, 4), dtype=torch.float32, device=self.device)
        raw = torch.empty((self.max_points,), dtype=torch.int32, device=self.device)
        n, has = self.fetch_into(ticket, xyzi, raw)
        return xyzi[:n], (raw[:n] if has else None)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value is not None:
            _lib.lib().slu_stager_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
