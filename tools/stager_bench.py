#!/usr/bin/env python
"""Files -> device tensors: np.fromfile + torch .to(device) (what the reference's loaders do before the GPU sees a scan)
against libslu's native stager (reader threads -> pinned slots -> async H2D), on 128 HDL-64-sized scans written to a
temporary directory (page cache warm: this measures the software path, not the disk).  Wall clock, best of 3."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semanticlidarunc_b200 import synth  # noqa: E402
from semanticlidarunc_b200.dataset.stager import ScanStager  # noqa: E402

dev = torch.device("cuda", 0)
N = 128


def timed(fn):
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


with tempfile.TemporaryDirectory() as d:
    paths = []
    for i in range(N):
        xyzi, raw = synth.synth_scan(i % 8, "hdl64")
        b, l = os.path.join(d, f"{i:06d}.bin"), os.path.join(d, f"{i:06d}.label")
        xyzi.tofile(b); raw.tofile(l)
        paths.append((b, l))
    nbytes = sum(os.path.getsize(b) + os.path.getsize(l) for b, l in paths)

    def numpy_path():
        outs = []
        for b, l in paths:
            x = torch.from_numpy(np.fromfile(b, dtype=np.float32).reshape(-1, 4)).to(dev, non_blocking=True)
            r = torch.from_numpy(np.fromfile(l, dtype=np.uint32).view(np.int32)).to(dev, non_blocking=True)
            outs.append((x, r))
        torch.cuda.synchronize()
        return outs

    def staged(threads):
        st = ScanStager(n_slots=16, max_points=130_000, n_io_threads=threads, device=dev)     # created once, as a loader would
        xyzi = torch.empty((130_000, 4), dtype=torch.float32, device=dev)
        raw = torch.empty((130_000,), dtype=torch.int32, device=dev)

        def run():
            tickets = [st.submit(b, l) for b, l in paths]
            for t in tickets:
                st.fetch_into(t, xyzi, raw)
            torch.cuda.synchronize()
        return run

    res = {}
    for name, fn in (("np.fromfile + .to(device)", numpy_path), ("stager, 1 reader thread", staged(1)), ("stager, 4 reader threads", staged(4)),
                     ("stager, 8 reader threads", staged(8))):
        fn()
        best = min(timed(fn) for _ in range(3))
        res[name] = {"scans_per_s": round(N / best, 1), "GBps": round(nbytes / best / 1e9, 2)}
    print(json.dumps({"scans": N, "bytes": nbytes, "cpus": os.cpu_count(), "results": res}, indent=1))
