"""Scratch: compare the cell pipeline with the four-launch path on one scan and print the pixels that differ."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semanticlidarunc_b200 import ops, synth, _lib
from semanticlidarunc_b200.dataset.definitions import build_id_lut
dev = torch.device("cuda", 0)
xyzi, raw = synth.synth_scan(0, "hdl64")
dx = torch.from_numpy(xyzi).to(dev); dr = torch.from_numpy(raw.view(np.int32)).to(dev); dl = torch.from_numpy(build_id_lut()).to(dev)
offs = [0, xyzi.shape[0]]
out = {}
for mode in (2, 0, 2):
    prev = _lib.lib().slu_debug_project_exact(mode)
    out[mode] = ops.project_batch(dx, dr, offs, 64, 2048, lut=dl)
    _lib.lib().slu_debug_project_exact(prev)
a, b = out[2], out[0]
print("pix equal", torch.equal(a["pix"], b["pix"]))
wa, wb = a["winner"].reshape(-1).cpu().numpy(), b["winner"].reshape(-1).cpu().numpy()
d = np.nonzero(wa != wb)[0]
print("winner diffs", d.size, "of", wa.size, "occupied", (wb >= 0).sum())
pix = a["pix"].cpu().numpy()
x = xyzi.astype(np.float64)
r2 = x[:, 0] ** 2 + x[:, 1] ** 2 + x[:, 2] ** 2
for px in d[:12]:
    members = np.nonzero(pix == px)[0]
    print("px", px, "cells", wa[px], "four", wb[px], "members", members, "r2", r2[members], "bits", [hex(v) for v in r2[members].view(np.uint64)])
