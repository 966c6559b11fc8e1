"""Parity of the projection / back-projection kernels: indices, winners, images and labels are
compared BIT-EXACTLY with the reference's golden vectors (tests/golden/projection_small.npz,
kitti_loader.npz, MANIFEST.json digests) and with the oracle on seeded scans."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import projection as oproj
from semanticlidarunc_b200 import ops, synth
from semanticlidarunc_b200.dataset.definitions import build_id_lut
from semanticlidarunc_b200.dataset.utils import spherical_projection

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def to_dev(xyzi, raw, cuda):
    return (torch.from_numpy(xyzi).to(cuda), torch.from_numpy(raw.view(np.int32)).to(cuda),
            torch.from_numpy(build_id_lut()).to(cuda))


def hwc_from_planes(img6):
    """[6,H,W] planes (x,y,z,range,intensity,label) -> the reference's [H,W,5] (x,y,z,i,label)."""
    p = img6.cpu().numpy()
    return np.stack([p[0], p[1], p[2], p[4], p[5]], axis=-1)


SMALL = ["tiny_auto", "tiny_range", "tiny_farthest", "edge", "ragged_1pt"]


@pytest.mark.parametrize("name", SMALL)
def test_batch_projection_vs_reference_golden(cuda, golden, name):
    g = golden("projection_small.npz")
    xyzi, raw = g[name + "/xyzi"], g[name + "/raw"]
    H, W = (int(v) for v in g[name + "/hw"])
    tr = g[name + "/theta_range"]
    tr = None if np.isnan(tr).any() else (float(tr[0]), float(tr[1]))
    far = bool(int(g[name + "/largest_first"]))
    dx, dr, dl = to_dev(xyzi, raw, cuda)
    res = ops.project_batch(dx, dr, [0, xyzi.shape[0]], H, W, lut=dl, theta_range=tr, farthest_wins=far)
    torch.cuda.synchronize()
    assert np.array_equal(res["pix"].cpu().numpy().astype(np.int64), g[name + "/pix"])
    assert np.array_equal(res["winner"].cpu().numpy().reshape(-1).astype(np.int64), g[name + "/winner"])
    assert np.array_equal(hwc_from_planes(res["img"][0]), g[name + "/img"])          # bit-exact float32 image
    if tr is None:
        assert np.array_equal(res["theta"][0].cpu().numpy(), g[name + "/theta"])
    assert int(res["diag"][0, 0]) == 0


@pytest.mark.parametrize("name", SMALL)
def test_generic_projection_vs_reference_golden(cuda, golden, name):
    g = golden("projection_small.npz")
    xyzi, raw = g[name + "/xyzi"], g[name + "/raw"]
    H, W = (int(v) for v in g[name + "/hw"])
    tr = g[name + "/theta_range"]
    tr = None if np.isnan(tr).any() else (float(tr[0]), float(tr[1]))
    far = bool(int(g[name + "/largest_first"]))
    sem = build_id_lut()[(raw & 0xFFFF).astype(np.int64)].astype(np.int64)
    pc = np.concatenate([xyzi, sem[:, None]], axis=-1)                 # float64 [N,5] as the loaders build it
    img, alpha, (tmin, tmax), (pmin, pmax) = spherical_projection(pc, H, W, theta_range=tr, sort_largest_first=far)
    assert img.dtype == np.float32 and img.shape == (H, W, 5)
    assert np.array_equal(img, g[name + "/img"])
    assert sha(alpha) == bytes(g[name + "/alpha_sha"]).hex()
    assert (pmin, pmax) == (-np.pi, np.pi)
    if tr is None:
        assert np.array_equal(np.array([tmin, tmax]), g[name + "/theta"])


def test_full_size_digests(cuda):
    """HDL-64 (120k pts -> 64x2048) and OS1-128 (262k pts -> 128x2048) against reference digests."""
    with open(os.path.join(HERE, "golden", "MANIFEST.json")) as f:
        cases = json.load(f)["projection_full"]
    for name, c in cases.items():
        xyzi, raw = synth.synth_scan(c["seed"], c["sensor"])
        assert sha(xyzi) == c["xyzi_sha"] and sha(raw) == c["raw_sha"], "synthetic generator drifted"
        dx, dr, dl = to_dev(xyzi, raw, cuda)
        tr = None if c["theta_range"] is None else tuple(c["theta_range"])
        res = ops.project_batch(dx, dr, [0, xyzi.shape[0]], c["H"], c["W"], lut=dl, theta_range=tr)
        assert sha(res["pix"].cpu().numpy().astype(np.int64)) == c["pix_sha"], name
        assert sha(res["winner"].cpu().numpy().reshape(-1).astype(np.int64)) == c["winner_sha"], name
        assert sha(hwc_from_planes(res["img"][0])) == c["img_sha"], name
        assert int((res["winner"] >= 0).sum()) == c["occupied"]
        if tr is None:
            assert float(res["theta"][0, 0]) == c["theta_min"] and float(res["theta"][0, 1]) == c["theta_max"]


def test_kitti_loader_planes_vs_reference(cuda, golden):
    g = golden("kitti_loader.npz")
    H, W = (int(v) for v in g["hw"])
    dx, dr, dl = to_dev(g["xyzi"], g["raw"], cuda)
    img = ops.project_batch(dx, dr, [0, g["xyzi"].shape[0]], H, W, lut=dl)["img"][0].cpu().numpy()
    assert np.array_equal(img[0:3], g["xyz"])
    assert np.array_equal(img[3:4], g["range"])                  # fp32 norm, bit-exact
    assert np.array_equal(img[4:5], g["reflectivity"])
    assert np.array_equal(img[5:6].astype(np.int64), g["semantics"])


def test_ragged_batch_vs_oracle_and_backprojection(cuda):
    """Batch of ragged scans incl. an empty one; back-projected labels are bit-exact vs the oracle."""
    lut = build_id_lut()
    scans = [synth.synth_scan(31, "tiny"), synth.synth_scan(32, "tiny", n_points=1234),
             (np.zeros((0, 4), np.float32), np.zeros((0,), np.uint32)), synth.synth_scan(33, "tiny", n_points=17)]
    H, W = 16, 256
    offs = np.concatenate([[0], np.cumsum([s[0].shape[0] for s in scans])])
    xyzi = np.concatenate([s[0] for s in scans])
    raw = np.concatenate([s[1] for s in scans])
    dx, dr, dl = to_dev(xyzi, raw, cuda)
    res = ops.project_batch(dx, dr, offs, H, W, lut=dl)
    rng = np.random.default_rng(0)
    label_img = torch.from_numpy(rng.integers(0, 20, (len(scans), H, W))).to(cuda)
    back = ops.backproject(label_img, res["pix"], offs).cpu().numpy()
    for b, (a, r) in enumerate(scans):
        if a.shape[0] == 0:
            assert int((res["winner"][b] >= 0).sum()) == 0 and float(res["img"][b].abs().sum()) == 0.0
            continue
        o = oproj.kitti_frame(a, r, H, W, lut)
        sl = slice(offs[b], offs[b + 1])
        assert np.array_equal(res["pix"][sl].cpu().numpy().astype(np.int64), o["pix"])
        assert np.array_equal(res["winner"][b].cpu().numpy().reshape(-1).astype(np.int64), o["winner"])
        img = res["img"][b].cpu().numpy()
        assert np.array_equal(img[0:3], o["xyz"]) and np.array_equal(img[3:4], o["range"])
        assert np.array_equal(img[5:6].astype(np.int64), o["semantics"])
        row, col = o["pix"] // W, o["pix"] % W
        assert np.array_equal(back[sl], oproj.backproject_labels(label_img[b].cpu().numpy(), row, col))


def test_projection_properties_at_full_size(cuda):
    """Size-independent properties on an OS1-128 batch: every point lands in an occupied pixel, each
    winner projects to its own pixel and is the nearest point of that pixel, idempotent re-projection."""
    B = 4
    scans = [synth.synth_scan(100 + i, "os1-128") for i in range(B)]
    H, W = 128, 2048
    offs = np.concatenate([[0], np.cumsum([s[0].shape[0] for s in scans])])
    dx, dr, dl = to_dev(np.concatenate([s[0] for s in scans]), np.concatenate([s[1] for s in scans]), cuda)
    res = ops.project_batch(dx, dr, offs, H, W, lut=dl)
    pix, win = res["pix"].long(), res["winner"].reshape(B, -1).long()
    r2 = (dx[:, :3].double() ** 2).sum(1)
    for b in range(B):
        sl = slice(int(offs[b]), int(offs[b + 1]))
        p, w = pix[sl], win[b]
        assert int(p.min()) >= 0 and int(p.max()) < H * W
        occ = torch.nonzero(w >= 0).squeeze(1)
        assert torch.equal(torch.unique(p), occ)
        assert torch.equal(p[w[occ]], occ)                                   # winner lies in its pixel
        best = torch.full((H * W,), float("inf"), dtype=torch.float64, device=cuda)
        best.scatter_reduce_(0, p, r2[sl], reduce="amin")
        assert torch.equal(r2[sl][w[occ]], best[occ])                        # and is the nearest there
    again = ops.project_batch(dx, dr, offs, H, W, lut=dl)
    assert torch.equal(again["img"], res["img"]) and torch.equal(again["winner"], res["winner"])


def test_organized_cloud_identity(cuda):
    """Ouster-style organised cloud: pixel = n, back-projection is label_img.reshape(-1) (dataset.md:109)."""
    H, W = 8, 64
    label_img = torch.arange(H * W, device=cuda).reshape(1, H, W)
    pix = torch.arange(H * W, dtype=torch.int32, device=cuda)
    out = ops.backproject(label_img, pix, [0, H * W])
    assert torch.equal(out, label_img.reshape(-1))


def test_fast_and_exact_kernels_agree_bit_for_bit(cuda):
    """The batched entry point classifies points with fp32 angles and falls back to fp64 near bin edges;
    slu_debug_project_exact(1) forces fp64 for every point.  Both must give identical bits, including on
    points placed a few fp32/fp64 ulps around bin edges."""
    from semanticlidarunc_b200 import _lib
    H, W = 64, 2048
    xyzi, raw = synth.synth_scan(77, "hdl64")
    # adversarial points: directions sitting (almost) exactly on column edges and on the seam
    edges = np.linspace(-np.pi, np.pi, W)
    k = np.arange(0, 400)
    for j, d in enumerate((0.0, 1e-7, -1e-7, 1e-15, -1e-15, 5e-6, -5e-6, 1e-5)):
        ang = edges[(k * 5 + j) % W] + d
        sl = slice(j * 400, (j + 1) * 400)
        rr = np.linalg.norm(xyzi[sl, :2], axis=1)
        xyzi[sl, 0] = (rr * np.cos(ang)).astype(np.float32)
        xyzi[sl, 1] = (rr * np.sin(ang)).astype(np.float32)
    scans = [(xyzi, raw), synth.synth_scan(78, "hdl64", n_points=50_000)]
    offs = np.concatenate([[0], np.cumsum([s[0].shape[0] for s in scans])])
    dx, dr, dl = to_dev(np.concatenate([s[0] for s in scans]), np.concatenate([s[1] for s in scans]), cuda)
    fast = ops.project_batch(dx, dr, offs, H, W, lut=dl)
    prev = _lib.lib().slu_debug_project_exact(1)
    try:
        exact = ops.project_batch(dx, dr, offs, H, W, lut=dl)
    finally:
        _lib.lib().slu_debug_project_exact(prev)
    for k in ("pix", "winner", "img", "label", "theta"):
        assert torch.equal(fast[k], exact[k]), k
    assert torch.equal(fast["diag"], exact["diag"])
    for b, (a, r) in enumerate(scans):                  # and both equal the oracle
        o = oproj.kitti_frame(a, r, H, W, build_id_lut())
        assert np.array_equal(fast["pix"][offs[b]:offs[b + 1]].cpu().numpy().astype(np.int64), o["pix"])
        assert np.array_equal(fast["winner"][b].cpu().numpy().reshape(-1).astype(np.int64), o["winner"])


@pytest.mark.parametrize("theta_range,H", [((-np.pi / 8, np.pi / 8), 128), ((-np.pi / 2, np.pi / 2), 64)])
def test_fixed_range_fused_pass_agrees_with_exact_kernels(cuda, theta_range, H):
    """With a fixed elevation range (SemanticCUDAL +-pi/8 at 128 rows, SemanticWADS +-pi/2) the batched entry point runs
    ONE fused point pass (columns + rows + depth test).  It must give the bits of the all-fp64 kernels and of the oracle,
    including on points placed a few ulps around row and column edges and outside the range."""
    from semanticlidarunc_b200 import _lib
    W = 2048
    xyzi, raw = synth.synth_scan(91, "os1-128")
    xyzi, raw = xyzi[:150_000].copy(), raw[:150_000].copy()
    row_edges = np.linspace(theta_range[0], theta_range[1], H)
    k = np.arange(300)
    for j, d in enumerate((0.0, 1e-7, -1e-7, 1e-15, -1e-15, 5e-6, -5e-6, 1e-5)):     # elevations on / around row edges
        th = row_edges[(k * 3 + j) % H] + d
        sl = slice(j * 300, (j + 1) * 300)
        rho = np.linalg.norm(xyzi[sl, :2], axis=1).astype(np.float64)
        xyzi[sl, 2] = (rho * np.tan(th)).astype(np.float32)
    scans = [(xyzi, raw), synth.synth_scan(92, "hdl64", n_points=40_000)]
    offs = np.concatenate([[0], np.cumsum([s[0].shape[0] for s in scans])])
    dx, dr, dl = to_dev(np.concatenate([s[0] for s in scans]), np.concatenate([s[1] for s in scans]), cuda)
    fused = ops.project_batch(dx, dr, offs, H, W, lut=dl, theta_range=theta_range)
    prev = _lib.lib().slu_debug_project_exact(1)
    try:
        exact = ops.project_batch(dx, dr, offs, H, W, lut=dl, theta_range=theta_range)
    finally:
        _lib.lib().slu_debug_project_exact(prev)
    for key in ("pix", "winner", "img", "label", "theta", "diag"):
        assert torch.equal(fused[key], exact[key]), key
    for b, (a, r) in enumerate(scans):
        pc = np.concatenate([a.astype(np.float64), build_id_lut()[(r & 0xFFFF).astype(np.int64)].astype(np.float64)[:, None]], axis=1)
        row, col, _ = oproj.projection_indices(pc, H, W, theta_range=theta_range)
        assert np.array_equal(fused["pix"][offs[b]:offs[b + 1]].cpu().numpy().astype(np.int64), row * W + col)


def _project_in_mode(mode, *args, **kw):
    from semanticlidarunc_b200 import _lib
    prev = _lib.lib().slu_debug_project_exact(mode)
    try:
        return ops.project_batch(*args, **kw)
    finally:
        _lib.lib().slu_debug_project_exact(prev)


@pytest.mark.parametrize("case", ["auto", "fixed", "farthest", "yaw", "ties", "pile_up", "tiny_and_empty"])
def test_cell_pipeline_agrees_with_four_launch_and_exact_kernels(cuda, case):
    """slu_debug_project_exact(0) = the default (angles -> rows with a 64-bit atomicMin on the range -> tie pass -> resolve),
    (3) the same kernels with a 128-bit compare-and-swap depth test on a (range, index) cell per pixel and no tie pass, (2) the
    fused cell pipeline (extremes -> one point pass -> resolve), (1) the all-fp64 kernels.  All four must give identical bits: pixel of every point, winner per pixel INCLUDING the lowest-index rule among
    exactly equal ranges, image, labels, theta range and the diagnostics."""
    H, W = 64, 2048
    kw = {}
    scans = [synth.synth_scan(301, "hdl64"), synth.synth_scan(302, "hdl64", n_points=33_333)]
    if case == "fixed":
        kw["theta_range"] = (-0.45, 0.05)                        # clips a few beams: points outside the range wrap like numpy's
    elif case == "farthest":
        kw["farthest_wins"] = True
    elif case == "yaw":
        kw["yaw_deg"] = [17.0, -123.4]
    elif case == "ties":
        # exact duplicates (same float64 range, same pixel, different index) and same-range points on a sphere:
        # the lowest index must win whichever thread's compare-and-swap lands first
        a, r = scans[0]
        a = np.concatenate([a[:20_000], a[:20_000][::-1], a[5_000:15_000]])
        r = np.concatenate([r[:20_000], r[:20_000][::-1], r[5_000:15_000]])
        rng = np.random.default_rng(5)
        d = rng.normal(size=(30_000, 3))
        d[:, 2] *= 0.2
        d = (d / np.linalg.norm(d, axis=1, keepdims=True) * 10.0).astype(np.float32)      # ranges equal up to fp32 rounding
        sph = np.concatenate([d, np.zeros((30_000, 1), np.float32)], axis=1)
        scans = [(a, r), (sph, scans[1][1][:30_000])]
    elif case == "pile_up":
        # 60 000 points into a patch of ~200 pixels: long compare-and-swap retry chains
        rng = np.random.default_rng(6)
        az = rng.uniform(0.30, 0.35, 60_000)
        el = rng.uniform(-0.10, -0.08, 60_000)
        rg = rng.uniform(3.0, 50.0, 60_000)
        pts = np.stack([rg * np.cos(el) * np.cos(az), rg * np.cos(el) * np.sin(az), rg * np.sin(el), rg * 0], axis=1).astype(np.float32)
        scans = [(pts, scans[0][1][:60_000]), scans[1]]
    elif case == "tiny_and_empty":
        H, W = 16, 256
        e = (np.zeros((0, 4), np.float32), np.zeros((0,), np.uint32))
        scans = [synth.synth_scan(303, "tiny", n_points=1), e, synth.synth_scan(304, "tiny"), synth.synth_scan(305, "tiny", n_points=7)]
    offs = np.concatenate([[0], np.cumsum([s[0].shape[0] for s in scans])])
    dx, dr, dl = to_dev(np.concatenate([s[0] for s in scans]), np.concatenate([s[1] for s in scans]), cuda)
    cells = _project_in_mode(0, dx, dr, offs, H, W, lut=dl, **kw)
    fused = _project_in_mode(2, dx, dr, offs, H, W, lut=dl, **kw)
    ties = _project_in_mode(3, dx, dr, offs, H, W, lut=dl, **kw)
    exact = _project_in_mode(1, dx, dr, offs, H, W, lut=dl, **kw)
    for key in ("pix", "winner", "img", "label", "theta", "diag"):
        assert torch.equal(cells[key], fused[key]), ("fused cell pipeline", key)
        assert torch.equal(cells[key], ties[key]), ("cells in the row kernel", key)
        assert torch.equal(cells[key], exact[key]), ("exact", key)
    if case == "ties":
        w = cells["winner"][0].reshape(-1).cpu().numpy()
        assert w.max() < 20_000                                     # every duplicate lost to its lower-index twin
    if case in ("auto", "ties", "pile_up"):
        for b, (a, r) in enumerate(scans):
            o = oproj.kitti_frame(a, r, H, W, build_id_lut())
            assert np.array_equal(cells["pix"][offs[b]:offs[b + 1]].cpu().numpy().astype(np.int64), o["pix"])
            if case != "ties":                                      # the oracle's winner among exact ties follows argsort
                assert np.array_equal(cells["winner"][b].cpu().numpy().reshape(-1).astype(np.int64), o["winner"])
