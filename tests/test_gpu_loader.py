"""The device loader (projection + slu_frame_tensors) against the reference Dataset's golden outputs:
copies (xyz, range, reflectivity, semantics) bit-exact; normals within 5e-5 absolute on every pixel where
the normal is well conditioned (tests/helpers.py::normals_condition_mask -- at the corners of
nearest-neighbour upsampled blocks the two gradient vectors are parallel and the reference's own normal
is float32 rounding noise), finite and of norm <= 1 everywhere."""
import hashlib
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import projection as oproj
from semanticlidarunc_b200 import synth
from semanticlidarunc_b200.dataset.dataloader_semantic_KITTI import SemanticKitti
from semanticlidarunc_b200.dataset.definitions import build_id_lut
from tests.helpers import normals_condition_mask

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def write_pair(d, xyzi, raw):
    fb, fl = os.path.join(d, "000000.bin"), os.path.join(d, "000000.label")
    xyzi.tofile(fb)
    raw.tofile(fl)
    return [(fb, fl)]


def close(a, b, rtol=1e-5, atol=1e-6):
    return np.abs(a.astype(np.float64) - b) .max() <= atol + rtol * np.abs(b).max()


def normals_close(got, ref, xyz):
    mask = normals_condition_mask(xyz)
    assert mask.mean() > 0.95
    assert np.isfinite(got).all() and (np.linalg.norm(got, axis=0) <= 1.0 + 1e-5).all()
    return np.abs(got.astype(np.float64) - ref)[:, mask].max() <= 5e-5


def test_getitem_native_resolution_vs_reference(cuda, golden):
    g = golden("kitti_loader.npz")
    with tempfile.TemporaryDirectory() as d:
        ds = SemanticKitti(write_pair(d, g["xyzi"], g["raw"]), projection=(16, 256), resize=False)
        rng, refl, xyz, normals, sem = ds[0]
    assert rng.device.type == "cpu" and sem.dtype == torch.int64 and tuple(normals.shape) == (3, 16, 256)
    for t, k in ((rng, "range"), (refl, "reflectivity"), (xyz, "xyz"), (sem, "semantics")):
        assert np.array_equal(t.numpy(), g[k]), k
    assert normals_close(normals.numpy(), g["normals"], g["xyz"])


def test_getitem_resize_vs_reference(cuda, golden):
    g = golden("kitti_loader_aug.npz")
    with tempfile.TemporaryDirectory() as d:
        ds = SemanticKitti(write_pair(d, g["xyzi"], g["raw"]), projection=(16, 256), resize=True)
        rng, refl, xyz, normals, sem = (t.numpy() for t in ds[0])
    assert xyz.shape == (3, 128, 2048)
    for a, k in ((rng, "range"), (refl, "reflectivity"), (xyz, "xyz"), (sem, "semantics")):
        assert sha(a) == bytes(g["resize/" + k + "_sha"]).hex(), k           # cv2 INTER_NEAREST index rule, bit-exact
    m = normals_condition_mask(xyz)[::4, ::16]
    assert np.abs(normals[:, ::4, ::16].astype(np.float64) - g["resize/normals_sub"])[:, m].max() <= 5e-5
    assert np.isfinite(normals).all()


def test_getitem_rotate_flip_vs_reference(cuda, golden):
    g = golden("kitti_loader_aug.npz")
    angle = float(g["aug/angle"])
    with tempfile.TemporaryDirectory() as d:
        ds = SemanticKitti(write_pair(d, g["xyzi"], g["raw"]), projection=(16, 256), resize=False)
        out = ds.device_batch([(g["xyzi"], g["raw"])], yaw_deg=[angle], flip=[True])
    sem = out["semantics"][0].cpu().numpy()
    xyz = out["xyz"][0].cpu().numpy()
    # np.dot's float64 accumulation order is BLAS-specific: coordinates agree to float32 rounding, pixels
    # may differ only where a rotated point sits within an ulp of a bin edge
    same = (sem == g["aug/semantics"]).mean()
    assert same > 0.999, same
    ok = sem == g["aug/semantics"]
    assert np.abs(xyz - g["aug/xyz"])[np.broadcast_to(ok, xyz.shape)].max() <= 1e-5
    assert close(out["range"][0].cpu().numpy(), g["aug/range"], atol=1e-5)


def test_getitem_draws_augmentation_like_reference(cuda, golden):
    """Same global-RNG protocol as the reference (:53 randint, :71 rand)."""
    g = golden("kitti_loader_aug.npz")
    with tempfile.TemporaryDirectory() as d:
        ds = SemanticKitti(write_pair(d, g["xyzi"], g["raw"]), rotate=True, flip=True, projection=(16, 256), resize=False)
        seed = 0
        while True:
            np.random.seed(seed)
            angle = float(np.random.randint(-180, 180))
            if np.random.rand() < 0.5:
                break
            seed += 1
        assert angle == float(g["aug/angle"])
        np.random.seed(seed)
        rng, refl, xyz, normals, sem = ds[0]
    assert (sem.numpy() == g["aug/semantics"]).mean() > 0.999


def test_batch_of_scans_vs_oracle_items(cuda):
    lut = build_id_lut()
    scans = [synth.synth_scan(40 + i, "tiny", n_points=None if i else 2000) for i in range(3)]
    ds = SemanticKitti([], projection=(16, 256), resize=True)
    out = ds.device_batch(scans, flip=[False, True, False])
    for b, (a, r) in enumerate(scans):
        ref = oproj.kitti_item(a, r, lut, projection=(16, 256), resize=True, flip=(b == 1))
        for k, t in zip(("range", "reflectivity", "xyz", "normals", "semantics"), ref):
            got = out[k][b].cpu().numpy()
            if k == "normals":
                assert normals_close(got, t, ref[2]), (b, k)
            else:
                assert np.array_equal(got, t), (b, k)


def test_unknown_label_id_raises_like_reference(cuda):
    xyzi, raw = synth.synth_scan(3, "tiny", n_points=100)
    raw = raw.copy()
    raw[5] = 7            # 7 is not a SemanticKITTI id: the reference's id_map[l] raises KeyError
    with tempfile.TemporaryDirectory() as d:
        ds = SemanticKitti(write_pair(d, xyzi, raw), projection=(16, 256), resize=False)
        with pytest.raises(KeyError):
            ds[0]


def test_cudal_getitem_vs_reference(cuda, golden):
    from semanticlidarunc_b200.dataset.dataloader_semantic_CUDAL import SemanticCUDAL
    g = golden("other_loaders.npz")
    with tempfile.TemporaryDirectory() as d:
        ds = SemanticCUDAL(write_pair(d, g["cudal/xyzi"], g["cudal/raw"]), projection=(32, 256), resize=True)
        rng, refl, xyz, normals, sem = (t.numpy() for t in ds[0])
    for a, k in ((rng, "range"), (refl, "reflectivity"), (xyz, "xyz"), (sem, "semantics")):
        assert sha(a) == bytes(g["cudal/" + k + "_sha"]).hex(), k
    assert refl.max() <= 1.0 and 12 in np.unique(sem)
    m = normals_condition_mask(xyz)[::4, ::16]
    assert np.abs(normals[:, ::4, ::16].astype(np.float64) - g["cudal/normals_sub"])[:, m].max() <= 5e-5


def test_thab_organised_cloud_vs_reference(cuda, golden):
    from semanticlidarunc_b200.dataset.dataloader_semantic_THAB import SemanticTHAB
    g = golden("other_loaders.npz")
    xyzi, raw = synth.synth_scan(int(g["thab/seed"]), "os1-128")
    with tempfile.TemporaryDirectory() as d:
        paths = write_pair(d, xyzi, raw)
        rng, refl, xyz, normals, sem = (t.numpy() for t in SemanticTHAB(paths)[0])
        for a, k in ((rng, "range"), (refl, "reflectivity"), (xyz, "xyz"), (sem, "semantics")):
            assert sha(a) == bytes(g["thab_plain/" + k + "_sha"]).hex(), k
        m = normals_condition_mask(xyz)[::4, ::16]
        assert m.mean() > 0.9
        assert np.abs(normals[:, ::4, ::16].astype(np.float64) - g["thab_plain/normals_sub"])[:, m].max() <= 5e-5
        # flip + yaw with the reference's RNG protocol (coin first, then the angle)
        np.random.seed(int(g["thab_aug/np_seed"]))
        rng, refl, xyz, normals, sem = (t.numpy() for t in SemanticTHAB(paths, rotate=True, flip=True)[0])
    assert sha(sem) == bytes(g["thab_aug/semantics_sha"]).hex() and sha(refl) == bytes(g["thab_aug/reflectivity_sha"]).hex()
    # rotated coordinates: np.dot's float64 accumulation is BLAS-specific, so float32(xyz) may differ in the last bit
    assert np.abs(xyz[:, ::4, ::16] - g["thab_aug/xyz_sub"]).max() <= 1e-5
    assert np.abs(rng[:, ::4, ::16] - g["thab_aug/range_sub"]).max() <= 1e-5


def test_wads_getitem_drops_empty_rows_like_reference(cuda, golden):
    from semanticlidarunc_b200.dataset.dataloader_semantic_WADS import SemanticWADS
    g = golden("other_loaders.npz")
    with tempfile.TemporaryDirectory() as d:
        paths = write_pair(d, g["wads/xyzi"], g["wads/raw"])
        rng, refl, xyz, normals, sem = (t.numpy() for t in SemanticWADS(paths, projection=(64, 256), resize=True)[0])
        assert xyz.shape == (3, 64, 1024) and 20 in np.unique(sem)
        for a, k in ((rng, "range"), (refl, "reflectivity"), (xyz, "xyz"), (sem, "semantics")):
            assert sha(a) == bytes(g["wads/" + k + "_sha"]).hex(), k
        m = normals_condition_mask(xyz)[::4, ::16]
        assert np.abs(normals[:, ::4, ::16].astype(np.float64) - g["wads/normals_sub"])[:, m].max() <= 5e-5
        # without the resize the image keeps only the rows that received points (variable height)
        rng, refl, xyz, normals, sem = (t.numpy() for t in SemanticWADS(paths, projection=(64, 256), resize=False)[0])
    assert xyz.shape == tuple(g["wads_native/shape"])
    for a, k in ((rng, "range"), (refl, "reflectivity"), (xyz, "xyz"), (sem, "semantics")):
        assert np.array_equal(a, g["wads_native/" + k]), k
    m = normals_condition_mask(xyz)
    assert np.abs(normals.astype(np.float64) - g["wads_native/normals"])[:, m].max() <= 5e-5


def test_stf_getitem_vs_reference(cuda, golden):
    from semanticlidarunc_b200.dataset.dataloader_semantic_STF import SemanticSTF
    g = golden("other_loaders.npz")
    with tempfile.TemporaryDirectory() as d:
        fb, fl = os.path.join(d, "0.bin"), os.path.join(d, "0.label")
        g["stf/five"].tofile(fb)
        g["stf/label"].tofile(fl)
        ds = SemanticSTF([(fb, fl)], projection=(16, 256), resize=True, remap_adverse_label=True, clip=True)
        rng, refl, xyz, normals, sem = (t.numpy() for t in ds[0])
    for a, k in ((rng, "range"), (refl, "reflectivity"), (xyz, "xyz"), (sem, "semantics")):
        assert sha(a) == bytes(g["stf/" + k + "_sha"]).hex(), k
    assert 20 not in np.unique(sem) and 21 in np.unique(sem)
    m = normals_condition_mask(xyz)[::4, ::16]
    assert np.abs(normals[:, ::4, ::16].astype(np.float64) - g["stf/normals_sub"])[:, m].max() <= 5e-5


def test_build_normal_xyz_standalone_vs_reference_golden(cuda, golden):
    """dataset.utils.build_normal_xyz ([h,w,3] in, [h,w,3] out, src/dataset/utils.py:30-59) against the normals the
    reference Dataset produced from the same xyz image."""
    from semanticlidarunc_b200.dataset.utils import build_normal_xyz
    g = golden("kitti_loader.npz")
    xyz = g["xyz"]                                             # [3,H,W] as the Dataset returns it
    n = build_normal_xyz(np.ascontiguousarray(xyz.transpose(1, 2, 0)))
    assert isinstance(n, np.ndarray) and n.shape == xyz.shape[1:] + (3,) and n.dtype == np.float32
    assert normals_close(n.transpose(2, 0, 1), g["normals"], xyz)
    nd = build_normal_xyz(torch.from_numpy(np.ascontiguousarray(xyz.transpose(1, 2, 0))).to(cuda))
    assert nd.is_cuda and np.array_equal(nd.cpu().numpy(), n)
