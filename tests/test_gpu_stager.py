"""Native scan stager (csrc/slu_stager.cu): files -> pinned slots -> device, byte-exact against np.fromfile, and the
SemanticKitti.staged_batches path against the same dataset fed from numpy arrays."""
import numpy as np
import pytest
import torch

from semanticlidarunc_b200 import synth
from semanticlidarunc_b200.dataset.dataloader_semantic_KITTI import SemanticKitti
from semanticlidarunc_b200.dataset.stager import ScanStager

pytestmark = pytest.mark.gpu


def write_scans(tmp_path, n, sensor="tiny", n_points=None):
    paths, arrays = [], []
    for i in range(n):
        xyzi, raw = synth.synth_scan(40 + i, sensor, n_points=None if n_points is None else n_points + 17 * i)
        b, l = tmp_path / f"{i:06d}.bin", tmp_path / f"{i:06d}.label"
        xyzi.tofile(b); raw.tofile(l)
        paths.append((str(b), str(l))); arrays.append((xyzi, raw))
    return paths, arrays


def test_round_trip_more_tickets_than_slots(cuda, tmp_path):
    paths, arrays = write_scans(tmp_path, 11, n_points=3000)
    with ScanStager(n_slots=2, max_points=4000, n_io_threads=3, device=cuda) as st:
        tickets = [st.submit(b, l) for b, l in paths]
        assert tickets == list(range(11))
        for t, (xyzi, raw) in zip(tickets, arrays):
            dx, dr = st.fetch(t)
            assert dx.shape == (xyzi.shape[0], 4) and dr.shape == (raw.shape[0],)
            assert np.array_equal(dx.cpu().numpy().view(np.uint32), np.fromfile(paths[t][0], dtype=np.float32).reshape(-1, 4).view(np.uint32))
            assert np.array_equal(dr.cpu().numpy().view(np.uint32), np.fromfile(paths[t][1], dtype=np.uint32))
        # no label file: points only
        t = st.submit(paths[0][0], None)
        dx, dr = st.fetch(t)
        assert dr is None and dx.shape[0] == arrays[0][0].shape[0]


def test_errors_do_not_poison_the_queue(cuda, tmp_path):
    paths, arrays = write_scans(tmp_path, 2, n_points=500)
    bad_size = tmp_path / "bad.bin"
    bad_size.write_bytes(b"\0" * 30)                                   # not a multiple of 16
    short_label = tmp_path / "short.label"
    arrays[0][1][:-3].tofile(short_label)
    with ScanStager(n_slots=2, max_points=600, device=cuda) as st:
        t_missing = st.submit(str(tmp_path / "nope.bin"), None)
        t_size = st.submit(str(bad_size), None)
        t_label = st.submit(paths[0][0], str(short_label))
        t_ok = st.submit(*paths[1])
        for t in (t_missing, t_size, t_label):
            with pytest.raises(OSError):
                st.fetch(t)
        dx, dr = st.fetch(t_ok)
        assert np.array_equal(dx.cpu().numpy(), arrays[1][0])
        with pytest.raises(ValueError):
            st.fetch(99)                                               # never issued
    with ScanStager(n_slots=1, max_points=100, device=cuda) as st:    # scan larger than the slots
        with pytest.raises(OSError):
            st.fetch(st.submit(*paths[0]))


@pytest.mark.parametrize("batch_size", [1, 4])
def test_dataset_staged_batches_match_numpy_fed_batches(cuda, tmp_path, batch_size):
    paths, arrays = write_scans(tmp_path, 6, n_points=2500)
    ds = SemanticKitti(paths, projection=(16, 256), resize=True, device=cuda, return_device=True)
    got = list(ds.staged_batches(range(6), batch_size=batch_size, max_points=4000))
    assert len(got) == (6 + batch_size - 1) // batch_size
    k = 0
    for out in got:
        nb = out["range"].shape[0]
        ref = ds.device_batch(arrays[k:k + nb])
        for name in ("range", "reflectivity", "xyz", "normals", "semantics", "pix"):
            assert torch.equal(out[name], ref[name]), name
        assert np.array_equal(out["offsets"], ref["offsets"])
        k += nb
    assert k == 6
