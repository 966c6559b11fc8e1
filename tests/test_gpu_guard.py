"""compute-sanitizer is closed on this pool, so out-of-bounds WRITES are checked by hand: every output
of every entry point is a window inside a larger sentinel-filled buffer, called through the raw C ABI on
ragged sizes (partial tiles, HW % 4 != 0, B = 1 / 3); the sentinels around the windows must survive."""
import numpy as np
import pytest
import torch

from semanticlidarunc_b200 import _lib, ops, synth
from semanticlidarunc_b200.dataset.definitions import build_id_lut

pytestmark = pytest.mark.gpu
PAD = 1024


class Guarded:
    def __init__(self, shape, dtype, dev):
        n = int(np.prod(shape))
        self.buf = torch.empty(n + 2 * PAD, dtype=dtype, device=dev)
        self.sent = 123 if dtype in (torch.int64, torch.int32) else 777.0
        self.buf.fill_(self.sent)
        self.win = self.buf[PAD:PAD + n].view(shape)

    def intact(self):
        return bool((self.buf[:PAD] == self.sent).all()) and bool((self.buf[-PAD:] == self.sent).all())


@pytest.mark.parametrize("shape", [(3, 1, 20, 5, 200), (2, 3, 7, 3, 37), (1, 2, 20, 1, 260), (20, 1, 13, 2, 514)])
@pytest.mark.parametrize("direct", [False, True])
def test_reduce_outputs_stay_in_bounds(cuda, shape, direct):
    T, B, C, H, W = shape
    x, lab = synth.synth_mc_logits(1, T, B, C, H, W)
    x, lab = x.to(cuda), lab.to(cuda)
    g = {"pbar": Guarded((B, C, H, W), torch.float32, cuda), "pred": Guarded((B, H, W), torch.int64, cuda),
         "conf": Guarded((B, H, W), torch.float32, cuda), "h": Guarded((B, H, W), torch.float32, cuda),
         "mi": Guarded((B, H, W), torch.float32, cuda), "cm": Guarded((C, C), torch.int64, cuda),
         "bins": Guarded((3, 15), torch.int64, cuda)}
    g["cm"].win.zero_(); g["bins"].win.zero_()
    fn = _lib.lib().slu_reduce_metrics_direct if direct else _lib.lib().slu_reduce_metrics
    rc = fn(_lib.ptr(x), _lib.ptr(lab), T, B, C, H * W, 0, 1, 1e-12, 1, 1, 0, 15, _lib.edges_array(ops.uniform_edges(15)),
            _lib.ptr(g["pbar"].win), _lib.ptr(g["pred"].win), _lib.ptr(g["conf"].win), _lib.ptr(g["h"].win),
            _lib.ptr(g["mi"].win), _lib.ptr(g["cm"].win), _lib.ptr(g["bins"].win), _lib.stream_ptr())
    _lib.check(rc, "slu_reduce_metrics")
    torch.cuda.synchronize()
    for k, v in g.items():
        assert v.intact(), k
    assert int(g["cm"].win.sum()) == B * H * W
    assert bool(torch.isfinite(g["h"].win).all()) and not bool((g["h"].win == 777.0).any())


@pytest.mark.parametrize("n_points", [1, 255, 257, 3200])
def test_projection_outputs_stay_in_bounds(cuda, n_points):
    H, W = 16, 256
    scans = [synth.synth_scan(5, "tiny", n_points=n_points), synth.synth_scan(6, "tiny", n_points=max(1, n_points // 3))]
    offs = np.concatenate([[0], np.cumsum([s[0].shape[0] for s in scans])]).astype(np.int64)
    xyzi = torch.from_numpy(np.concatenate([s[0] for s in scans])).to(cuda)
    raw = torch.from_numpy(np.concatenate([s[1] for s in scans]).view(np.int32)).to(cuda)
    lut = torch.from_numpy(build_id_lut()).to(cuda)
    n, B = int(offs[-1]), 2
    need = _lib.lib().slu_project_workspace_bytes(n, B, H * W)
    g = {"work": Guarded((need,), torch.uint8, cuda) if False else None}
    work = torch.empty(need + 2 * PAD, dtype=torch.uint8, device=cuda)
    work.fill_(0x5A)
    # the workspace window must start 16-byte aligned
    off = (-work.data_ptr()) % 16 + 16 * (PAD // 16)
    wwin = work[off:off + need]
    outs = {"img": Guarded((B, 6, H, W), torch.float32, cuda), "label": Guarded((B, H, W), torch.int64, cuda),
            "pix": Guarded((n,), torch.int32, cuda), "winner": Guarded((B, H, W), torch.int32, cuda),
            "theta": Guarded((B, 2), torch.float64, cuda), "diag": Guarded((B, 2), torch.int32, cuda)}
    rc = _lib.lib().slu_project_batch(_lib.ptr(xyzi), _lib.ptr(raw), _lib.ptr(lut), offs.ctypes.data_as(_lib.C.c_void_p), n, B, H, W,
                                      0, 0.0, 0.0, 0, None, _lib.ptr(wwin), _lib.ptr(outs["img"].win), _lib.ptr(outs["label"].win),
                                      _lib.ptr(outs["pix"].win), _lib.ptr(outs["winner"].win), _lib.ptr(outs["theta"].win),
                                      _lib.ptr(outs["diag"].win), _lib.stream_ptr())
    _lib.check(rc, "slu_project_batch")
    back = Guarded((n,), torch.int64, cuda)
    rc = _lib.lib().slu_backproject(_lib.ptr(outs["label"].win), _lib.ptr(outs["pix"].win), offs.ctypes.data_as(_lib.C.c_void_p),
                                    n, B, H * W, _lib.ptr(back.win), _lib.stream_ptr())
    _lib.check(rc, "slu_backproject")
    torch.cuda.synchronize()
    for k, v in outs.items():
        assert v.intact(), k
    assert back.intact()
    assert bool((work[:off] == 0x5A).all()) and bool((work[off + need:] == 0x5A).all())
    assert int(outs["pix"].win.min()) >= 0 and int(outs["pix"].win.max()) < H * W


def test_frame_loss_evidential_hist_outputs_stay_in_bounds(cuda):
    B, C, H, W = 3, 13, 5, 77
    img = torch.randn((B, 6, H, W), device=cuda)
    o = {k: Guarded(s, d, cuda) for k, s, d in (("range", (B, 1, 9, 101), torch.float32), ("refl", (B, 1, 9, 101), torch.float32),
                                                ("xyz", (B, 3, 9, 101), torch.float32), ("nrm", (B, 3, 9, 101), torch.float32),
                                                ("sem", (B, 1, 9, 101), torch.int64))}
    flip = np.array([1, 0, 1], dtype=np.uint8)
    rc = _lib.lib().slu_frame_tensors(_lib.ptr(img), B, H, W, 9, 101, flip.ctypes.data_as(_lib.C.c_void_p), 0.25,
                                      _lib.ptr(o["range"].win), _lib.ptr(o["refl"].win), _lib.ptr(o["xyz"].win),
                                      _lib.ptr(o["nrm"].win), _lib.ptr(o["sem"].win), None, 1, _lib.stream_ptr())
    _lib.check(rc, "slu_frame_tensors")
    alpha = torch.rand((B, C, H, W), device=cuda) * 5 + 1
    tgt = torch.randint(0, C, (B, H, W), device=cuda)
    gm, gk = Guarded((B, C, H, W), torch.float32, cuda), Guarded((B, C, H, W), torch.float32, cuda)
    sums = Guarded((3,), torch.float64, cuda)
    sums.win.zero_()
    ign = (_lib.C.c_int64 * 1)(0)
    rc = _lib.lib().slu_dirichlet_loss(_lib.ptr(alpha), _lib.ptr(tgt), None, B, C, H * W, ign, 1, 1e-8, 1e-8, 1, 1,
                                       _lib.ptr(sums.win), _lib.ptr(gm.win), _lib.ptr(gk.win), _lib.stream_ptr())
    _lib.check(rc, "slu_dirichlet_loss")
    ev = {k: Guarded((B, H, W), torch.float32, cuda) for k in ("conf", "h", "au", "eu", "mi")}
    ev["alpha"] = Guarded((B, C, H, W), torch.float32, cuda)
    ev["pred"] = Guarded((B, H, W), torch.int64, cuda)
    outs = torch.randn((B, C + 1, H, W), device=cuda)
    rc = _lib.lib().slu_evidential_reduce(_lib.ptr(outs), None, _lib.ptr(tgt), B, C, H * W, 1.0, 1e-8, 1e-12, 1, 1, 0, 0, None,
                                          _lib.ptr(ev["alpha"].win), _lib.ptr(ev["pred"].win), _lib.ptr(ev["conf"].win),
                                          _lib.ptr(ev["h"].win), _lib.ptr(ev["au"].win), _lib.ptr(ev["eu"].win), _lib.ptr(ev["mi"].win),
                                          None, None, _lib.stream_ptr())
    _lib.check(rc, "slu_evidential_reduce")
    cm, bins = Guarded((C, C), torch.int64, cuda), Guarded((3, 15), torch.int64, cuda)
    cm.win.zero_(); bins.win.zero_()
    pred = torch.randint(-2, C + 2, (B * H * W + 3,), device=cuda)
    lab = torch.randint(-2, C + 2, (B * H * W + 3,), device=cuda)
    conf = torch.rand((B * H * W + 3,), device=cuda)
    rc = _lib.lib().slu_confusion_ece(_lib.ptr(pred), _lib.ptr(lab), _lib.ptr(conf), pred.numel(), C, 1, 0, 15,
                                      _lib.edges_array(ops.uniform_edges(15)), _lib.ptr(cm.win), _lib.ptr(bins.win), _lib.stream_ptr())
    _lib.check(rc, "slu_confusion_ece")
    torch.cuda.synchronize()
    for d in (o, ev):
        for k, v in d.items():
            assert v.intact(), k
    assert gm.intact() and gk.intact() and sums.intact() and cm.intact() and bins.intact()


def test_loss_term_outputs_stay_in_bounds(cuda):
    """slu_evidence_term (all five terms), slu_dirichlet_term, slu_logit_regularizer on a ragged shape."""
    B, C, H, W = 3, 13, 5, 77
    alpha = torch.rand((B, C, H, W), device=cuda) * 5 + 1
    tgt = torch.randint(0, C, (B, H, W), device=cuda)
    ign = (_lib.C.c_int64 * 1)(0)
    prms = {ops.TERM_COMP_KL: (2.0, 0.55, 0.12, -1.0, 1.0, 1e-8, 1.0), ops.TERM_WRONG_LOW: (0.0, 0.05, 0.08, 1e-8),
            ops.TERM_EVID_BAND: (30.0, 0.1), ops.TERM_EVID_REG: (30.0, 0.0, 0.1, 0.0), ops.TERM_KL_CONF: (1.0, 1e-8)}
    guards = []
    for term, prm in prms.items():
        g, s = Guarded((B, C, H, W), torch.float32, cuda), Guarded((2,), torch.float64, cuda)
        s.win.zero_()
        arr = (_lib.C.c_float * len(prm))(*prm)
        rc = _lib.lib().slu_evidence_term(_lib.ptr(alpha), _lib.ptr(tgt), None, B, C, H * W, ign, 1, term, arr, len(prm),
                                          _lib.ptr(s.win), _lib.ptr(g.win), _lib.stream_ptr())
        _lib.check(rc, "slu_evidence_term")
        guards += [g, s]
    for term in (ops.TERM_NLL, ops.TERM_DIGAMMA_CE, ops.TERM_BRIER):
        g, s = Guarded((B, C, H, W), torch.float32, cuda), Guarded((2,), torch.float64, cuda)
        s.win.zero_()
        rc = _lib.lib().slu_dirichlet_term(_lib.ptr(alpha), _lib.ptr(tgt), None, B, C, H * W, ign, 1, term, 1e-8, -1.0,
                                           _lib.ptr(s.win), _lib.ptr(g.win), _lib.stream_ptr())
        _lib.check(rc, "slu_dirichlet_term")
        guards += [g, s]
    z = torch.randn((B, C + 1, H, W), device=cuda)
    g, s = Guarded((B, C + 1, H, W), torch.float32, cuda), Guarded((2,), torch.float64, cuda)
    s.win.zero_()
    rc = _lib.lib().slu_logit_regularizer(_lib.ptr(z), _lib.ptr(tgt), None, B, C + 1, H * W, ign, 1, 1, 0.5,
                                          _lib.ptr(s.win), _lib.ptr(g.win), _lib.stream_ptr())
    _lib.check(rc, "slu_logit_regularizer")
    guards += [g, s]
    torch.cuda.synchronize()
    assert all(x.intact() for x in guards)
    assert not any(bool((x.win == 777.0).any()) for x in guards)        # every output element was written
    # wrong parameter count / missing target are refused before any launch
    arr = (_lib.C.c_float * 2)(1.0, 1e-8)
    assert _lib.lib().slu_evidence_term(_lib.ptr(alpha), _lib.ptr(tgt), None, B, C, H * W, ign, 1, ops.TERM_COMP_KL, arr, 2,
                                        _lib.ptr(guards[1].win), None, _lib.stream_ptr()) == -1
    assert _lib.lib().slu_evidence_term(_lib.ptr(alpha), None, None, B, C, H * W, None, 0, ops.TERM_KL_CONF, arr, 2,
                                        _lib.ptr(guards[1].win), None, _lib.stream_ptr()) == -1
