"""Round-2 additions: launch counter, stand-alone normals, int32 histogram entry point, ScanEvaluator (all outputs on the
host path, idempotent summary, graph capture), AUROC score overrides outside [0,1], long ignore lists."""
import numpy as np
import pytest
import torch

from oracle import metrics as om
from oracle import projection as oproj
from oracle import uncertainty as ou
from semanticlidarunc_b200 import _lib, ops, synth
from semanticlidarunc_b200.dataset.definitions import build_id_lut
from semanticlidarunc_b200.pipeline import ScanEvaluator

pytestmark = pytest.mark.gpu


def test_launch_counter_counts_kernels(cuda):
    x, lab = synth.synth_mc_logits(1, 3, 1, 20, 8, 128)
    n0 = _lib.launch_count()
    ops.reduce_metrics(x.to(cuda), lab.to(cuda), kind="logits")
    assert _lib.launch_count() - n0 == 1
    n0 = _lib.launch_count()
    ops.confusion_ece(torch.zeros(10, dtype=torch.int64, device=cuda), torch.zeros(10, dtype=torch.int64, device=cuda), None,
                      num_classes=3, confmat=ops.new_confmat(3, cuda))
    assert _lib.launch_count() - n0 == 1


def test_frame_normals_equals_loader_normals(cuda):
    """slu_frame_normals on the projection image in place == the normals slu_frame_tensors derives from its xyz copy,
    and == the oracle's build_normal_xyz where the normal is well conditioned."""
    from tests.helpers import normals_condition_mask
    scans = [synth.synth_scan(i, "tiny") for i in range(3)]
    offs = np.concatenate([[0], np.cumsum([s[0].shape[0] for s in scans])]).astype(np.int64)
    xyzi = torch.from_numpy(np.concatenate([s[0] for s in scans])).to(cuda)
    raw = torch.from_numpy(np.concatenate([s[1] for s in scans]).view(np.int32)).to(cuda)
    proj = ops.project_batch(xyzi, raw, offs, 16, 256, lut=torch.from_numpy(build_id_lut()).to(cuda))
    a = ops.frame_normals(proj["img"])
    b = ops.frame_tensors(proj["img"])["normals"]
    assert torch.equal(a, b)
    c = ops.frame_normals(proj["img"][:, :3].contiguous())
    assert torch.equal(a, c)
    xyz = proj["img"][0, :3].cpu().numpy()
    ref = oproj.build_normal_xyz(np.ascontiguousarray(xyz.transpose(1, 2, 0))).transpose(2, 0, 1)
    m = normals_condition_mask(xyz)
    assert np.abs(a[0].cpu().numpy() - ref)[:, m].max() < 5e-5


@pytest.mark.parametrize("n", [17, 4096, 300001])
def test_int32_histogram_entry_point_equals_int64(cuda, n):
    g = torch.Generator().manual_seed(n)
    C = 20
    pred = torch.randint(-1, C + 1, (n,), generator=g)
    lab = torch.randint(-1, C + 1, (n,), generator=g)
    conf = torch.rand(n, generator=g)
    conf[::97] = float("nan")
    cm64, b64 = ops.new_confmat(C, cuda), ops.new_ece_bins(15, cuda)
    cm32, b32 = ops.new_confmat(C, cuda), ops.new_ece_bins(15, cuda)
    ops.confusion_ece(pred.to(cuda), lab.to(cuda), conf.to(cuda), num_classes=C, ignore_index=0, confmat=cm64, ece_bins=b64)
    ops.confusion_ece(pred.int().to(cuda), lab.int().to(cuda), conf.to(cuda), num_classes=C, ignore_index=0, confmat=cm32, ece_bins=b32)
    assert torch.equal(cm64, cm32) and torch.equal(b64, b32)
    assert torch.equal(cm64.cpu(), om.confusion_counts(pred, lab, C))


def _batch(cuda, n_scans, T=3, H=16, W=256, C=20):
    scans = [synth.synth_scan(10 + i, "tiny") for i in range(n_scans)]
    offs = np.concatenate([[0], np.cumsum([s[0].shape[0] for s in scans])]).astype(np.int64)
    xyzi = torch.from_numpy(np.concatenate([s[0] for s in scans])).to(cuda)
    raw = torch.from_numpy(np.concatenate([s[1] for s in scans]).view(np.int32)).to(cuda)
    logits, _ = synth.synth_mc_logits(7, T, n_scans, C, H, W)
    return scans, offs, xyzi, raw, logits


def test_scan_evaluator_host_path_returns_every_map_and_matches_device_path(cuda):
    H, W, C, T = 16, 256, 20, 3
    scans, offs, xyzi, raw, logits = _batch(cuda, 4, T, H, W, C)
    ev = ScanEvaluator(H, W, C, device=cuda)
    dev_out = ev.step_device(xyzi, raw, offs, logits.to(cuda))
    assert dev_out["normals"].shape == (4, 3, H, W)
    cm_dev = ev.confmat.clone()
    ev.reset()
    host = [(torch.from_numpy(s[0]).pin_memory(), torch.from_numpy(s[1].view(np.int32)).pin_memory(),
             logits[:, i:i + 1].contiguous().pin_memory()) for i, s in enumerate(scans)]
    outs = ev.step_host(host)
    assert torch.equal(ev.confmat, cm_dev)
    for i, o in enumerate(outs):
        assert set(o) == {"point_labels", "pred", "conf", "H_norm", "MI_norm"}
        assert not o["pred"].is_cuda and o["point_labels"].numel() == scans[i][0].shape[0]
        assert torch.equal(o["pred"], dev_out["pred"][i].cpu())
        assert torch.equal(o["H_norm"], dev_out["H_norm"][i].cpu()) and torch.equal(o["MI_norm"], dev_out["MI_norm"][i].cpu())
        assert torch.equal(o["conf"], dev_out["conf"][i].cpu())
        assert torch.equal(o["point_labels"], dev_out["point_labels"][offs[i]:offs[i + 1]].cpu())
        ref = ou.mc_reduce(logits[:, i:i + 1])
        assert (o["H_norm"] - ref["H_norm"][0]).abs().max() < 1e-5
    # pinned result buffers are per batch position, not per distinct point count: a second call reuses them
    ids = [id(s["pred"]) for s in ev._host_out]
    ev.step_host(host[:2])
    assert [id(s["pred"]) for s in ev._host_out] == ids
    h2d, d2h = ev.host_bytes_per_scan(scans[0][0].shape[0], T)
    assert h2d == scans[0][0].shape[0] * 20 + T * C * H * W * 4 and d2h == scans[0][0].shape[0] * 8 + H * W * 20


def test_summary_is_idempotent_and_leaves_accumulators_alone(cuda):
    H, W, C = 16, 256, 20
    scans, offs, xyzi, raw, logits = _batch(cuda, 2)
    ev = ScanEvaluator(H, W, C, device=cuda)
    ev.step_device(xyzi, raw, offs, logits.to(cuda))
    a = ev.summary()
    cm = ev.confmat.clone()
    b = ev.summary()
    assert np.array_equal(a["confmat"], b["confmat"]) and a["mIoU"] == b["mIoU"] or (np.isnan(a["mIoU"]) and np.isnan(b["mIoU"]))
    assert torch.equal(ev.confmat, cm)
    ev.step_device(xyzi, raw, offs, logits.to(cuda))
    assert int(ev.confmat.sum()) == 2 * int(cm.sum())


def test_scan_evaluator_capture_replay(cuda):
    H, W, C = 16, 256, 20
    scans, offs, xyzi, raw, logits = _batch(cuda, 3)
    lg = logits.to(cuda)
    ev = ScanEvaluator(H, W, C, device=cuda)
    eager = ev.step_device(xyzi, raw, offs, lg)
    cm1 = ev.confmat.clone()
    ev.reset()
    out = ev.capture(xyzi, raw, offs, lg)
    assert int(ev.confmat.sum()) == 0
    l0 = ev.launches
    ev.replay()
    torch.cuda.synchronize()
    assert ev.launches - l0 >= 6
    assert torch.equal(ev.confmat, cm1)
    for k in ("pred", "H_norm", "MI_norm", "conf", "point_labels", "normals"):
        assert torch.equal(out[k], eager[k]), k
    lg.mul_(0.5)                                            # new data written INTO the static input
    ev.replay()
    torch.cuda.synchronize()
    ref = ops.reduce_metrics(lg, eager["label"], kind="logits", conf_mode=ops.CONF_RENORM)
    assert torch.equal(out["H_norm"], ref["H_norm"])


def test_auroc_override_outside_unit_range_is_rescaled_not_saturated(cuda):
    from semanticlidarunc_b200.metrics.auroc import AUROCAggregator
    g = torch.Generator().manual_seed(0)
    C = 20
    probs = torch.softmax(torch.randn(1, C, 16, 64, generator=g) * 2, dim=1)
    labels = torch.randint(1, C, (1, 16, 64), generator=g)
    ent = -(probs * probs.clamp_min(1e-12).log()).sum(1)                # un-normalised entropy, up to ln C > 1
    a = AUROCAggregator(mode="probs", score="entropy_norm")
    a.update(probs.to(cuda), labels.to(cuda), score_override=ent.to(cuda))
    b = AUROCAggregator(mode="probs", score="entropy_norm")
    b.update(probs.to(cuda), labels.to(cuda), score_override=(ent / np.log(C)).to(cuda))
    pred = probs.argmax(1)
    ref = om.auroc_error_detection(ent.reshape(-1).numpy(), (pred != labels).reshape(-1).numpy())
    assert a.compute()[0] == pytest.approx(b.compute()[0], abs=1e-9)
    assert a.compute()[0] == pytest.approx(ref, abs=2e-4)
    with pytest.raises(ValueError):
        a.update(probs.to(cuda), labels.to(cuda), score_override=(-ent).to(cuda))
    s, e = a._scores, a._is_error
    assert s.numel() == e.numel() == 16 * 64 and abs(float(e.float().mean()) - float((pred != labels).float().mean())) < 1e-6


def test_accuracy_aggregator_accepts_long_ignore_lists(cuda):
    from semanticlidarunc_b200.models.evaluator import UncertaintyAccuracyAggregator
    g = torch.Generator().manual_seed(1)
    lab = torch.randint(0, 20, (2, 16, 64), generator=g)
    pred = torch.randint(0, 20, (2, 16, 64), generator=g)
    unc = torch.rand(2, 16, 64, generator=g)
    ign = (0, 2, 3, 4, 5, 7, 8)
    a = UncertaintyAccuracyAggregator()
    a.update(lab.to(cuda), pred.to(cuda), unc.to(cuda), ignore_ids=ign)
    keep = ~torch.isin(lab, torch.tensor(ign))
    assert a._seen == int(keep.sum())
    st = a.binned_accuracy(num_bins=10)
    n_ref, acc_ref = om.binned_accuracy(unc[keep].numpy(), (pred == lab)[keep].numpy(), 10)
    assert np.array_equal(st["n"].to_numpy(), n_ref)


# ---------------------------------------------------------------------------------------------- packed (f32x2) kernels
def _switch(name, on):
    return getattr(_lib.lib(), name)(int(on))


@pytest.mark.parametrize("C", [20, 7])
@pytest.mark.parametrize("from_outputs", [True, False])
def test_packed_evidential_kernel_agrees_with_one_pixel_kernel(cuda, C, from_outputs):
    B, H, W = 3, 8, 130
    x, lab = synth.synth_evidential_logits(3, B, C, H, W)
    x, lab = x.to(cuda), lab.to(cuda)
    if not from_outputs:
        x = ops.evidential_reduce(x, from_outputs=True, want=("alpha",))["alpha"]
    want = ("alpha", "pred", "conf", "H", "AU", "EU", "MI")
    res = []
    for off in (0, 1):
        _switch("slu_debug_no_packed_evidential", off)
        cm, bins = ops.new_confmat(C, cuda), ops.new_ece_bins(15, cuda)
        n0 = _lib.launch_count()
        r = ops.evidential_reduce(x, lab, from_outputs=from_outputs, ignore_index=0, confmat=cm, ece_bins=bins, want=want)
        assert _lib.launch_count() - n0 == 1
        res.append((r, cm, bins))
    _switch("slu_debug_no_packed_evidential", 0)
    (a, cma, ba), (b, cmb, bb) = res
    # same formulas, but ptxas schedules / contracts the two kernels differently: concentrations agree to an ulp or two,
    # the integer results exactly
    assert ((a["alpha"] - b["alpha"]).abs() <= 6e-7 * b["alpha"].abs()).all()
    assert torch.equal(a["pred"], b["pred"]) and (a["conf"] - b["conf"]).abs().max().item() < 1e-7
    assert torch.equal(cma, cmb) and torch.equal(ba[:2], bb[:2])
    for k in ("H", "AU", "EU", "MI"):
        assert (a[k] - b[k]).abs().max().item() < 5e-6, k


@pytest.mark.parametrize("C", [20, 6])
def test_packed_loss_kernels_agree_with_one_pixel_kernels(cuda, C):
    B, H, W = 2, 8, 258
    x, lab = synth.synth_evidential_logits(5, B, C, H, W)
    x, lab = x.to(cuda), lab.to(cuda)
    alpha = ops.evidential_reduce(x, from_outputs=True, want=("alpha",))["alpha"]
    alpha[0, 1, 0, :8] = 0.25                              # concentrations below 1: the out-of-line special-function path
    out = {}
    for off in (0, 1):
        _switch("slu_debug_no_packed_loss", off)
        out[off] = (ops.evidential_loss_fused(x, lab, ignore=(0,)), ops.dirichlet_loss(alpha, lab, ignore=(0,)))
    _switch("slu_debug_no_packed_loss", 0)
    (fa, da), (fb, db) = out[0], out[1]
    assert fa["sums"][2] == fb["sums"][2] and da["sums"][2] == db["sums"][2]
    np.testing.assert_allclose(fa["sums"].cpu().numpy(), fb["sums"].cpu().numpy(), rtol=2e-6)
    np.testing.assert_allclose(da["sums"].cpu().numpy(), db["sums"].cpu().numpy(), rtol=2e-6)
    for ga, gb in ((fa["grad"], fb["grad"]), (da["grad_mse"], db["grad_mse"]), (da["grad_kl"], db["grad_kl"])):
        scale = gb.abs().max().item()
        assert (ga - gb).abs().max().item() <= 1e-5 * scale
        assert torch.equal(ga == 0, gb == 0)               # masked pixels / the true class: exact zeros in both


def test_loss_step_api_self_cleaning_state_and_graph(cuda):
    from semanticlidarunc_b200.losses.evidential import EvidentialLoss
    B, C, H, W = 2, 20, 16, 128
    x, lab = synth.synth_evidential_logits(9, B, C, H, W)
    x, lab = x.to(cuda), lab.to(cuda)
    ref = ops.evidential_loss_fused(x, lab, ignore=(0,))
    n = float(ref["sums"][2])
    crit = EvidentialLoss(1.0, 0.05, ignore_index=0)
    for _ in range(3):                                     # the same buffers serve every step
        l4, g = crit.forward_backward(x, lab)
        assert torch.equal(g, ref["grad"])
        assert float(l4[3]) == n
        assert float(l4[1]) == pytest.approx(float(ref["sums"][0]) / n, rel=1e-6)
        assert float(l4[2]) == pytest.approx(float(ref["sums"][1]) / n, rel=1e-6)
        assert float(l4[0]) == pytest.approx((float(ref["sums"][0]) + 0.05 * float(ref["sums"][1])) / n, rel=1e-6)
    count, state = crit._work_buffers(x.device)
    assert float(count) == 0.0 and bool((state == 0).all())
    # autograd path: same values, gradient = kernel gradient times the upstream gradient
    xo = x.clone().requires_grad_(True)
    loss, mse, kl = crit(xo, lab)
    (2.0 * loss).backward()
    assert torch.equal(xo.grad, 2.0 * ref["grad"]) and float(loss) == float(l4[0])
    # one CUDA graph over static buffers
    xs = x.clone()
    l4s, gs = crit.capture(xs, lab)
    crit.replay(); torch.cuda.synchronize()
    assert torch.equal(gs, ref["grad"]) and torch.equal(l4s, l4)
    xs.mul_(0.5); crit.replay(); torch.cuda.synchronize()
    ref2 = ops.evidential_loss_fused(x * 0.5, lab, ignore=(0,))
    assert torch.equal(gs, ref2["grad"])


# ---------------------------------------------------------------------------------------------- edge points settled on the host
def test_drop_in_projection_settles_points_on_bin_edges_like_numpy(cuda):
    """Points planted exactly ON column / row edges and 1-2 ulp either side of them: the literal drop-in must return
    numpy's (the reference's) image whatever side CUDA's atan2 lands on; the kernel flags such scans and the host settles
    them (dataset/utils.py::_settle_edge_points)."""
    from semanticlidarunc_b200.dataset.utils import spherical_projection
    H, W = 16, 256
    xyzi, raw = synth.synth_scan(21, "tiny")
    pc = np.concatenate([xyzi.astype(np.float64), np.arange(xyzi.shape[0], dtype=np.float64)[:, None] + 1.0], axis=1)
    edges_w = np.linspace(-np.pi, np.pi, W)
    rng = np.random.default_rng(0)
    planted = []
    for k, e in enumerate(edges_w[5:250:7]):
        for ulps in (-2, -1, 0, 1, 2):
            phi = e
            for _ in range(abs(ulps)):
                phi = np.nextafter(phi, np.inf if ulps > 0 else -np.inf)
            r, el = rng.uniform(5, 60), rng.uniform(-0.3, 0.1)
            planted.append([r * np.cos(el) * np.cos(phi), r * np.cos(el) * np.sin(phi), r * np.sin(el), 0.5, 0.0])
    planted = np.asarray(planted)
    planted[:, 4] = np.arange(len(planted)) + 1e6
    pc2 = np.concatenate([pc, planted], axis=0)
    img, _, (tmin, tmax), _ = spherical_projection(pc2, H, W)
    ref, _, (rmin, rmax), _ = oproj.spherical_projection(pc2, H, W)
    assert (tmin, tmax) == (rmin, rmax)
    assert np.array_equal(img, ref)
    assert spherical_projection.last_near_edge > 0          # the planted points were seen as ambiguous ...
    # ... and a scan without them is not touched by the host at all
    img0, *_ = spherical_projection(pc, H, W)
    assert spherical_projection.last_near_edge == 0 and np.array_equal(img0, oproj.spherical_projection(pc, H, W)[0])


# ---------------------------------------------------------------------------------------------- single-sample logits kernel
@pytest.mark.parametrize("shape", [(16, 20, 64, 2048), (3, 7, 5, 37), (1, 20, 9, 130)])
def test_single_sample_logits_kernel_agrees_with_general_single_kernel(cuda, shape):
    """reduce_single_logits_kernel (arg-max on the logits, e_max / S confidence, thread-private reliability cells) against
    reduce_single_kernel (arg-max on the probabilities, shared-atomic histograms): identical confidence / entropy maps,
    identical counters up to pixels whose two best probabilities round to the same float."""
    B, C, H, W = shape
    g = torch.Generator().manual_seed(B * 31 + C)
    x = (torch.randn((B, C, H, W), generator=g) * 3.0).to(cuda)
    lab = torch.randint(-1, C + 1, (B, H, W), generator=g).to(cuda)
    x[0, :, 0, 0] = float("nan"); x[0, 1, 0, 1] = float("inf"); x[0, :, 0, 2] = -float("inf")
    out = []
    for off in (0, 1):
        _lib.lib().slu_debug_reduce_no_private(off)
        cm, bins = ops.new_confmat(C, cuda), ops.new_ece_bins(15, cuda)
        r = ops.reduce_metrics(x, lab, kind="logits", ignore_index=0, confmat=cm, ece_bins=bins, want=("p_bar", "pred", "conf", "H_norm", "MI_norm"))
        out.append((r, cm, bins))
    _lib.lib().slu_debug_reduce_no_private(0)
    (a, cma, ba), (b, cmb, bb) = out
    same = a["pred"] == b["pred"]
    assert int((~same).sum()) <= max(3, same.numel() // 20000)
    for k in ("conf", "H_norm", "MI_norm"):
        assert torch.equal(torch.nan_to_num(a[k][same], nan=-7.0), torch.nan_to_num(b[k][same], nan=-7.0)), k
    assert torch.equal(torch.nan_to_num(a["p_bar"], nan=-7.0), torch.nan_to_num(b["p_bar"], nan=-7.0))
    assert int(cma.sum()) == int(cmb.sum()) and int((cma - cmb).abs().sum()) <= 2 * int((~same).sum())
    assert torch.equal(ba[0], bb[0]) and torch.equal(ba[2], bb[2]) and int((ba[1] - bb[1]).abs().sum()) <= int((~same).sum())
    ref = ou.mc_reduce(x[None].cpu())
    ok = torch.isfinite(ref["H_norm"])
    assert (a["H_norm"].cpu()[ok] - ref["H_norm"][ok]).abs().max() < 1e-5


def test_prefilter_arctangent_error_bound(cuda):
    """The fp32 arctangent that decides which points may skip the fp64 angles: its error must stay far below the
    prefilter's margin (8e-6 rad), and degenerate inputs must come out as NaN (-> exact path)."""
    g = torch.Generator().manual_seed(0)
    n = 4_000_000
    ang = torch.rand(n, generator=g, dtype=torch.float64) * (2 * np.pi) - np.pi
    rad = 10 ** (torch.rand(n, generator=g, dtype=torch.float64) * 4.5 - 2)
    x, y = (rad * torch.cos(ang)).float(), (rad * torch.sin(ang)).float()
    x[:4] = torch.tensor([0.0, float("inf"), float("nan"), 1e38]); y[:4] = torch.tensor([0.0, 1.0, 1.0, 1.0])
    out = torch.empty(n, dtype=torch.float32, device=cuda)
    xd, yd = x.to(cuda), y.to(cuda)
    _lib.check(_lib.lib().slu_diag_fast_atan2(_lib.ptr(yd), _lib.ptr(xd), n, _lib.ptr(out), _lib.stream_ptr()), "slu_diag_fast_atan2")
    out = out.cpu()
    assert bool(torch.isnan(out[:4]).all())
    ref = torch.atan2(y[4:].double(), x[4:].double())
    err = (out[4:].double() - ref).abs()
    err = torch.minimum(err, 2 * np.pi - err)
    assert float(err.max()) < 1.0e-6, float(err.max())


@pytest.mark.parametrize("increasing", [False, True])
def test_drop_in_projection_with_caller_supplied_row_edges(cuda, increasing):
    """bins_h (src/dataset/utils.py:330-338): non-uniform row edges, in either direction, against the numpy oracle."""
    from semanticlidarunc_b200.dataset.utils import spherical_projection
    H, W = 16, 256
    xyzi, raw = synth.synth_scan(33, "tiny")
    pc = np.concatenate([xyzi.astype(np.float64), np.arange(xyzi.shape[0], dtype=np.float64)[:, None] + 1.0], axis=1)
    rng = np.random.default_rng(5)
    edges = np.sort(rng.uniform(-0.40, 0.20, H))                       # non-uniform, strictly increasing
    bins = edges if increasing else edges[::-1]
    img, alpha, tr, _ = spherical_projection(pc, H, W, bins_h=bins)
    ref, alpha_ref, tr_ref, _ = oproj.spherical_projection(pc, H, W, bins_h=bins)
    assert np.array_equal(img, ref) and tr == tr_ref
    assert np.array_equal(alpha, alpha_ref)
    with pytest.raises(ValueError):
        spherical_projection(pc, H, W, bins_h=np.array([0.0, 1.0, 0.5] + [2.0] * (H - 3)))


def test_kitti_loader_settles_edge_points(cuda, tmp_path):
    """SemanticKitti.__getitem__ on a scan with points planted on / 1-2 ulp around column edges: the five tensors must
    equal the reference's (oracle kitti_item), whatever side the device's angles fall on."""
    from semanticlidarunc_b200.dataset.dataloader_semantic_KITTI import SemanticKitti
    H, W = 16, 256
    xyzi, raw = synth.synth_scan(41, "tiny")
    edges_w = np.linspace(-np.pi, np.pi, W)
    rng = np.random.default_rng(2)
    extra = []
    for e in edges_w[9:240:11]:
        for ulps in (-1, 0, 1):
            phi = e
            for _ in range(abs(ulps)):
                phi = np.nextafter(phi, np.inf if ulps > 0 else -np.inf)
            r, el = rng.uniform(5, 60), rng.uniform(-0.3, 0.1)
            extra.append([r * np.cos(el) * np.cos(phi), r * np.cos(el) * np.sin(phi), r * np.sin(el), 0.25])
    xyzi2 = np.concatenate([xyzi, np.asarray(extra, dtype=np.float32)]).astype(np.float32)
    raw2 = np.concatenate([raw, np.full(len(extra), raw[0], dtype=np.uint32)])
    b, l = tmp_path / "000000.bin", tmp_path / "000000.label"
    xyzi2.tofile(b); raw2.tofile(l)
    ds = SemanticKitti([(str(b), str(l))], projection=(H, W), resize=False)
    got = ds[0]
    ref = oproj.kitti_item(xyzi2, raw2, build_id_lut(), projection=(H, W), resize=False)
    for k, (a, r_) in enumerate(zip(got, ref)):
        if k == 3:
            continue                                     # normals: tolerance-checked elsewhere
        assert np.array_equal(a.numpy(), r_), k


def test_class_pair_loss_kernel_on_odd_shapes(cuda):
    """Odd HW cannot take the pixel-pair kernel: the class-pair packed kernel runs instead and must agree with the
    one-pixel-per-thread kernel (same formulas as the pixel-pair kernel, horizontal sums in another order)."""
    B, C, H, W = 2, 20, 7, 37
    x, lab = synth.synth_evidential_logits(11, B, C, H, W)
    x, lab = x.to(cuda), lab.to(cuda)
    n0 = _lib.launch_count()
    a = ops.evidential_loss_fused(x, lab, ignore=(0,))
    assert _lib.launch_count() - n0 == 2
    _switch("slu_debug_no_packed_loss", 1)
    b = ops.evidential_loss_fused(x, lab, ignore=(0,))
    _switch("slu_debug_no_packed_loss", 0)
    assert a["sums"][2] == b["sums"][2]
    np.testing.assert_allclose(a["sums"].cpu().numpy(), b["sums"].cpu().numpy(), rtol=3e-6)
    assert (a["grad"] - b["grad"]).abs().max().item() <= 1e-5 * b["grad"].abs().max().item()
    assert torch.equal(a["grad"] == 0, b["grad"] == 0)
    for Cc in (5, 6):                                     # padded class pairs
        x2, lab2 = synth.synth_evidential_logits(12, 1, Cc, 5, 9)
        x2, lab2 = x2.to(cuda), lab2.to(cuda)
        a = ops.evidential_loss_fused(x2, lab2, ignore=(0,))
        _switch("slu_debug_no_packed_loss", 1)
        b = ops.evidential_loss_fused(x2, lab2, ignore=(0,))
        _switch("slu_debug_no_packed_loss", 0)
        np.testing.assert_allclose(a["sums"].cpu().numpy(), b["sums"].cpu().numpy(), rtol=3e-6)
        assert (a["grad"] - b["grad"]).abs().max().item() <= 1e-5 * b["grad"].abs().max().item()


def test_vectorised_normals_kernel_is_bit_identical_to_scalar(cuda):
    """frame_normals4_kernel (four pixels per thread, 16-byte accesses) against frame_normals_kernel, which an unaligned view
    of the same planes falls back to."""
    g = torch.Generator().manual_seed(9)
    B, H, W = 3, 16, 256
    xyz = torch.randn((B, 3, H, W), generator=g).to(cuda)
    a = ops.frame_normals(xyz)                                             # aligned: vectorised kernel
    pad = torch.empty(B * 3 * H * W + 1, dtype=torch.float32, device=cuda)
    view = pad[1:].view(B, 3, H, W)                                        # 4-byte offset: scalar kernel
    view.copy_(xyz)
    b = ops.frame_normals(view)
    assert torch.equal(a, b)


def test_near_edge_queue_overflow_falls_back_inline(cuda):
    """The per-block queue of near-edge points holds 1024 entries (DEFER_CAP); with one CTA per SM over a 16-scan batch a
    block sees ~2 200 points, ALL planted within the prefilter's margin of a column edge here, so most of them overflow the
    queue and take the inline fp64 path.  Result must equal the all-fp64 kernels bit for bit (and the oracle)."""
    import os
    H, W = 16, 2048
    edges = np.linspace(-np.pi, np.pi, W)
    rng = np.random.default_rng(3)
    scans = []
    for b in range(16):
        n = 20_000
        ang = edges[rng.integers(2, W - 2, n)] + rng.choice([0.0, 1e-7, -1e-7, 2e-6, -2e-6], n)
        el = rng.uniform(-0.4, 0.05, n)
        r = rng.uniform(3, 70, n)
        xyzi = np.stack([r * np.cos(el) * np.cos(ang), r * np.cos(el) * np.sin(ang), r * np.sin(el), rng.uniform(0, 1, n)], 1).astype(np.float32)
        scans.append((xyzi, synth.synth_scan(b, "tiny")[1][:1].repeat(n)))
    offs = np.concatenate([[0], np.cumsum([s[0].shape[0] for s in scans])]).astype(np.int64)
    dx = torch.from_numpy(np.concatenate([s[0] for s in scans])).to(cuda)
    dr = torch.from_numpy(np.concatenate([s[1] for s in scans]).view(np.int32)).to(cuda)
    dl = torch.from_numpy(build_id_lut()).to(cuda)
    old = os.environ.get("SLU_PT_CTAS_PER_SM")
    os.environ["SLU_PT_CTAS_PER_SM"] = "1"
    try:
        fast = ops.project_batch(dx, dr, offs, H, W, lut=dl)
        prev = _lib.lib().slu_debug_project_exact(1)
        try:
            exact = ops.project_batch(dx, dr, offs, H, W, lut=dl)
        finally:
            _lib.lib().slu_debug_project_exact(prev)
    finally:
        if old is None:
            del os.environ["SLU_PT_CTAS_PER_SM"]
        else:
            os.environ["SLU_PT_CTAS_PER_SM"] = old
    for k in ("pix", "winner", "img", "label", "theta", "diag"):
        assert torch.equal(fast[k], exact[k]), k
    o = oproj.kitti_frame(scans[3][0], scans[3][1], H, W, build_id_lut())
    assert np.array_equal(fast["pix"][offs[3]:offs[4]].cpu().numpy().astype(np.int64), o["pix"])


def test_config1_full_size_single_scan_vs_oracle(cuda):
    """BASELINE.json configs[0] at its real size: one 64x2048 scan, single-pass logits (T=1, C=20) -> entropy map within 1e-5 of
    the oracle, confusion counts and reliability-bin counts exact on margin-enforced inputs, ECE within 1e-5 relative."""
    from tests.helpers import enforce_margins
    C, H, W = 20, 64, 2048
    x, lab = synth.synth_mc_logits(17, 1, 1, C, H, W)
    x, lab = enforce_margins(x, lab, conf_renorm=False)
    cm, bins = ops.new_confmat(C, cuda), ops.new_ece_bins(15, cuda)
    out = ops.reduce_metrics(x[0].to(cuda), lab.to(cuda), kind="logits", conf_mode=ops.CONF_RAW, ignore_index=0, confmat=cm, ece_bins=bins,
                             want=("p_bar", "pred", "conf", "H_norm"))
    ref = ou.mc_reduce(x)
    assert torch.equal(out["pred"].cpu(), ref["pred"])
    assert (out["H_norm"].cpu() - ref["H_norm"]).abs().max().item() < 1e-5
    assert torch.equal(cm.cpu(), om.confusion_counts(ref["pred"], lab, C))
    p = torch.softmax(x[0], dim=1)
    cf, pr = p.max(1)
    valid = lab != 0
    n_ref, c_ref, s_ref = om.ece_bin_counts(cf[valid].numpy(), (pr == lab)[valid].numpy(), 15)
    assert np.array_equal(bins[0].cpu().numpy(), n_ref) and np.array_equal(bins[1].cpu().numpy(), c_ref)
    ece = ops.ece_from_bins(bins)[0]
    ece_ref = om.ece_from_counts(n_ref, c_ref, s_ref)[0]
    assert abs(ece - ece_ref) <= 1e-5 * abs(ece_ref)


def test_frame_tensors_views_equal_the_copies(cuda):
    """ops.frame_tensors(img, label=...) hands out range / reflectivity / xyz / semantics as views of the projection's planes
    (no resample copy when nothing is resized, flipped or dropped): same values, shapes and dtypes as the copying path,
    contiguous per scan; any resize / flip / row dropping takes the copying kernel as before."""
    scans = [synth.synth_scan(70 + i, "hdl64", n_points=30_000 + 1000 * i) for i in range(3)]
    offs = np.concatenate([[0], np.cumsum([s[0].shape[0] for s in scans])])
    xyzi = torch.from_numpy(np.concatenate([s[0] for s in scans])).to(cuda)
    raw = torch.from_numpy(np.concatenate([s[1] for s in scans]).view(np.int32)).to(cuda)
    lut = torch.from_numpy(build_id_lut()).to(cuda)
    proj = ops.project_batch(xyzi, raw, offs, 64, 2048, lut=lut)
    a = ops.frame_tensors(proj["img"])
    b = ops.frame_tensors(proj["img"], label=proj["label"])
    for k in ("range", "reflectivity", "xyz", "normals", "semantics"):
        assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), k
    assert b["xyz"].data_ptr() == proj["img"].data_ptr() and b["semantics"].data_ptr() == proj["label"].data_ptr()      # views, not copies
    assert b["xyz"][1].is_contiguous() and b["range"][2].is_contiguous()
    c = ops.frame_tensors(proj["img"], label=proj["label"], flip=[True, False, True])                               # flip: the copying kernel
    d = ops.frame_tensors(proj["img"], flip=[True, False, True])
    assert c["xyz"].data_ptr() != proj["img"].data_ptr() and all(torch.equal(c[k], d[k]) for k in c)
