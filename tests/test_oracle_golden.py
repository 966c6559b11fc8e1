"""Pins the oracle: every oracle function is checked against outputs of the UNMODIFIED reference
stored in tests/golden/ (written by oracle/gen_golden.py).  When /root/reference is present (build
container) the reference is also re-run live on fresh seeds."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import _refshim
from oracle import losses as ol
from oracle import metrics as om
from oracle import projection as oproj
from oracle import uncertainty as ou
from semanticlidarunc_b200 import synth
from semanticlidarunc_b200.dataset.definitions import build_id_lut

HERE = os.path.dirname(os.path.abspath(__file__))
SMALL = ["tiny_auto", "tiny_range", "tiny_farthest", "edge", "ragged_1pt"]


def _pc(xyzi, raw):
    sem = build_id_lut()[(raw & 0xFFFF).astype(np.int64)].astype(np.int64)
    return np.concatenate([xyzi, sem[:, None]], axis=-1)


@pytest.mark.parametrize("name", SMALL)
def test_projection_oracle_vs_golden(golden, name):
    g = golden("projection_small.npz")
    H, W = (int(v) for v in g[name + "/hw"])
    tr = g[name + "/theta_range"]
    tr = None if np.isnan(tr).any() else (float(tr[0]), float(tr[1]))
    far = bool(int(g[name + "/largest_first"]))
    pc = _pc(g[name + "/xyzi"], g[name + "/raw"])
    img, alpha, (tmin, tmax), _ = oproj.spherical_projection(pc, H, W, theta_range=tr, sort_largest_first=far)
    assert img.dtype == np.float32 and np.array_equal(img, g[name + "/img"])
    row, col, _ = oproj.projection_indices(pc, H, W, tr)
    pix = row * W + col
    assert np.array_equal(pix, g[name + "/pix"])
    win = oproj.depth_test(oproj.point_range(pc), pix, H * W, farthest_wins=far)
    assert np.array_equal(win, g[name + "/winner"])
    if tr is None:
        assert np.array_equal(np.array([tmin, tmax]), g[name + "/theta"])


def test_kitti_frame_oracle_vs_golden(golden):
    g = golden("kitti_loader.npz")
    H, W = (int(v) for v in g["hw"])
    o = oproj.kitti_frame(g["xyzi"], g["raw"], H, W, build_id_lut())
    for k in ("xyz", "range", "reflectivity", "semantics"):
        assert o[k].dtype == g[k].dtype and np.array_equal(o[k], g[k]), k


def test_mc_oracle_vs_golden(golden):
    g = golden("mc_reduce.npz")
    for name in ("mc_small", "mc_peaked", "mc_c7"):
        r = ou.mc_reduce(torch.from_numpy(g[name + "/logits"]))
        assert np.array_equal(r["pred"].numpy(), g[name + "/pred"])
        for k in ("p_bar", "H_norm", "MI_norm"):
            assert np.allclose(r[k].numpy(), g[name + "/" + k], rtol=1e-6, atol=1e-7), (name, k)
        probs = torch.softmax(torch.from_numpy(g[name + "/logits"]), dim=2)
        assert np.allclose(ou.predictive_entropy_mc(probs).numpy(), g[name + "/H_mc"], rtol=1e-6, atol=1e-7)


def test_evidential_oracle_vs_golden(golden):
    g = golden("evidential.npz")
    for name in ("ev_small", "ev_strong"):
        o = torch.from_numpy(g[name + "/outputs"])
        r = ou.evidential_reduce(o, 20)
        assert np.array_equal(r["pred"].numpy(), g[name + "/pred"])
        for k in ("alpha", "H_norm", "AU", "EU", "MI_norm"):
            assert np.allclose(r[k].numpy(), g[name + "/" + k], rtol=1e-6, atol=1e-7), (name, k)


def test_metrics_oracle_vs_golden(golden):
    g = golden("metrics.npz")
    C = 20
    cm = om.confusion_counts(torch.from_numpy(g["iou/preds"]), torch.from_numpy(g["iou/targets"]), C)
    assert np.array_equal(cm.numpy(), g["iou/confmat"])
    miou, iou = om.iou_from_confmat(cm, test_mask=[0] + [1] * (C - 1), ignore_gt=[0])
    assert miou == float(g["iou/miou"]) and np.array_equal(iou.numpy(), g["iou/per_class"], equal_nan=True)
    for mode in ("alpha", "logits", "probs"):
        x, lab = torch.from_numpy(g[f"ece_{mode}/preds"]), torch.from_numpy(g[f"ece_{mode}/labels"])
        conf, corr = om.ece_samples(x, lab, mode, ignore_index=0)
        assert np.array_equal(conf.numpy(), g[f"ece_{mode}/conf"]) and np.array_equal(corr.numpy(), g[f"ece_{mode}/correct"])
        n, acc, avg = om.ece_reference_stats(conf.numpy(), corr.numpy(), 15)
        assert np.array_equal(n, g[f"ece_{mode}/n"])
        ece, mce = om.ece_from_stats(n, acc, avg)
        assert np.allclose([ece, mce], g[f"ece_{mode}/ece_mce"], rtol=1e-12)
        # the exact streaming-histogram form agrees with the reference's float32 np.histogram sums
        n2, nc2, cs2 = om.ece_bin_counts(conf.numpy(), corr.numpy(), 15)
        assert np.array_equal(n2, n)
        ece2, mce2 = om.ece_from_counts(n2, nc2, cs2)
        assert abs(ece2 - ece) <= 1e-5 * ece and abs(mce2 - mce) <= 1e-5 * mce
    assert int(g["ece_empty/len"]) == 2


def test_losses_oracle_vs_golden(golden):
    g = golden("losses.npz")
    target = torch.from_numpy(g["target"])
    import functools
    for name, fn in (("mse", ol.dirichlet_mse), ("kl", ol.kl_offclasses_to_uniform), ("nll", ol.nll_dirichlet_categorical),
                     ("dce", ol.digamma_dirichlet_ce), ("brier", ol.brier_dirichlet),
                     ("brier_sref", functools.partial(ol.brier_dirichlet, s_ref=30.0))):
        a = torch.from_numpy(g["alpha"]).clone().requires_grad_(True)
        loss = fn(a, target, ignore_index=0)
        (grad,) = torch.autograd.grad(loss, a)
        assert np.allclose(loss.detach().numpy(), g[name + "/loss"], rtol=1e-6)
        assert np.allclose(grad.numpy(), g[name + "/grad"], rtol=1e-5, atol=1e-9)


def test_loss_terms_oracle_vs_golden(golden):
    """ComplementKLUniform, WrongLowEvidence, EvidenceReg(Band), conf-weighted KL, LogitRegularizer restatements
    against the reference modules' recorded values and autograd gradients."""
    from tests.loss_term_cases import NAMES, cases
    g = golden("loss_terms.npz")
    cs = cases(torch.from_numpy(g["target"]), torch.from_numpy(g["keep"]))
    assert sorted(cs) == sorted(NAMES)
    for name in NAMES:
        key, fn, _ = cs[name]
        x = torch.from_numpy(g[key]).clone().requires_grad_(True)
        loss = fn(x)
        (grad,) = torch.autograd.grad(loss, x)
        assert np.allclose(loss.detach().numpy(), g[name + "/loss"], rtol=1e-6), name
        assert np.allclose(grad.numpy(), g[name + "/grad"], rtol=1e-5, atol=1e-9), name
        assert float(np.abs(g[name + "/grad"]).max()) > 0.0, name        # every case exercises its gradient


def _np_score_hist(score, is_err, M):
    b = np.minimum((np.clip(score, 0, 1).astype(np.float64) * M).astype(np.int64), M - 1)
    hist = np.zeros((2, M), np.int64)
    np.add.at(hist, (is_err.astype(np.int64), b), 1)
    return hist


@pytest.mark.parametrize("name", ["cont", "ties", "two", "one", "allwrong"])
def test_aurc_oracle_and_histogram_form_vs_golden(golden, name):
    """rc_curve_stats / aurc_from_risks_confids (src/metrics/aurc.py:7-45): the restatement is exact; the
    fixed-resolution histogram form the device path uses stays within 1e-6 on continuous confidences and within
    1e-4 where the reference's own result depends on how argsort orders exact ties."""
    from semanticlidarunc_b200.metrics import aurc as A
    g = golden("aurc.npz")
    conf, risks = g[name + "/conf"], g[name + "/risks"]
    cov, sel, w = om.rc_curve_stats(risks, conf)
    assert np.array_equal(cov, g[name + "/cov"]) and np.array_equal(sel, g[name + "/sel"]) and np.array_equal(w, g[name + "/w"])
    a, e, _, _ = om.aurc_from_risks_confids(risks, conf)
    assert a == float(g[name + "/aurc"]) and e == float(g[name + "/eaurc"])
    hist = _np_score_hist(np.float32(1.0) - conf, risks, A.AURC_BINS)
    ah, eh, _, _ = A.aurc_from_hist(hist)
    tol = 1e-4 if name == "ties" else 1e-6
    assert abs(ah - a) <= tol and abs(eh - e) <= tol, (ah - a, eh - e)
    n, ne = int(risks.size), int(risks.sum())
    opt_ref = float((np.cumsum(np.sort(risks.astype(np.float64))) / np.arange(1, n + 1)).sum() / n)
    assert abs(A.optimal_aurc(n, ne) - opt_ref) <= 1e-12


def test_full_size_projection_digests_cpu():
    """One full HDL-64 scan through the oracle must hit the reference's digests (others run on GPU)."""
    import hashlib
    with open(os.path.join(HERE, "golden", "MANIFEST.json")) as f:
        c = json.load(f)["projection_full"]["hdl64_seed0"]
    xyzi, raw = synth.synth_scan(c["seed"], c["sensor"])
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert sha(xyzi) == c["xyzi_sha"] and sha(raw) == c["raw_sha"]
    pc = _pc(xyzi, raw)
    row, col, (tmin, tmax) = oproj.projection_indices(pc, c["H"], c["W"])
    pix = row * c["W"] + col
    assert sha(pix) == c["pix_sha"]
    win = oproj.depth_test(oproj.point_range(pc), pix, c["H"] * c["W"])
    assert sha(win) == c["winner_sha"] and int((win >= 0).sum()) == c["occupied"]
    assert float(tmin) == c["theta_min"] and float(tmax) == c["theta_max"]


def test_linspace_restatement_matches_numpy():
    """The CUDA kernels rebuild numpy.linspace as fl(fl(i*step)+start) with the last edge = stop."""
    rng = np.random.default_rng(0)
    for num in (2, 3, 16, 64, 128, 2048):
        for _ in range(20):
            a, b = np.sort(rng.uniform(-np.pi, np.pi, 2))
            ref = np.linspace(a, b, num)
            step = (b - a) / (num - 1)
            mine = np.arange(num, dtype=np.float64) * step + a
            mine[-1] = b
            assert np.array_equal(ref, mine)
    ref = np.linspace(-np.pi, np.pi, 2048)
    step = (np.pi - (-np.pi)) / 2047
    mine = np.arange(2048, dtype=np.float64) * step + (-np.pi)
    mine[-1] = np.pi
    assert np.array_equal(ref, mine)


@pytest.mark.skipif(not _refshim.available(), reason="reference tree not present (GPU box)")
def test_oracle_vs_live_reference_on_fresh_seeds():
    _refshim.install()
    from dataset.utils import spherical_projection as ref_proj
    import models.probability_helper as ph
    for seed in (5, 6):
        xyzi, raw = synth.synth_scan(seed, "tiny")
        pc = _pc(xyzi, raw)
        a = ref_proj(pc, 16, 256)
        b = oproj.spherical_projection(pc, 16, 256)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
    o, _ = synth.synth_evidential_logits(9, 1, 20, 4, 32)
    alpha = ph.to_alpha_concentrations_from_shape_and_scale(o[:, :20], o[:, 20:21])
    assert torch.equal(alpha, ou.to_alpha_concentrations_from_shape_and_scale(o[:, :20], o[:, 20:21]))
    assert torch.equal(ph.get_aleatoric_uncertainty(alpha), ou.get_aleatoric_uncertainty(alpha))
    assert torch.equal(ph.get_predictive_entropy_norm(alpha), ou.get_predictive_entropy_norm(alpha))


def test_kitti_item_oracle_vs_golden(golden):
    """Full __getitem__ restatement (resize, flip, yaw, normals through cv2) against the reference Dataset."""
    import hashlib
    g = golden("kitti_loader_aug.npz")
    lut = build_id_lut()
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    r = oproj.kitti_item(g["xyzi"], g["raw"], lut, projection=(16, 256), resize=True)
    for k, a in zip(("range", "reflectivity", "xyz", "normals", "semantics"), r):
        assert sha(a) == bytes(g["resize/" + k + "_sha"]).hex(), k
    r = oproj.kitti_item(g["xyzi"], g["raw"], lut, projection=(16, 256), resize=False, flip=True, yaw_deg=float(g["aug/angle"]))
    for k, a in zip(("range", "reflectivity", "xyz", "normals", "semantics"), r):
        assert np.array_equal(a, g["aug/" + k]), k


def test_auroc_and_binned_accuracy_oracle_vs_golden(golden):
    g = golden("metrics.npz")
    x, lab = torch.from_numpy(g["auroc/probs"]), torch.from_numpy(g["auroc/labels"])
    valid = (lab != 0).numpy().reshape(-1)
    p = om.to_probs(x, "probs")
    pred = p.argmax(1)
    err = (pred != lab).numpy().reshape(-1)[valid]
    H = (-(p.clamp_min(1e-12) * p.clamp_min(1e-12).log()).sum(1) / np.log(20)).numpy().reshape(-1)[valid]
    assert abs(om.auroc_error_detection(H, err) - float(g["auroc/probs_entropy_norm"])) < 1e-12
    ov = g["auroc/override"].reshape(-1)[valid]
    assert abs(om.auroc_error_detection(ov, err) - float(g["auroc/probs_override"])) < 1e-12
    for nb in (10, 20):
        n, acc = om.binned_accuracy(ov, ~err, nb)
        assert np.array_equal(n, g[f"ua/n_{nb}"]) and np.allclose(acc, g[f"ua/acc_{nb}"], equal_nan=True)


def test_cudal_and_thab_items_oracle_vs_golden(golden):
    import hashlib
    from semanticlidarunc_b200.dataset.dataloader_semantic_CUDAL import id_map as cudal_map
    g = golden("other_loaders.npz")
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    names = ("range", "reflectivity", "xyz", "normals", "semantics")
    r = oproj.kitti_item(g["cudal/xyzi"], g["cudal/raw"], build_id_lut(cudal_map), projection=(32, 256), resize=True,
                         theta_range=[-np.pi / 8, np.pi / 8], normalise_reflectivity=True)
    for k, a in zip(names, r):
        assert sha(a) == bytes(g[f"cudal/{k}_sha"]).hex(), k
    xyzi, raw = synth.synth_scan(int(g["thab/seed"]), "os1-128")
    for tag, kw in (("thab_plain", {}), ("thab_aug", {"flip": True, "yaw_deg": int(g["thab_aug/angle"])})):
        r = oproj.thab_item(xyzi, raw, build_id_lut(), **kw)
        for k, a in zip(names, r):
            assert sha(a) == bytes(g[f"{tag}/{k}_sha"]).hex(), (tag, k)


def test_wads_item_oracle_vs_golden(golden):
    import hashlib
    from semanticlidarunc_b200.dataset.dataloader_semantic_WADS import id_map as wads_map
    g = golden("other_loaders.npz")
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    names = ("range", "reflectivity", "xyz", "normals", "semantics")
    kw = dict(projection=(64, 256), theta_range=[-np.pi / 2, np.pi / 2], drop_empty_rows=True, resize_to=(1024, 64))
    r = oproj.kitti_item(g["wads/xyzi"], g["wads/raw"], build_id_lut(wads_map), resize=True, **kw)
    for k, a in zip(names, r):
        assert sha(a) == bytes(g[f"wads/{k}_sha"]).hex(), k
    r = oproj.kitti_item(g["wads/xyzi"], g["wads/raw"], build_id_lut(wads_map), resize=False, **kw)
    for k, a in zip(names, r):
        assert np.array_equal(a, g[f"wads_native/{k}"]), k
