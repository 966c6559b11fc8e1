#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):
    python oracle/gen_golden.py
Every array written here is an output of reference code imported from
/root/reference/src through oracle/_refshim.py; the only restated lines are the
per-point (row, col) formula, which is asserted to reproduce the reference
image before it is stored.  The two MC closures that live inside
Tester.test_epoch (src/models/tester.py:419-451) cannot be imported, so their
source is extracted from the reference file with `ast` at generation time and
executed as is.
"""
from __future__ import annotations

import ast
import hashlib
import json
import math
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the script's own directory must not shadow the reference's top-level packages (metrics, losses ...)
sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != os.path.join(ROOT, "oracle")]
sys.path.insert(0, ROOT)

from oracle import _refshim  # noqa: E402

_refshim.install()

from semanticlidarunc_b200 import synth  # noqa: E402
from semanticlidarunc_b200.dataset.definitions import build_id_lut  # noqa: E402
from oracle import projection as oproj  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def extract_closures(path, names):
    """Compile nested function definitions out of a reference source file, unmodified."""
    with open(path) as f:
        src = f.read()
    tree = ast.parse(src)
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names and node.name not in found:
            node.decorator_list = []           # @torch.no_grad() only
            found[node.name] = node
    mod = ast.Module(body=[found[n] for n in names], type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {"torch": torch, "math": math}
    exec(compile(mod, path, "exec"), ns)
    return [ns[n] for n in names]


# ---------------------------------------------------------------- projection
def gen_projection():
    from dataset.utils import spherical_projection as ref_proj

    cases = {}

    def run(name, xyzi, raw, H, W, theta_range=None, largest_first=False):
        lut = build_id_lut()
        sem = lut[(raw & 0xFFFF).astype(np.int64)].astype(np.int64)
        pc = np.concatenate([xyzi, sem[:, None]], axis=-1)                   # float64, as the loaders
        pc_id = np.concatenate([pc, np.arange(pc.shape[0])[:, None] + 1.0], axis=-1)
        img, alpha, (tmin, tmax), (pmin, pmax) = ref_proj(pc_id, H, W, theta_range=theta_range,
                                                          sort_largest_first=largest_first)
        winner = img[..., -1].reshape(-1).astype(np.int64) - 1                # -1 = empty
        row, col, _ = oproj.projection_indices(pc, H, W, theta_range)
        pix = row * W + col
        # the restated per-point indices must reproduce the reference image
        occ = winner >= 0
        assert (pix[winner[occ]] == np.nonzero(occ)[0]).all(), name
        assert set(np.unique(pix)) == set(np.nonzero(occ)[0]), name
        return {"img": img[..., :-1], "winner": winner, "pix": pix, "alpha": alpha,
                "theta": np.array([tmin, tmax], dtype=np.float64)}

    # small cases, stored in full (inputs included)
    small = {}
    xyzi, raw = synth.synth_scan(11, "tiny")
    small["tiny_auto"] = (xyzi, raw, 16, 256, None, False)
    small["tiny_range"] = (xyzi, raw, 16, 256, (-np.pi / 8, np.pi / 8), False)   # CUDAL-style clip
    small["tiny_farthest"] = (xyzi, raw, 16, 256, None, True)
    # edge cases: points on the +-pi seam, straight up/down, duplicates in a pixel
    e = xyzi[:400].copy()
    e[0, :3] = (-5.0, 0.0, 0.1)        # phi = +pi  -> wraps to last column
    e[1, :3] = (-5.0, -0.0, 0.1)       # phi = -pi
    e[2, :3] = (0.0, 0.0, 7.0)         # straight up
    e[3, :3] = (0.0, 0.0, -7.0)        # straight down
    e[4, :3] = (3.0, 0.0, 0.0)         # phi = 0, theta = 0
    e[5, :3] = (0.0, 3.0, 0.0)         # phi = pi/2
    e[6, :3] = e[7, :3] * 0.5          # same direction, nearer -> must win its pixel
    small["edge"] = (e, raw[:400], 8, 32, None, False)
    small["ragged_1pt"] = (xyzi[:1].copy(), raw[:1], 4, 8, (-0.5, 0.5), False)
    out = {}
    for name, (a, b, H, W, tr, lf) in small.items():
        res = run(name, a, b, H, W, tr, lf)
        out[name + "/xyzi"] = a
        out[name + "/raw"] = b
        out[name + "/hw"] = np.array([H, W])
        out[name + "/theta_range"] = np.array(tr if tr is not None else [np.nan, np.nan])
        out[name + "/largest_first"] = np.array(int(lf))
        for k, v in res.items():
            if k != "alpha":
                out[name + "/" + k] = v
        out[name + "/alpha_sha"] = np.frombuffer(bytes.fromhex(sha(res["alpha"])), dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLD, "projection_small.npz"), **out)

    # full-size cases: digests only (inputs regenerate from the seed)
    for name, sensor, seed, tr in (("hdl64_seed0", "hdl64", 0, None), ("hdl64_seed1", "hdl64", 1, None),
                                   ("os1_128_seed0", "os1-128", 0, None),
                                   ("os1_128_cudal", "os1-128", 2, (-np.pi / 8, np.pi / 8))):
        xyzi, raw = synth.synth_scan(seed, sensor)
        H, W = synth.SENSORS[sensor][4:6]
        res = run(name, xyzi, raw, H, W, tr)
        cases[name] = {
            "sensor": sensor, "seed": seed, "H": H, "W": W,
            "theta_range": None if tr is None else list(tr),
            "n_points": int(xyzi.shape[0]), "xyzi_sha": sha(xyzi), "raw_sha": sha(raw),
            "img_sha": sha(res["img"]), "winner_sha": sha(res["winner"]), "pix_sha": sha(res["pix"]),
            "theta_min": float(res["theta"][0]), "theta_max": float(res["theta"][1]),
            "occupied": int((res["winner"] >= 0).sum()),
        }
    return cases


def gen_kitti_loader():
    """SemanticKitti.__getitem__ (resize off) on a temp .bin/.label pair."""
    from dataset.dataloader_semantic_KITTI import SemanticKitti

    xyzi, raw = synth.synth_scan(21, "tiny")
    with tempfile.TemporaryDirectory() as d:
        fb, fl = os.path.join(d, "000000.bin"), os.path.join(d, "000000.label")
        xyzi.tofile(fb)
        raw.tofile(fl)
        ds = SemanticKitti([(fb, fl)], rotate=False, flip=False, projection=(16, 256), resize=False)
        rng_img, refl, xyz, normals, sem = ds[0]
    np.savez_compressed(os.path.join(GOLD, "kitti_loader.npz"), xyzi=xyzi, raw=raw, hw=np.array([16, 256]),
                        range=rng_img.numpy(), reflectivity=refl.numpy(), xyz=xyz.numpy(),
                        normals=normals.numpy(), semantics=sem.numpy())


def gen_kitti_loader_aug():
    """SemanticKitti.__getitem__ with resize / flip / rotate; np.random is seeded so the draws are known."""
    from dataset.dataloader_semantic_KITTI import SemanticKitti

    xyzi, raw = synth.synth_scan(22, "tiny")
    out = {"xyzi": xyzi, "raw": raw}
    with tempfile.TemporaryDirectory() as d:
        fb, fl = os.path.join(d, "000000.bin"), os.path.join(d, "000000.label")
        xyzi.tofile(fb)
        raw.tofile(fl)
        # case A: resize to 128x2048 (hard-coded in the reference), no augmentation
        ds = SemanticKitti([(fb, fl)], rotate=False, flip=False, projection=(16, 256), resize=True)
        r = [t.numpy() for t in ds[0]]
        for k, a in zip(("range", "reflectivity", "xyz", "normals", "semantics"), r):
            out["resize/" + k + "_sha"] = np.frombuffer(bytes.fromhex(sha(a)), dtype=np.uint8)
        out["resize/normals_sub"] = r[3][:, ::4, ::16].copy()
        out["resize/xyz_sub"] = r[2][:, ::4, ::16].copy()
        # case B: rotate + flip at native resolution; find a seed whose coin flips heads
        seed = 0
        while True:
            np.random.seed(seed)
            angle = float(np.random.randint(-180, 180))
            if np.random.rand() < 0.5:
                break
            seed += 1
        np.random.seed(seed)
        ds = SemanticKitti([(fb, fl)], rotate=True, flip=True, projection=(16, 256), resize=False)
        r = [t.numpy() for t in ds[0]]
        out["aug/angle"] = np.array(angle)
        for k, a in zip(("range", "reflectivity", "xyz", "normals", "semantics"), r):
            out["aug/" + k] = a
    np.savez_compressed(os.path.join(GOLD, "kitti_loader_aug.npz"), **out)


def gen_other_loaders():
    """SemanticCUDAL (projected, +-pi/8) and SemanticTHAB (organised 128x2048) items from the reference."""
    from dataset.dataloader_semantic_CUDAL import SemanticCUDAL
    from dataset.dataloader_semantic_THAB import SemanticTHAB

    def subs(r, tag, out):
        for k, a in zip(("range", "reflectivity", "xyz", "normals", "semantics"), r):
            out[f"{tag}/{k}_sha"] = np.frombuffer(bytes.fromhex(sha(a)), dtype=np.uint8)
            out[f"{tag}/{k}_sub"] = a[:, ::4, ::16].copy()

    out = {}
    xyzi, raw = synth.synth_scan(23, "tiny")
    raw = raw.copy()
    raw[:50] = (raw[:50] & np.uint32(0xFFFF0000)) | np.uint32(2)          # CUDAL-only id 2 -> 12
    xyzi = xyzi.copy()
    xyzi[:, 3] *= 3.0                                                      # reflectivity above 1: the max-normalisation matters
    out["cudal/xyzi"], out["cudal/raw"] = xyzi, raw
    with tempfile.TemporaryDirectory() as d:
        fb, fl = os.path.join(d, "0.bin"), os.path.join(d, "0.label")
        xyzi.tofile(fb); raw.tofile(fl)
        ds = SemanticCUDAL([(fb, fl)], rotate=False, flip=False, projection=(32, 256), resize=True)
        subs([t.numpy() for t in ds[0]], "cudal", out)
    # THAB: organised OS1-128 cloud; seed so that flip = True, then the angle is drawn
    xyzi, raw = synth.synth_scan(24, "os1-128")
    out["thab/seed"] = np.array(24)
    with tempfile.TemporaryDirectory() as d:
        fb, fl = os.path.join(d, "0.bin"), os.path.join(d, "0.label")
        xyzi.tofile(fb); raw.tofile(fl)
        subs([t.numpy() for t in SemanticTHAB([(fb, fl)])[0]], "thab_plain", out)
        seed = 0
        while True:
            np.random.seed(seed)
            if np.random.choice([True, False]):
                angle = int(np.random.randint(-180, 180))
                break
            seed += 1
        np.random.seed(seed)
        subs([t.numpy() for t in SemanticTHAB([(fb, fl)], rotate=True, flip=True)[0]], "thab_aug", out)
        out["thab_aug/angle"] = np.array(angle)
        out["thab_aug/np_seed"] = np.array(seed)
    # WADS: +-pi/2 elevation range, empty rows dropped, 64x1024 target; STF: 5-column file, clip, identity labels
    from dataset.dataloader_semantic_STF import SemanticSTF
    from dataset.dataloader_semantic_WADS import SemanticWADS
    xyzi, raw = synth.synth_scan(25, "tiny")
    raw = raw.copy()
    raw[:40] = (raw[:40] & np.uint32(0xFFFF0000)) | np.uint32(110)
    out["wads/xyzi"], out["wads/raw"] = xyzi, raw
    with tempfile.TemporaryDirectory() as d:
        fb, fl = os.path.join(d, "0.bin"), os.path.join(d, "0.label")
        xyzi.tofile(fb); raw.tofile(fl)
        subs([t.numpy() for t in SemanticWADS([(fb, fl)], projection=(64, 256), resize=True)[0]], "wads", out)
        r = [t.numpy() for t in SemanticWADS([(fb, fl)], projection=(64, 256), resize=False)[0]]
        out["wads_native/shape"] = np.array(r[2].shape)
        for k, a in zip(("range", "reflectivity", "xyz", "normals", "semantics"), r):
            out[f"wads_native/{k}"] = a
    rng = np.random.default_rng(26)
    xyzi, _ = synth.synth_scan(26, "tiny")
    xyzi = xyzi.copy()
    xyzi[:300, :3] *= 0.02                                                   # inside the 1.8 m sensor clip
    five = np.concatenate([xyzi[:, :3], xyzi[:, 3:4] * 255.0, rng.random((xyzi.shape[0], 1))], axis=1).astype(np.float32)
    lab = rng.integers(0, 22, xyzi.shape[0]).astype(np.uint32)
    out["stf/five"], out["stf/label"] = five, lab
    with tempfile.TemporaryDirectory() as d:
        fb, fl = os.path.join(d, "0.bin"), os.path.join(d, "0.label")
        five.tofile(fb); lab.tofile(fl)
        subs([t.numpy() for t in SemanticSTF([(fb, fl)], projection=(16, 256), resize=True, remap_adverse_label=True, clip=True)[0]], "stf", out)
    np.savez_compressed(os.path.join(GOLD, "other_loaders.npz"), **out)


# ---------------------------------------------------------------- uncertainty
def gen_mc():
    pe, mi = extract_closures(os.path.join(_refshim.REF_SRC, "models", "tester.py"),
                              ["mc_predictive_entropy_norm", "mc_mutual_information_norm"])
    from utils.mc_dropout import predictive_entropy_mc
    import torch.nn.functional as F

    out = {}
    for name, (T, B, C, H, W, scale) in {"mc_small": (5, 2, 20, 4, 64, 3.0),
                                         "mc_peaked": (3, 1, 20, 2, 64, 25.0),
                                         "mc_c7": (4, 1, 7, 3, 40, 2.0)}.items():
        logits, labels = synth.synth_mc_logits(101, T, B, C, H, W, scale=scale)
        with torch.no_grad():
            probs = F.softmax(logits, dim=2)                 # tester.py:412
            p_bar = probs.mean(dim=0)                        # :415
            preds = p_bar.argmax(dim=1)                      # :417
            H_norm = pe(probs)
            MI_norm = mi(probs)
            H_mc = predictive_entropy_mc(probs)
        out.update({name + "/logits": logits.numpy(), name + "/labels": labels.numpy(),
                    name + "/p_bar": p_bar.numpy(), name + "/pred": preds.numpy(),
                    name + "/H_norm": H_norm.numpy(), name + "/MI_norm": MI_norm.numpy(),
                    name + "/H_mc": H_mc.numpy()})
    np.savez_compressed(os.path.join(GOLD, "mc_reduce.npz"), **out)


def gen_evidential():
    import models.probability_helper as ph
    from metrics.auroc import AUROCAggregator

    out = {}
    for name, (B, C, H, W, scale) in {"ev_small": (2, 20, 4, 64, 3.0), "ev_strong": (1, 20, 2, 64, 12.0)}.items():
        o, labels = synth.synth_evidential_logits(202, B, C, H, W, scale=scale)
        with torch.no_grad():
            alpha = ph.to_alpha_concentrations_from_shape_and_scale(o[:, :C], o[:, C:C + 1])
            Hn = ph.get_predictive_entropy_norm(alpha)
            Hp = ph.get_predictive_entropy(alpha)
            AU = ph.get_aleatoric_uncertainty(alpha)
            EU = ph.get_epistemic_uncertainty(alpha)
            MI = AUROCAggregator(mode="alpha", score="mi_norm")._uncertainty_score(alpha)
            pred = torch.softmax(o[:, :C], dim=1).argmax(dim=1)            # tester.py:493-495
        out.update({name + "/outputs": o.numpy(), name + "/labels": labels.numpy(), name + "/alpha": alpha.numpy(),
                    name + "/H_norm": Hn.numpy(), name + "/H": Hp.numpy(), name + "/AU": AU.numpy(),
                    name + "/EU": EU.numpy(), name + "/MI_norm": MI.numpy(), name + "/pred": pred.numpy()})
    np.savez_compressed(os.path.join(GOLD, "evidential.npz"), **out)


# ---------------------------------------------------------------- metrics
def gen_metrics():
    from models.evaluator import IoUEvaluator
    from metrics.ece import ECEAggregator

    out = {}
    g = torch.Generator().manual_seed(303)
    C = 20
    preds = torch.randint(-1, C + 1, (3, 8, 64), generator=g)     # includes out-of-range -1 and C
    targets = torch.randint(-1, C + 1, (3, 8, 64), generator=g)
    ev = IoUEvaluator(C)
    ev.update(preds[:2], targets[:2])
    ev.update(preds[2:], targets[2:])
    names = {i: str(i) for i in range(C)}
    miou, per = ev.compute(names, test_mask=[0] + [1] * (C - 1), ignore_gt=[0])
    miou_all, per_all = ev.compute(names)
    out.update({"iou/preds": preds.numpy(), "iou/targets": targets.numpy(), "iou/confmat": ev.confmat.numpy(),
                "iou/miou": np.array(miou), "iou/per_class": np.array([per[str(i)] for i in range(C)]),
                "iou/miou_all": np.array(miou_all), "iou/per_class_all": np.array([per_all[str(i)] for i in range(C)])})

    for mode in ("alpha", "logits", "probs"):
        x = torch.randn((2, C, 8, 64), generator=g) * 3.0
        if mode == "alpha":
            x = torch.nn.functional.softplus(x) + 1.0
        elif mode == "probs":
            x = torch.softmax(x, dim=1)
        labels = torch.randint(0, C, (2, 8, 64), generator=g)
        # make about half the predictions correct so acc is not ~1/C
        with torch.no_grad():
            am = x.argmax(1)
            take = torch.rand((2, 8, 64), generator=g) < 0.5
            labels = torch.where(take, am, labels)
        agg = ECEAggregator(n_bins=15, mode=mode, ignore_index=0, max_samples=None)
        agg.update(x[:1], labels[:1])
        agg.update(x[1:], labels[1:])
        (ece, mce), stats, _ = agg.compute(save_plot_path=os.path.join(tempfile.gettempdir(), "slu_gold.png"))
        out.update({f"ece_{mode}/preds": x.numpy(), f"ece_{mode}/labels": labels.numpy(),
                    f"ece_{mode}/conf": agg._conf.numpy(), f"ece_{mode}/correct": agg._correct.numpy(),
                    f"ece_{mode}/n": stats["n"].to_numpy(), f"ece_{mode}/acc": stats["acc"].to_numpy(),
                    f"ece_{mode}/avg_conf": stats["conf"].to_numpy(),
                    f"ece_{mode}/ece_mce": np.array([ece, mce])})
    # AUROC (error detection) and accuracy-vs-uncertainty bins, every pixel kept (max_samples=None)
    from metrics.auroc import AUROCAggregator
    from models.evaluator import UncertaintyAccuracyAggregator
    x = torch.softmax(torch.randn((2, C, 16, 128), generator=g) * 2.0, dim=1)
    labels = torch.randint(0, C, (2, 16, 128), generator=g)
    with torch.no_grad():
        labels = torch.where(torch.rand((2, 16, 128), generator=g) < 0.6, x.argmax(1), labels)
    out["auroc/probs"] = x.numpy(); out["auroc/labels"] = labels.numpy()
    for score in ("entropy_norm", "entropy", "1-maxprob"):
        a = AUROCAggregator(mode="probs", score=score, ignore_index=0, max_samples=None)
        a.update(x[:1], labels[:1]); a.update(x[1:], labels[1:])
        out[f"auroc/probs_{score}"] = np.array(a.compute(save_plot_path=os.path.join(tempfile.gettempdir(), "slu_gold_roc.png"))[0])
    override = torch.rand((2, 16, 128), generator=g)
    a = AUROCAggregator(mode="probs", score="mi_norm", ignore_index=0, max_samples=None)
    a.update(x, labels, score_override=override)
    out["auroc/override"] = override.numpy(); out["auroc/probs_override"] = np.array(a.compute(save_plot_path=os.path.join(tempfile.gettempdir(), "slu_gold_roc.png"))[0])
    alpha = torch.nn.functional.softplus(torch.randn((2, C, 16, 128), generator=g) * 3.0) + 1.0
    out["auroc/alpha"] = alpha.numpy()
    for score in ("mi_norm", "entropy_norm"):
        a = AUROCAggregator(mode="alpha", score=score, ignore_index=0, max_samples=None)
        a.update(alpha, labels)
        out[f"auroc/alpha_{score}"] = np.array(a.compute(save_plot_path=os.path.join(tempfile.gettempdir(), "slu_gold_roc.png"))[0])
    ua = UncertaintyAccuracyAggregator(max_samples=None)
    ua.update(labels=labels, preds=x.argmax(1), uncertainty=override, ignore_ids=(0,))
    for nb in (10, 20):
        df = ua.binned_accuracy(num_bins=nb)
        out[f"ua/n_{nb}"] = df["n"].to_numpy(); out[f"ua/acc_{nb}"] = df["accuracy"].to_numpy()

    # empty aggregator: 2-tuple with NaNs (ece.py:157-158)
    r = ECEAggregator(n_bins=15, mode="probs", ignore_index=0).compute(save_plot_path=None)
    out["ece_empty/len"] = np.array(len(r))
    np.savez_compressed(os.path.join(GOLD, "metrics.npz"), **out)


# ---------------------------------------------------------------- losses
def gen_losses():
    from losses.dirichlet_losses import BrierDirichlet, DigammaDirichletCE, DirichletMSELoss, NLLDirichletCategorical
    from losses.regularizers import KL_offClasses_to_uniform

    g = torch.Generator().manual_seed(404)
    B, C, H, W = 2, 20, 4, 32
    alpha = (torch.nn.functional.softplus(torch.randn((B, C, H, W), generator=g) * 3.0) + 1.0)
    target = torch.randint(0, C, (B, H, W), generator=g)
    out = {"alpha": alpha.numpy(), "target": target.numpy()}
    for name, mod in (("mse", DirichletMSELoss(ignore_index=0)), ("kl", KL_offClasses_to_uniform(ignore_index=0)),
                      ("nll", NLLDirichletCategorical(ignore_index=0)), ("dce", DigammaDirichletCE(ignore_index=0)),
                      ("brier", BrierDirichlet(ignore_index=0)), ("brier_sref", BrierDirichlet(ignore_index=0, s_ref=30.0))):
        a = alpha.clone().requires_grad_(True)
        loss = mod(a, target)
        (grad,) = torch.autograd.grad(loss, a)
        out[name + "/loss"] = loss.detach().numpy()
        out[name + "/grad"] = grad.numpy()
    np.savez_compressed(os.path.join(GOLD, "losses.npz"), **out)


def gen_loss_terms():
    """The remaining terms, run through the unmodified reference modules + autograd."""
    from losses.dirichlet_losses import ComplementKLUniform
    from losses.regularizers import (EvidenceReg, EvidenceRegBand, KL_offClasses_to_uniform, LogitRegularizer,
                                     WrongLowEvidence)

    g = torch.Generator().manual_seed(505)
    B, C, H, W = 2, 20, 4, 32
    alpha = (torch.nn.functional.softplus(torch.randn((B, C, H, W), generator=g) * 3.0) + 1.0)
    target = torch.randint(0, C, (B, H, W), generator=g)
    # make ~half of the pixels "correct" so WrongLowEvidence / the complement gate see both kinds
    fix = torch.rand((B, H, W), generator=g) < 0.5
    target = torch.where(fix, alpha.argmax(dim=1), target)
    logits = torch.randn((B, C + 1, H, W), generator=g) * 4.0
    keep = torch.rand((B, H, W), generator=g) > 0.3
    out = {"alpha": alpha.numpy(), "target": target.numpy(), "logits": logits.numpy(), "keep": keep.numpy()}

    def run(name, fn, x):
        x = x.clone().requires_grad_(True)
        loss = fn(x)
        (grad,) = torch.autograd.grad(loss, x)
        out[name + "/loss"] = loss.detach().numpy()
        out[name + "/grad"] = grad.numpy()

    run("comp", lambda a: ComplementKLUniform(ignore_index=0)(a, target), alpha)
    run("comp_trainer", lambda a: ComplementKLUniform(ignore_index=0, gamma=1.25, tau=0.65, sigma=0.15)(a, target), alpha)   # trainer.py:339
    run("comp_evid_gate", lambda a: ComplementKLUniform(ignore_index=0, s_target=40.0, normalize=False)(a, target), alpha)
    run("comp_attached", lambda a: ComplementKLUniform(ignore_index=0, detach_uncert=False)(a, target), alpha)
    run("wle", lambda a: WrongLowEvidence(ignore_index=0, s_low=0.0, margin=0.05, soft_margin_k=0.08)(a, target), alpha)  # trainer.py:376-381
    run("wle_hard", lambda a: WrongLowEvidence(ignore_index=0, s_low=2.0, margin=0.1, soft_margin_k=0.0)(a, target), alpha)
    run("wle_nomargin", lambda a: WrongLowEvidence(ignore_index=None, margin=0.0)(a, target), alpha)
    run("band", lambda a: EvidenceRegBand(60.0, band=0.10, ignore_index=0)(a, target=target), alpha)
    run("band_mask", lambda a: EvidenceRegBand(200.0, band=0.25)(a, mask=keep), alpha)
    run("band_nomask", lambda a: EvidenceRegBand(30.0)(a), alpha)
    run("ereg_log", lambda a: EvidenceReg(60.0, ignore_index=0)(a, target=target), alpha)
    run("ereg_log_sc", lambda a: EvidenceReg(60.0, scale_correct=True)(a, mask=keep), alpha)
    run("ereg_one_sided", lambda a: EvidenceReg(60.0, mode="one_sided", margin=0.2, ignore_index=(0, 3))(a, target=target), alpha)
    run("ereg_l2", lambda a: EvidenceReg(60.0, mode="l2")(a), alpha)
    run("klw", lambda a: KL_offClasses_to_uniform(ignore_index=0, with_conf_weighting=True, gamma=1.0)(a, target), alpha)
    run("klw_g2", lambda a: KL_offClasses_to_uniform(ignore_index=0, with_conf_weighting=True, gamma=2.0)(a, target), alpha)
    run("logit", lambda z: LogitRegularizer()(z), logits)
    run("logit_thr_target", lambda z: LogitRegularizer(threshold=3.0, ignore_index=0)(z, target=target), logits)
    run("logit_mask", lambda z: LogitRegularizer(threshold=None)(z, mask=keep), logits)
    run("logit_target_noignore", lambda z: LogitRegularizer(threshold=1.0)(z, target=target), logits)
    np.savez_compressed(os.path.join(GOLD, "loss_terms.npz"), **out)


def gen_aurc():
    """Risk-coverage curves from the reference's metrics/aurc.py on continuous, tied and degenerate inputs."""
    from metrics.aurc import UncertaintyAggregator, aurc_from_risks_confids, rc_curve_stats
    rng = np.random.default_rng(606)
    out = {}
    for name, n, quant in (("cont", 4000, None), ("ties", 3000, 0.01), ("two", 2, None), ("one", 1, None), ("allwrong", 50, None)):
        conf = rng.random(n).astype(np.float32)
        if quant:
            conf = (np.round(conf / quant) * quant).astype(np.float32)
        risks = (rng.random(n) < (1.0 - conf) * 0.8).astype(np.float32) if name != "allwrong" else np.ones(n, np.float32)
        cov, sel, w = rc_curve_stats(risks, conf)
        aurc, eaurc, _, _ = aurc_from_risks_confids(risks, conf)
        out.update({f"{name}/conf": conf, f"{name}/risks": risks, f"{name}/cov": cov, f"{name}/sel": sel, f"{name}/w": w,
                    f"{name}/aurc": np.array(aurc), f"{name}/eaurc": np.array(eaurc)})
    # the aggregator end to end: softmax probs [B,C,H,W] + labels [B,1,H,W], both confidence definitions
    g = torch.Generator().manual_seed(607)
    probs = torch.softmax(torch.randn((3, 20, 16, 128), generator=g) * 2.0, dim=1)
    lab = torch.randint(0, 20, (3, 1, 16, 128), generator=g)
    lab = torch.where(torch.rand(lab.shape, generator=g) < 0.6, probs.argmax(dim=1, keepdim=True), lab)
    out["agg/probs"], out["agg/labels"] = probs.numpy(), lab.numpy()
    for tag, mp in (("entropy", False), ("maxprob", True)):
        agg = UncertaintyAggregator(ignore_index=0, use_max_prob_confidence=mp)
        agg.add_batch(probs[:2], lab[:2])
        agg.add_batch(probs[2:], lab[2:])
        r = agg.finalize(make_plots=False)
        out[f"agg/{tag}"] = np.array([r["AURC"], r["EAURC"], r["num_pixels"]], dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, "aurc.npz"), **out)


def gen_per_class():
    """UncertaintyPerClassAggregator (src/models/evaluator.py:191-281): per-class sample counts, means and quartiles of
    the reference's per-pixel arrays, and the class order plot_iou_sorted_by_uncertainty derives from them (:559-566)."""
    from models.evaluator import UncertaintyPerClassAggregator
    g = torch.Generator().manual_seed(808)
    C = 20
    labels = torch.randint(-1, C + 1, (4, 16, 256), generator=g)              # includes ids outside [0, C)
    base = torch.rand((4, 16, 256), generator=g)
    unc = (base ** (1.0 + 0.15 * labels.clamp(0, C - 1).float())).float()    # class-dependent distributions
    agg = UncertaintyPerClassAggregator(C)
    agg.update(labels[:1], unc[:1])
    agg.update(labels[1:], unc[1:])
    vals = [agg._values[c].double().numpy() for c in range(C)]
    stats = np.array([[v.size, v.mean(), np.quantile(v, 0.25), np.median(v), np.quantile(v, 0.75)] for v in vals])
    df = agg.as_dataframe([str(i) for i in range(C)], ignore_ids=(0,))
    order = df.groupby("class_id")["uncertainty"].mean().sort_values().index.to_numpy()
    np.savez_compressed(os.path.join(GOLD, "per_class.npz"), labels=labels.numpy(), unc=unc.numpy(), stats=stats,
                        seen=np.array(agg._seen_counts), order_by_mean=order, df_rows=np.array(len(df)))


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(1)
    if len(sys.argv) > 1:                      # regenerate selected files only: python oracle/gen_golden.py loss_terms
        for name in sys.argv[1:]:
            globals()["gen_" + name]()
        return
    manifest = {
        "generator": "oracle/gen_golden.py",
        "reference_root": _refshim.REF_ROOT,
        "versions": {"python": sys.version.split()[0], "numpy": np.__version__, "torch": torch.__version__},
        "projection_full": gen_projection(),
    }
    gen_kitti_loader()
    gen_kitti_loader_aug()
    gen_other_loaders()
    gen_mc()
    gen_evidential()
    gen_metrics()
    gen_losses()
    gen_loss_terms()
    gen_aurc()
    gen_per_class()
    with open(os.path.join(GOLD, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    for fn in sorted(os.listdir(GOLD)):
        print(fn, os.path.getsize(os.path.join(GOLD, fn)))


if __name__ == "__main__":
    main()
