"""Standalone confusion / reliability histograms (slu_confusion_ece) and the IoUEvaluator /
ECEAggregator classes: integer counts bit-exact against the reference's golden vectors."""
import numpy as np
import pytest
import torch

from oracle import metrics as om
from semanticlidarunc_b200 import ops
from semanticlidarunc_b200.metrics.ece import ECEAggregator
from semanticlidarunc_b200.models.evaluator import IoUEvaluator

pytestmark = pytest.mark.gpu


def test_iou_evaluator_vs_reference_golden(cuda, golden):
    g = golden("metrics.npz")
    preds, targets = torch.from_numpy(g["iou/preds"]), torch.from_numpy(g["iou/targets"])
    C = 20
    ev = IoUEvaluator(C)
    ev.update(preds[:2], targets[:2])            # CPU tensors in, as the reference's callers may pass
    ev.update(preds[2:].to(cuda), targets[2:].to(cuda))
    assert ev.confmat.device.type == "cpu" and ev.confmat.dtype == torch.long
    assert np.array_equal(ev.confmat.numpy(), g["iou/confmat"])
    names = {i: str(i) for i in range(C)}
    miou, per = ev.compute(names, test_mask=[0] + [1] * (C - 1), ignore_gt=[0])
    assert miou == float(g["iou/miou"])
    assert np.array_equal(np.array([per[str(i)] for i in range(C)]), g["iou/per_class"], equal_nan=True)
    miou_all, per_all = ev.compute(names)
    assert miou_all == float(g["iou/miou_all"])
    ev.reset()
    assert int(ev.confmat.sum()) == 0


@pytest.mark.parametrize("mode", ["alpha", "logits", "probs"])
def test_ece_aggregator_vs_reference_golden(cuda, golden, mode):
    g = golden("metrics.npz")
    x, lab = torch.from_numpy(g[f"ece_{mode}/preds"]), torch.from_numpy(g[f"ece_{mode}/labels"])
    agg = ECEAggregator(n_bins=15, mode=mode, ignore_index=0, max_samples=None)
    agg.update(x[:1], lab[:1])
    agg.update(x[1:].to(cuda), lab[1:].to(cuda))
    (ece, mce), stats, fig = agg.compute(save_plot_path=None)
    # exact integer state vs the reference's stored samples
    n, nc, cs = om.ece_bin_counts(g[f"ece_{mode}/conf"], g[f"ece_{mode}/correct"], 15)
    assert np.array_equal(stats["n"].to_numpy(), g[f"ece_{mode}/n"])
    assert np.array_equal(agg._bins[0].cpu().numpy(), n) and np.array_equal(agg._bins[1].cpu().numpy(), nc)
    assert agg._seen == int(n.sum())
    # ECE / MCE within 1e-5 relative of the reference's own numbers
    ece_ref, mce_ref = g[f"ece_{mode}/ece_mce"]
    assert abs(ece - ece_ref) <= 1e-5 * abs(ece_ref) + 1e-9
    assert abs(mce - mce_ref) <= 1e-5 * abs(mce_ref) + 1e-9
    assert np.allclose(stats["acc"].to_numpy(), g[f"ece_{mode}/acc"], rtol=1e-6, equal_nan=True)
    assert np.allclose(stats["conf"].to_numpy(), g[f"ece_{mode}/avg_conf"], rtol=1e-5, equal_nan=True)


def test_ece_empty_returns_two_tuple(cuda):
    r = ECEAggregator(n_bins=15, mode="probs", ignore_index=0).compute()
    assert len(r) == 2 and np.isnan(r[0][0])


@pytest.mark.parametrize("C,n", [(20, 1_000_003), (3, 257), (100, 50_000)])
def test_standalone_histograms_vs_oracle(cuda, C, n):
    g = torch.Generator().manual_seed(C + n)
    pred = torch.randint(-1, C + 1, (n,), generator=g)
    lab = torch.randint(-1, C + 1, (n,), generator=g)
    conf = torch.rand((n,), generator=g)
    conf[:5] = torch.tensor([0.0, 1.0, 1.0 / 15, float(np.float32(14 / 15)), 0.5])   # on-edge values
    confmat = torch.zeros((C, C), dtype=torch.int64, device=cuda)
    bins = ops.new_ece_bins(15, cuda)
    ops.confusion_ece(pred.to(cuda), lab.to(cuda), conf.to(cuda), num_classes=C, ignore_index=0, confmat=confmat, ece_bins=bins)
    assert torch.equal(confmat.cpu(), om.confusion_counts(pred, lab, C))
    valid = lab != 0
    nn, nc, cs = om.ece_bin_counts(conf[valid].numpy(), (pred[valid] == lab[valid]).numpy(), 15)
    assert np.array_equal(bins[0].cpu().numpy(), nn) and np.array_equal(bins[1].cpu().numpy(), nc)
    assert np.allclose(bins[2].cpu().numpy() / 2.0 ** 32, cs, rtol=1e-9, atol=1e-6)
    # np.histogram itself (the reference's binning) agrees on the counts
    assert np.array_equal(np.histogram(conf[valid].numpy(), bins=om.ece_edges(15))[0], nn)


def test_coherent_labels_same_counts(cuda):
    """Warp-aggregated path with long runs of equal keys (what real label maps look like)."""
    from semanticlidarunc_b200 import synth
    C = 20
    lab = synth.synth_coherent_labels(1, 2, C, 64, 2048)
    pred = synth.synth_coherent_labels(2, 2, C, 64, 2048, block=16)
    confmat = ops.new_confmat(C, cuda)
    ops.confusion_ece(pred.to(cuda), lab.to(cuda), None, num_classes=C, confmat=confmat)
    assert torch.equal(confmat.cpu(), om.confusion_counts(pred, lab, C))
    assert int(confmat.sum()) == lab.numel()


def test_auroc_aggregator_vs_reference_golden(cuda, golden):
    """Histogram AUROC (60000 score bins) vs the reference's sort-based value on every pixel: 1e-4."""
    from semanticlidarunc_b200.metrics.auroc import AUROCAggregator
    g = golden("metrics.npz")
    x, lab = torch.from_numpy(g["auroc/probs"]), torch.from_numpy(g["auroc/labels"])
    for score in ("entropy_norm", "entropy", "1-maxprob"):
        a = AUROCAggregator(mode="probs", score=score, ignore_index=0, max_samples=None)
        a.update(x[:1], lab[:1])
        a.update(x[1:].to(cuda), lab[1:].to(cuda))
        auroc, curves, fig = a.compute()
        ref = float(g[f"auroc/probs_{score}"])
        assert abs(auroc - ref) < 1e-4, (score, auroc, ref)
        assert curves["fpr"][0] == 0.0 and curves["tpr"][-1] == 1.0
    a = AUROCAggregator(mode="probs", score="mi_norm", ignore_index=0)
    a.update(x.to(cuda), lab.to(cuda), score_override=torch.from_numpy(g["auroc/override"]).to(cuda))
    assert abs(a.compute()[0] - float(g["auroc/probs_override"])) < 1e-4
    alpha = torch.from_numpy(g["auroc/alpha"]).to(cuda)
    for score in ("mi_norm", "entropy_norm"):
        a = AUROCAggregator(mode="alpha", score=score, ignore_index=0)
        a.update(alpha, lab.to(cuda))
        assert abs(a.compute()[0] - float(g[f"auroc/alpha_{score}"])) < 1e-4, score
    assert a._seen == int((lab != 0).sum())
    e = AUROCAggregator(mode="probs")
    assert np.isnan(e.compute()[0])


def test_uncertainty_accuracy_aggregator_vs_reference_golden(cuda, golden):
    from semanticlidarunc_b200.models.evaluator import UncertaintyAccuracyAggregator
    g = golden("metrics.npz")
    x, lab = torch.from_numpy(g["auroc/probs"]), torch.from_numpy(g["auroc/labels"])
    ua = UncertaintyAccuracyAggregator(max_samples=None)
    ua.update(labels=lab.to(cuda), preds=x.argmax(1).to(cuda), uncertainty=torch.from_numpy(g["auroc/override"]).to(cuda), ignore_ids=(0,))
    for nb in (10, 20):
        df = ua.binned_accuracy(num_bins=nb)
        assert np.array_equal(df["n"].to_numpy(), g[f"ua/n_{nb}"])                 # 10 and 20 divide 60000: exact
        assert np.allclose(df["accuracy"].to_numpy(), g[f"ua/acc_{nb}"], rtol=1e-6, equal_nan=True)
    assert list(df.columns) == ["low", "high", "label", "n", "pct", "accuracy"]


def test_whole_mc_block_updates_every_aggregator(cuda):
    """mc_reduce_from_logits == the reference tester's MC block: IoU, ECE, both AUROCs and the UA bins."""
    from oracle import uncertainty as ou
    from semanticlidarunc_b200 import synth
    from semanticlidarunc_b200.metrics.auroc import AUROCAggregator
    from semanticlidarunc_b200.models.evaluator import UncertaintyAccuracyAggregator
    from semanticlidarunc_b200.utils.mc_dropout import mc_reduce_from_logits
    C = 20
    x, lab = synth.synth_mc_logits(21, 6, 2, C, 16, 256)
    iou, ece = IoUEvaluator(C), ECEAggregator(n_bins=15, mode="probs", ignore_index=0)
    au, au_mi = AUROCAggregator(mode="probs", score="entropy_norm", ignore_index=0), AUROCAggregator(mode="probs", score="mi_norm", ignore_index=0)
    ua = UncertaintyAccuracyAggregator()
    out = mc_reduce_from_logits(x.to(cuda), lab.to(cuda), iou_evaluator=iou, ece_eval=ece, auroc_eval=au, auroc_eval_mi=au_mi, ua_agg=ua)
    ref = ou.mc_reduce(x)
    valid = (lab != 0).numpy().reshape(-1)
    err = (ref["pred"] != lab).numpy().reshape(-1)[valid]
    assert abs(au.compute()[0] - om.auroc_error_detection(ref["H_norm"].numpy().reshape(-1)[valid], err)) < 2e-4
    assert abs(au_mi.compute()[0] - om.auroc_error_detection(ref["MI_norm"].numpy().reshape(-1)[valid], err)) < 2e-4
    n, acc = om.binned_accuracy(ref["H_norm"].numpy().reshape(-1)[valid].clip(0, 1), ~err, 10)
    df = ua.binned_accuracy(num_bins=10)
    assert abs(df["n"].to_numpy() - n).sum() <= 4          # a pixel within 1e-6 of a coarse edge may move
    assert int(iou.confmat.sum()) == lab.numel() and ece._seen == int(valid.sum())


def test_uncertainty_per_class_aggregator(cuda):
    """Per-class mean is exact (what the reference's plot_iou_sorted_by_uncertainty consumes), quartiles and the
    expanded dataframe agree with the reference's per-pixel arrays to one histogram bin (1/2048)."""
    from semanticlidarunc_b200.models.evaluator import UncertaintyPerClassAggregator
    g = torch.Generator().manual_seed(3)
    C = 20
    labels = torch.randint(-1, C + 1, (3, 16, 256), generator=g)
    unc = torch.rand((3, 16, 256), generator=g) ** 2
    agg = UncertaintyPerClassAggregator(C)
    agg.update(labels[:1], unc[:1])
    agg.update(labels[1:].to(cuda), unc[1:].to(cuda))
    st = agg.class_stats().set_index("class_id")
    names = [str(i) for i in range(C)]
    df = agg.as_dataframe(names, ignore_ids=(0,))
    assert 0 not in set(df["class_id"]) and set(df["class_id"]) == set(range(1, C))
    for c in range(C):
        v = unc[labels == c].double().numpy()
        assert agg._seen_counts[c] == v.size == int(st.loc[c, "n"])
        assert abs(st.loc[c, "mean"] - v.mean()) < 1e-7
        assert abs(st.loc[c, "median"] - np.median(v)) < 0.01      # ~550 samples per class: neighbouring samples are ~0.004 apart
        if c:
            assert abs(df[df["class_id"] == c]["uncertainty"].mean() - v.mean()) < 1.0 / 2048
    agg.reset()
    assert sum(agg._seen_counts) == 0


def test_summary_cache_round_trip(cuda, tmp_path):
    from semanticlidarunc_b200 import synth
    from semanticlidarunc_b200.metrics.auroc import AUROCAggregator
    from semanticlidarunc_b200.models.evaluator import UncertaintyAccuracyAggregator, UncertaintyPerClassAggregator
    from semanticlidarunc_b200.models.summary_cache import load_summary, save_summary
    from semanticlidarunc_b200.utils.mc_dropout import mc_reduce_from_logits
    C = 20

    def fresh():
        return dict(iou_evaluator=IoUEvaluator(C), ece_eval=ECEAggregator(n_bins=15, mode="probs", ignore_index=0),
                    auroc_eval=AUROCAggregator(mode="probs", ignore_index=0), auroc_eval_mi=AUROCAggregator(mode="probs", score="mi_norm", ignore_index=0),
                    ua_agg=UncertaintyAccuracyAggregator(), unc_agg=UncertaintyPerClassAggregator(C))
    a = fresh()
    x, lab = synth.synth_mc_logits(31, 4, 1, C, 16, 256)
    out = mc_reduce_from_logits(x.to(cuda), lab.to(cuda), **{k: v for k, v in a.items() if k != "unc_agg"})
    a["unc_agg"].update(lab.to(cuda), out["H_norm"])
    path = str(tmp_path / "summary_epoch_000001.pt")
    save_summary(path, {"epoch_name": "000001", "num_frames": 1}, **a)
    b = fresh()
    cache = load_summary(path, **b)
    assert cache["meta"]["num_frames"] == 1
    assert torch.equal(a["iou_evaluator"].confmat, b["iou_evaluator"].confmat)
    assert a["ece_eval"].compute()[0] == b["ece_eval"].compute()[0]
    assert a["auroc_eval"].compute()[0] == b["auroc_eval"].compute()[0]
    assert a["auroc_eval_mi"].compute()[0] == b["auroc_eval_mi"].compute()[0]
    assert a["ua_agg"].binned_accuracy(10).equals(b["ua_agg"].binned_accuracy(10))
    assert a["unc_agg"].class_stats().equals(b["unc_agg"].class_stats())
    import pytest as _pt
    with _pt.raises(KeyError):
        from semanticlidarunc_b200.models.summary_cache import restore_summary
        restore_summary({"meta": {}}, iou_evaluator=IoUEvaluator(C))


@pytest.mark.parametrize("tag,maxprob", [("entropy", False), ("maxprob", True)])
def test_uncertainty_aggregator_aurc_vs_reference_golden(cuda, golden, tag, maxprob):
    """UncertaintyAggregator (src/metrics/aurc.py:210-350) on the device histogram: pixel count exact, AURC and
    E-AURC within 1e-5 of the reference's sort-based values."""
    from semanticlidarunc_b200.metrics.aurc import UncertaintyAggregator, compute_batch_uncertainty_metrics, aurc_from_risks_confids
    g = golden("aurc.npz")
    probs, lab = torch.from_numpy(g["agg/probs"]).to(cuda), torch.from_numpy(g["agg/labels"]).to(cuda)
    agg = UncertaintyAggregator(ignore_index=0, use_max_prob_confidence=maxprob)
    agg.add_batch(probs[:2], lab[:2])
    agg.add_batch(probs[2:].cpu(), lab[2:].cpu())                 # host tensors are accepted, as in the reference
    r = agg.finalize(make_plots=False)
    ref = g["agg/" + tag]
    assert r["num_pixels"] == int(ref[2])
    assert abs(r["AURC"] - ref[0]) <= 1e-5 and abs(r["EAURC"] - ref[1]) <= 1e-5, (r, ref)
    m = compute_batch_uncertainty_metrics(probs, lab, ignore_index=0, use_max_prob_confidence=maxprob)
    assert abs(m["AURC"] - ref[0]) <= 1e-5 and m["num_pixels"] == int(ref[2])
    assert m["recalls"].shape == (8,) and np.all(np.diff(m["recalls"]) >= 0) and m["recalls"][-1] <= 1.0
    a, e, _, _ = aurc_from_risks_confids(g["cont/risks"], g["cont/conf"])
    assert abs(a - float(g["cont/aurc"])) <= 1e-6 and abs(e - float(g["cont/eaurc"])) <= 1e-6
    agg.reset()
    with pytest.raises(RuntimeError):
        agg.finalize(make_plots=False)


@pytest.mark.parametrize("C,n,n_bins,offset", [(20, 1_000_003, 15, 0), (20, 262_144, 15, 0), (20, 4099, 15, 0), (64, 300_001, 32, 0),
                                                (20, 100_000, 64, 0), (100, 50_000, 15, 0), (20, 70_001, 15, 1), (7, 16_384, 10, 2)])
@pytest.mark.parametrize("coherent", [False, True])
def test_streaming_and_generic_histogram_kernels_agree(cuda, C, n, n_bins, offset, coherent):
    """slu_confusion_ece picks the streaming kernel (vector loads, private cells) for aligned inputs and the generic
    warp-aggregated one otherwise (offset != 0: views that are not 16-byte aligned; n_bins = 64: cells do not fit);
    integer counts must be bit-identical between the two and equal to the oracle's."""
    from semanticlidarunc_b200 import _lib
    g = torch.Generator().manual_seed(C * 7 + n)
    if coherent:
        lab = torch.randint(0, C, ((n + offset + 63) // 64,), generator=g).repeat_interleave(64)[: n + offset]
        pred = torch.where(torch.rand((n + offset,), generator=g) < 0.9, lab, torch.randint(0, C, (n + offset,), generator=g))
        conf = 1.0 - torch.rand((n + offset,), generator=g) * 0.05
    else:
        pred = torch.randint(-1, C + 1, (n + offset,), generator=g)
        lab = torch.randint(-1, C + 1, (n + offset,), generator=g)
        conf = torch.rand((n + offset,), generator=g)
        conf[offset:offset + 6] = torch.tensor([0.0, 1.0, 1.0 / n_bins, 0.5, float("nan"), 1.5])
    pd, ld, cd = pred.to(cuda)[offset:], lab.to(cuda)[offset:], conf.to(cuda)[offset:]
    out = []
    for generic in (0, 1):
        _lib.lib().slu_debug_hist_generic(generic)
        try:
            cm = torch.zeros((C, C), dtype=torch.int64, device=cuda)
            bins = ops.new_ece_bins(n_bins, cuda)
            ops.confusion_ece(pd, ld, cd, num_classes=C, ignore_index=0, confmat=cm, ece_bins=bins)
            ops.confusion_ece(pd, ld, cd, num_classes=C, ignore_index=0, confmat=cm, ece_bins=bins)      # accumulates
            out.append((cm.cpu(), bins.cpu()))
        finally:
            _lib.lib().slu_debug_hist_generic(0)
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
    p, l, c = pred[offset:], lab[offset:], conf[offset:]
    assert torch.equal(out[0][0], 2 * om.confusion_counts(p, l, C))
    valid = (l != 0) & ~torch.isnan(c)
    nn, nc, cs = om.ece_bin_counts(c[valid].clamp(0, 1).numpy(), (p[valid] == l[valid]).numpy(), n_bins)
    assert np.array_equal(out[0][1][0].numpy(), 2 * nn) and np.array_equal(out[0][1][1].numpy(), 2 * nc)
    # bins only / confusion only
    b2 = ops.new_ece_bins(n_bins, cuda)
    ops.confusion_ece(pd, ld, cd, num_classes=C, ignore_index=0, ece_bins=b2)
    cm2 = torch.zeros((C, C), dtype=torch.int64, device=cuda)
    ops.confusion_ece(pd, ld, None, num_classes=C, confmat=cm2)
    assert torch.equal(2 * b2.cpu(), out[0][1]) and torch.equal(2 * cm2.cpu(), out[0][0])


def test_uncertainty_per_class_aggregator_vs_reference_golden(cuda, golden):
    """Against the reference's own UncertaintyPerClassAggregator (golden from oracle/gen_golden.py::gen_per_class): sample
    counts and per-class means exact (fixed-point sums), quartiles to one histogram bin, and the same class order by mean
    uncertainty that plot_iou_sorted_by_uncertainty uses."""
    from semanticlidarunc_b200.models.evaluator import UncertaintyPerClassAggregator, plot_iou_sorted_by_uncertainty
    g = golden("per_class.npz")
    labels, unc = torch.from_numpy(g["labels"]), torch.from_numpy(g["unc"])
    C = 20
    agg = UncertaintyPerClassAggregator(C)
    agg.update(labels[:1].to(cuda), unc[:1].to(cuda))
    agg.update(labels[1:].to(cuda), unc[1:].to(cuda))
    assert agg._seen_counts == [int(v) for v in g["seen"]]
    st = agg.class_stats().set_index("class_id")
    for c in range(C):
        n, mean, q25, med, q75 = g["stats"][c]
        assert int(st.loc[c, "n"]) == int(n)
        assert abs(st.loc[c, "mean"] - mean) < 2e-7
        for ours, ref in ((st.loc[c, "q25"], q25), (st.loc[c, "median"], med), (st.loc[c, "q75"], q75)):
            assert abs(ours - ref) < 0.01
    names = [str(i) for i in range(C)]
    order = plot_iou_sorted_by_uncertainty(agg, {n: 0.5 for n in names}, names, {i: [0, 0, 0] for i in range(C)}, ignore_ids=(0,))
    assert [int(v) for v in order["class_id"]] == [int(v) for v in g["order_by_mean"]]
