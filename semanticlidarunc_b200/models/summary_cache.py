"""Cached evaluation summary.

The reference Tester saves the RAW per-pixel buffers of its aggregators to `outputs_summary/summary_epoch_*.pt`
so a finished evaluation can be re-plotted without inference (src/models/tester.py:615-654) and restores them by
writing private attributes (:306-360).  The device aggregators have no per-pixel buffers; their whole state is a
handful of small integer histograms, which is what this cache stores (a few MB instead of gigabytes).

Keys: "meta" and "iou_confmat" as in the reference; "ece_bins" [3,n_bins], "auroc_hist" / "auroc_mi_hist" /
"ua_hist" [2,M], "unc_hist" [C,Mc] + "unc_sum_fx" [C] replace ece_conf/ece_correct, auroc_scores/auroc_is_error,
ua_uncert/ua_correct and unc_values/unc_seen_counts.
"""
from __future__ import annotations

import torch


def _cpu(t):
    return None if t is None else t.detach().cpu()


def build_summary(meta: dict, iou_evaluator=None, ece_eval=None, auroc_eval=None, auroc_eval_mi=None, ua_agg=None, unc_agg=None) -> dict:
    return {
        "meta": dict(meta),
        "iou_confmat": None if iou_evaluator is None else iou_evaluator.confmat.detach().cpu(),
        "ece_bins": None if ece_eval is None else _cpu(ece_eval._bins),
        "auroc_hist": None if auroc_eval is None else _cpu(auroc_eval._hist),
        "auroc_mi_hist": None if auroc_eval_mi is None else _cpu(auroc_eval_mi._hist),
        "ua_hist": None if ua_agg is None else _cpu(ua_agg._hist),
        "unc_hist": None if unc_agg is None else _cpu(unc_agg._hist),
        "unc_sum_fx": None if unc_agg is None else _cpu(unc_agg._sum),
    }


def save_summary(path: str, meta: dict, **aggregators) -> None:
    torch.save(build_summary(meta, **aggregators), path)


def restore_summary(cache: dict, iou_evaluator=None, ece_eval=None, auroc_eval=None, auroc_eval_mi=None, ua_agg=None, unc_agg=None) -> None:
    """Load counters back into (fresh) aggregators; raises KeyError on a missing section, like the reference does."""
    def need(key):
        v = cache.get(key)
        if v is None:
            raise KeyError(f"Summary file missing keys: ['{key}']")
        return v
    from .. import _lib
    dev = _lib.require_cuda()
    if iou_evaluator is not None:
        iou_evaluator.confmat = need("iou_confmat").clone().long()
    if ece_eval is not None:
        ece_eval._bins = need("ece_bins").clone().to(dev)
    for agg, key in ((auroc_eval, "auroc_hist"), (auroc_eval_mi, "auroc_mi_hist"), (ua_agg, "ua_hist")):
        if agg is not None:
            agg._hist = need(key).clone().to(dev)
    if unc_agg is not None:
        unc_agg._hist = need("unc_hist").clone().to(dev)
        unc_agg._sum = need("unc_sum_fx").clone().to(dev)


def load_summary(path: str, **aggregators) -> dict:
    cache = torch.load(path, map_location="cpu")
    restore_summary(cache, **aggregators)
    return cache
