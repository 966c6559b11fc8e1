"""MC-dropout reductions with the reference's names (src/utils/mc_dropout.py:121-133 and the
closures of src/models/tester.py:419-451), computed by the fused kernel.

`mc_forward` and the dropout toggles are the backbone's side of the interface (the network only supplies the
logits, BASELINE.json north_star); they are provided with the reference's names and behaviour
(src/utils/mc_dropout.py:13-35,99-119) so that `utils.mc_dropout` can be swapped as a whole module.
"""
from __future__ import annotations

import contextlib

import torch

from .. import ops

_DROPOUT_TYPES = (torch.nn.Dropout, torch.nn.Dropout2d, torch.nn.Dropout3d, torch.nn.AlphaDropout,
                  torch.nn.FeatureAlphaDropout)


def set_dropout_mode(module: torch.nn.Module, train: bool) -> None:
    """Switch the dropout layers (and nothing else: batch-norm statistics stay frozen) to train / eval."""
    for m in module.modules():
        if isinstance(m, _DROPOUT_TYPES):
            m.train(train)


@contextlib.contextmanager
def dropout_sampling(module: torch.nn.Module, enable: bool = True):
    """Dropout layers sample inside the block and are put back to eval afterwards."""
    if enable:
        set_dropout_mode(module, True)
    try:
        yield
    finally:
        if enable:
            set_dropout_mode(module, False)


@torch.no_grad()
def mc_forward(model: torch.nn.Module, inputs, T: int = 30) -> torch.Tensor:
    """T stochastic forward passes with the model in eval mode and only its dropout layers sampling:
    [T,B,C,H,W] raw model outputs, written straight into one preallocated tensor (the layout
    slu_reduce_metrics streams) instead of a list + cat."""
    model.eval()
    out = None
    with dropout_sampling(model, enable=True):
        for t in range(int(T)):
            y = model(*inputs)
            if out is None:
                out = torch.empty((int(T),) + tuple(y.shape), dtype=y.dtype, device=y.device)
            out[t].copy_(y)
    return out


@torch.no_grad()
def predictive_entropy_mc(mc_probs: torch.Tensor, eps: float = 1e-12, normalize: bool = True):
    """mc_probs [T,B,C,H,W] probabilities -> entropy of the mean [B,H,W] (optionally / log C)."""
    return ops.reduce_metrics(mc_probs, kind="probs", eps=eps, want=("H_norm",), normalize=normalize)["H_norm"]


@torch.no_grad()
def mc_predictive_entropy_norm(probs: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """tester.py:419-425 on probabilities [T,B,C,H,W]."""
    return ops.reduce_metrics(probs, kind="probs", eps=eps, want=("H_norm",))["H_norm"]


@torch.no_grad()
def mc_mutual_information_norm(probs: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """tester.py:427-451 on probabilities [T,B,C,H,W]."""
    return ops.reduce_metrics(probs, kind="probs", eps=eps, want=("MI_norm",))["MI_norm"]


@torch.no_grad()
def mc_reduce_from_logits(mc_logits: torch.Tensor, labels: torch.Tensor | None = None, *, eps: float = 1e-12,
                          ignore_index=None, iou_evaluator=None, ece_eval=None, auroc_eval=None, auroc_eval_mi=None,
                          ua_agg=None, unc_agg=None, ua_ignore_ids=(0,), want=("pred", "conf", "H_norm", "MI_norm")) -> dict:
    """The whole MC block of Tester.test_epoch (tester.py:412-471) in one pass over the logits.

    mc_logits [T,B,C,H,W] straight from `mc_forward`; returns pred / conf / H_norm / MI_norm (and
    p_bar if asked) and, when given, updates `iou_evaluator` (models.evaluator.IoUEvaluator) and
    `ece_eval` (metrics.ece.ECEAggregator, mode 'probs' semantics) in the same kernel, then
    `auroc_eval` (entropy score, :468), `auroc_eval_mi` (MI score override, :470-471) and `ua_agg`
    (accuracy vs H_norm, :460-464) from the small per-pixel maps -- p_bar is never materialised.
    """
    confmat = iou_evaluator._accumulator(mc_logits.device) if iou_evaluator is not None else None
    bins = ece_eval._accumulator(mc_logits.device) if ece_eval is not None else None
    edges = ece_eval._edges if ece_eval is not None else None
    if ece_eval is not None and ignore_index is None:
        ignore_index = ece_eval.ignore_index
    need = set(want)
    if auroc_eval is not None or ua_agg is not None or unc_agg is not None:
        need |= {"pred", "H_norm"}
    if auroc_eval_mi is not None:
        need |= {"pred", "MI_norm"}
    out = ops.reduce_metrics(mc_logits, labels, kind="logits", conf_mode=ops.CONF_RENORM, eps=eps,
                             ignore_index=ignore_index, edges=edges, confmat=confmat, ece_bins=bins, want=tuple(need))
    if labels is not None:
        lab = labels[:, 0] if labels.dim() == 4 else labels
        if auroc_eval is not None:
            auroc_eval.add_maps(out["H_norm"], out["pred"], lab)
        if auroc_eval_mi is not None:
            auroc_eval_mi.add_maps(out["MI_norm"], out["pred"], lab)
        if ua_agg is not None:
            ua_agg.update(labels=lab, preds=out["pred"], uncertainty=out["H_norm"], ignore_ids=ua_ignore_ids)
        if unc_agg is not None:
            unc_agg.update(lab, out["H_norm"])          # tester.py:516
    return out
