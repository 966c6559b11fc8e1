"""Parity of the evidential (Dirichlet) kernel against the reference's golden vectors and the oracle:
alpha / H / AU / EU / MI within 1e-5 relative (+1e-6 absolute where a difference of O(1) terms is
taken), pred and histogram counts bit-exact on margin-safe pixels."""
import math

import numpy as np
import pytest
import torch

from oracle import metrics as om
from oracle import uncertainty as ou
from semanticlidarunc_b200 import ops, synth
from semanticlidarunc_b200.metrics.ece import ECEAggregator
from semanticlidarunc_b200.models import probability_helper as ph
from semanticlidarunc_b200.models.evaluator import IoUEvaluator
from tests.helpers import rel_close

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-6


@pytest.mark.parametrize("name", ["ev_small", "ev_strong"])
def test_from_outputs_vs_reference_golden(cuda, golden, name):
    g = golden("evidential.npz")
    o = torch.from_numpy(g[name + "/outputs"]).to(cuda)
    r = ops.evidential_reduce(o, from_outputs=True, want=("alpha", "pred", "H", "AU", "EU", "MI"))
    for k, gk, at in (("alpha", "alpha", 0.0), ("H", "H_norm", ATOL), ("AU", "AU", ATOL), ("EU", "EU", 2 * ATOL), ("MI", "MI_norm", ATOL)):
        ok, aerr, rerr = rel_close(r[k].cpu().numpy(), g[name + "/" + gk], RTOL, at)
        assert ok, f"{name}/{k}: abs {aerr:.3e} rel {rerr:.3e}"
    z = torch.from_numpy(g[name + "/outputs"])[:, :20]
    top = torch.softmax(z, 1).topk(2, dim=1).values
    safe = (top[:, 0] - top[:, 1]) > 1e-6
    assert torch.equal(r["pred"].cpu()[safe], torch.from_numpy(g[name + "/pred"])[safe])


def test_probability_helper_functions_vs_golden(cuda, golden):
    g = golden("evidential.npz")
    o = torch.from_numpy(g["ev_small/outputs"])
    alpha = ph.to_alpha_concentrations_from_shape_and_scale(o[:, :20], o[:, 20:21])          # CPU tensors in
    assert alpha.is_cuda
    ok, aerr, rerr = rel_close(alpha.cpu().numpy(), g["ev_small/alpha"], RTOL, 0.0)
    assert ok, (aerr, rerr)
    a = torch.from_numpy(g["ev_small/alpha"]).to(cuda)
    for fn, key, at in ((ph.get_predictive_entropy, "H", ATOL), (ph.get_aleatoric_uncertainty, "AU", ATOL),
                        (ph.get_epistemic_uncertainty, "EU", 2 * ATOL), (ph.get_predictive_entropy_norm, "H_norm", ATOL)):
        ok, aerr, rerr = rel_close(fn(a).cpu().numpy(), g["ev_small/" + key], RTOL, at)
        assert ok, f"{key}: abs {aerr:.3e} rel {rerr:.3e}"
    # running-mean decorator (src/utils/agg.py:85-89)
    ph.get_predictive_entropy_norm.reset()
    ph.get_predictive_entropy_norm.accumulate(a)
    ph.get_predictive_entropy_norm.accumulate(a)
    assert abs(ph.get_predictive_entropy_norm.mean(reset=True) - float(g["ev_small/H_norm"].mean())) < 1e-6
    # normalised variants against the oracle formulas
    AU = ou.get_aleatoric_uncertainty(torch.from_numpy(g["ev_small/alpha"]))
    ok, *_ = rel_close(ph.get_aleatoric_uncertainty_norm(a, mode="max").cpu().numpy(), (AU / math.log(20)).clamp(0, 1).numpy(), RTOL, ATOL)
    assert ok


def test_tester_branch_with_aggregators(cuda):
    C = 20
    o, lab = synth.synth_evidential_logits(77, 2, C, 8, 256)
    # margin-safe labels: drop pixels whose alpha-confidence is within 1e-4 of a bin edge
    r = ou.evidential_reduce(o, C)
    conf = om.to_probs(r["alpha"], "alpha").max(1).values.numpy()
    near = np.abs(conf[..., None] - om.ece_edges(15)[None, None, None, :].astype(np.float64)).min(-1) < 1e-4
    lab[torch.from_numpy(near)] = 0
    iou, ece = IoUEvaluator(C), ECEAggregator(n_bins=15, mode="alpha", ignore_index=0)
    out = ph.evidential_reduce_from_outputs(o, lab, num_classes=C, iou_evaluator=iou, ece_eval=ece)
    top = torch.softmax(o[:, :C], 1).topk(2, dim=1).values
    safe = (top[:, 0] - top[:, 1]) > 1e-5
    assert torch.equal(out["pred"].cpu()[safe], r["pred"][safe])
    if bool(safe.all()):
        assert torch.equal(iou.confmat, om.confusion_counts(r["pred"], lab, C))
    c, k = om.ece_samples(r["alpha"], lab, "alpha", ignore_index=0)
    n, nc, cs = om.ece_bin_counts(c.numpy(), k.numpy(), 15)
    assert np.array_equal(ece._bins[0].cpu().numpy(), n) and np.array_equal(ece._bins[1].cpu().numpy(), nc)
    for kk, rk in (("H", "H_norm"), ("AU", "AU"), ("EU", "EU"), ("MI", "MI_norm")):
        ok, aerr, rerr = rel_close(out[kk].cpu().numpy(), r[rk].numpy(), RTOL, 2 * ATOL)
        assert ok, f"{kk}: abs {aerr:.3e} rel {rerr:.3e}"


def test_digamma_range_through_AU(cuda):
    """alpha from 1 to 1e4 (SURVEY: digamma must hold to 1e-5 over this range)."""
    g = torch.Generator().manual_seed(5)
    alpha = 1.0 + torch.exp(torch.rand((1, 20, 8, 256), generator=g) * math.log(1e4)) - 1.0 + 1e-3
    r = ops.evidential_reduce(alpha.to(cuda), from_outputs=False, want=("AU", "H", "EU"))
    ok, aerr, rerr = rel_close(r["AU"].cpu().numpy(), ou.get_aleatoric_uncertainty(alpha.double()).numpy(), RTOL, ATOL)
    assert ok, f"AU: abs {aerr:.3e} rel {rerr:.3e}"


def test_small_concentrations_take_the_two_eps_path(cuda):
    """alpha0 << 1 (concentrations passed in directly, not from the head): alpha0 + 1e-8 and alpha0 + 1e-12 differ in
    fp32, so the AUROC-convention MI is evaluated against its own psi(alpha0 + 1e-12 + 1) in the kernel's rare path."""
    g = torch.Generator().manual_seed(9)
    alpha = torch.rand((1, 20, 4, 128), generator=g) * 2e-3 + 1e-4
    alpha[:, :, 2:] += 1.0                                   # half of the pixels stay on the common path
    r = ops.evidential_reduce(alpha.to(cuda), from_outputs=False, want=("MI", "AU", "H"))
    ref_mi = ou.dirichlet_mi(alpha.double())
    ok, aerr, rerr = rel_close(r["MI"].cpu().numpy(), ref_mi.numpy(), RTOL, 2 * ATOL)
    assert ok, f"MI: abs {aerr:.3e} rel {rerr:.3e}"
    ok, aerr, rerr = rel_close(r["AU"].cpu().numpy(), ou.get_aleatoric_uncertainty(alpha.double()).numpy(), RTOL, 2 * ATOL)
    assert ok, f"AU: abs {aerr:.3e} rel {rerr:.3e}"
