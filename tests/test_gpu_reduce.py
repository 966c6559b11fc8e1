"""Parity of the fused reduction + metrics kernel (slu_reduce_metrics) against the oracle and the
reference's golden vectors.  Tolerances (BASELINE.json): counts and indices bit-exact on
margin-enforced inputs; entropy / MI / confidence within 1e-5 relative (atol 1e-6 absorbs the
cancellation in MI = H - E[H], whose operands are O(1))."""
import math

import numpy as np
import pytest
import torch

from oracle import metrics as om
from oracle import uncertainty as ou
from semanticlidarunc_b200 import ops, synth
from tests.helpers import enforce_margins, rel_close

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-6


def run_both(x, lab, C, cuda, direct=False, ignore_index=0, n_bins=15, want_pbar=True):
    confmat = ops.new_confmat(C, cuda)
    bins = ops.new_ece_bins(n_bins, cuda)
    want = ("p_bar", "pred", "conf", "H_norm", "MI_norm") if want_pbar else ("pred", "conf", "H_norm", "MI_norm")
    out = ops.reduce_metrics(x.to(cuda), lab.to(cuda), kind="logits", conf_mode=ops.CONF_RENORM,
                             ignore_index=ignore_index, confmat=confmat, ece_bins=bins, want=want, direct=direct)
    torch.cuda.synchronize()
    ref = ou.mc_reduce(x)
    ref["confmat"] = om.confusion_counts(ref["pred"], lab, C)
    conf, corr = om.ece_samples(ref["p_bar"], lab, "probs", ignore_index=ignore_index)
    ref["bins"] = om.ece_bin_counts(conf.numpy(), corr.numpy(), n_bins)
    ref["conf_map"] = om.to_probs(ref["p_bar"], "probs").max(dim=1).values
    return out, confmat.cpu(), bins.cpu(), ref


def check(out, confmat, bins, ref, exact=True):
    for k, rk in (("H_norm", "H_norm"), ("MI_norm", "MI_norm"), ("conf", "conf_map")):
        ok, aerr, rerr = rel_close(out[k].cpu().numpy(), ref[rk].numpy(), RTOL, ATOL)
        assert ok, f"{k}: max abs err {aerr:.3e}, max rel err {rerr:.3e}"
    if "p_bar" in out:
        ok, aerr, rerr = rel_close(out["p_bar"].cpu().numpy(), ref["p_bar"].numpy(), RTOL, 1e-9)
        assert ok, f"p_bar: max abs err {aerr:.3e}, max rel err {rerr:.3e}"
    if exact:
        assert torch.equal(out["pred"].cpu(), ref["pred"])
        assert torch.equal(confmat, ref["confmat"])
        n, nc, cs = ref["bins"]
        assert np.array_equal(bins[0].numpy(), n)
        assert np.array_equal(bins[1].numpy(), nc)
        got = bins[2].numpy().astype(np.float64) / 2.0 ** 32
        assert np.allclose(got, cs, rtol=1e-6, atol=1e-6)
    else:
        flips = int((out["pred"].cpu() != ref["pred"]).sum())
        assert flips <= max(2, out["pred"].numel() // 20000), f"{flips} argmax flips on unconstrained input"


@pytest.mark.parametrize("direct", [False, True])
@pytest.mark.parametrize("shape", [(20, 2, 20, 8, 256), (5, 3, 20, 4, 192), (1, 2, 20, 4, 256), (4, 1, 7, 3, 40),
                                   (3, 2, 13, 5, 100), (2, 1, 32, 2, 128), (6, 1, 2, 2, 64)])
def test_mc_logits_vs_oracle(cuda, shape, direct):
    T, B, C, H, W = shape
    x, lab = synth.synth_mc_logits(7 + T + C, T, B, C, H, W)
    x, lab = enforce_margins(x, lab)
    check(*run_both(x, lab, C, cuda, direct=direct))


def test_ragged_width_falls_back_to_direct(cuda):
    # HW % 4 != 0 cannot use 16-byte bulk copies: slu_reduce_metrics must dispatch the direct kernel itself
    x, lab = synth.synth_mc_logits(3, 4, 2, 20, 3, 37)
    x, lab = enforce_margins(x, lab)
    check(*run_both(x, lab, 20, cuda))


def test_unconstrained_inputs_flip_count(cuda):
    x, lab = synth.synth_mc_logits(5, 20, 1, 20, 16, 512)
    check(*run_both(x, lab, 20, cuda), exact=False)


def test_golden_reference_vectors(cuda, golden):
    g = golden("mc_reduce.npz")
    for name in ("mc_small", "mc_peaked", "mc_c7"):
        x = torch.from_numpy(g[name + "/logits"])
        out = ops.reduce_metrics(x.to(cuda), None, kind="logits", want=("p_bar", "pred", "H_norm", "MI_norm"))
        for k in ("H_norm", "MI_norm"):
            ok, aerr, rerr = rel_close(out[k].cpu().numpy(), g[name + "/" + k], RTOL, ATOL)
            assert ok, f"{name}/{k}: abs {aerr:.3e} rel {rerr:.3e}"
        ok, aerr, rerr = rel_close(out["p_bar"].cpu().numpy(), g[name + "/p_bar"], RTOL, 1e-9)
        assert ok, f"{name}/p_bar: abs {aerr:.3e} rel {rerr:.3e}"
        # argmax may only differ where the reference's own top-2 gap is at rounding level
        pb = torch.from_numpy(g[name + "/p_bar"])
        top = pb.topk(2, dim=1).values
        safe = (top[:, 0] - top[:, 1]) > 1e-5
        assert torch.equal(out["pred"].cpu()[safe], torch.from_numpy(g[name + "/pred"])[safe])


def test_probs_kind_and_entropy_mc(cuda, golden):
    from semanticlidarunc_b200.utils.mc_dropout import (mc_mutual_information_norm, mc_predictive_entropy_norm,
                                                        predictive_entropy_mc)
    g = golden("mc_reduce.npz")
    x = torch.from_numpy(g["mc_small/logits"])
    probs = torch.softmax(x, dim=2).to(cuda)
    for fn, key in ((predictive_entropy_mc, "H_mc"), (mc_predictive_entropy_norm, "H_norm"), (mc_mutual_information_norm, "MI_norm")):
        ok, aerr, rerr = rel_close(fn(probs).cpu().numpy(), g["mc_small/" + key], RTOL, ATOL)
        assert ok, f"{key}: abs {aerr:.3e} rel {rerr:.3e}"
    raw = predictive_entropy_mc(probs, normalize=False).cpu().numpy()
    ok, aerr, rerr = rel_close(raw, g["mc_small/H_mc"] * math.log(20), RTOL, ATOL)
    assert ok


def test_large_eps_uses_literal_clamp(cuda):
    x, lab = synth.synth_mc_logits(9, 6, 1, 20, 4, 128, scale=6.0)
    out = ops.reduce_metrics(x.to(cuda), None, kind="logits", eps=1e-3, want=("H_norm", "MI_norm"))
    ref = ou.mc_reduce(x, eps=1e-3)
    for k in ("H_norm", "MI_norm"):
        ok, aerr, rerr = rel_close(out[k].cpu().numpy(), ref[k].numpy(), RTOL, ATOL)
        assert ok, f"{k}: abs {aerr:.3e} rel {rerr:.3e}"


def test_neg_inf_logits(cuda):
    x, lab = synth.synth_mc_logits(10, 3, 1, 20, 2, 64)
    x[:, :, 5] = float("-inf")                      # a masked class
    out = ops.reduce_metrics(x.to(cuda), None, kind="logits", want=("H_norm", "MI_norm", "pred"))
    ref = ou.mc_reduce(x)
    for k in ("H_norm", "MI_norm"):
        got = out[k].cpu().numpy()
        assert np.isfinite(got).all()
        ok, aerr, rerr = rel_close(got, ref[k].numpy(), RTOL, ATOL)
        assert ok, f"{k}: abs {aerr:.3e} rel {rerr:.3e}"


def test_accumulators_add_and_ignore_semantics(cuda):
    C = 20
    x, lab = synth.synth_mc_logits(11, 2, 2, C, 4, 64)
    x, lab = enforce_margins(x, lab)
    lab[0, 0, :8] = -1          # out-of-range labels: dropped from confmat, still ECE samples (ece.py:78)
    lab[0, 1, :8] = C + 3
    confmat = ops.new_confmat(C, cuda)
    bins = ops.new_ece_bins(15, cuda)
    for _ in range(2):           # two updates accumulate
        ops.reduce_metrics(x.to(cuda), lab.to(cuda), kind="logits", conf_mode=ops.CONF_RENORM, ignore_index=0,
                           confmat=confmat, ece_bins=bins, want=())
    ref = ou.mc_reduce(x)
    cm = om.confusion_counts(ref["pred"], lab, C)
    conf, corr = om.ece_samples(ref["p_bar"], lab, "probs", ignore_index=0)
    n, nc, cs = om.ece_bin_counts(conf.numpy(), corr.numpy(), 15)
    assert torch.equal(confmat.cpu(), 2 * cm)
    assert np.array_equal(bins[0].cpu().numpy(), 2 * n) and np.array_equal(bins[1].cpu().numpy(), 2 * nc)
    assert int(confmat.sum()) == 2 * int(((lab >= 0) & (lab < C)).sum())      # histogram invariant
    assert int(bins[0].sum()) == 2 * int((lab != 0).sum())


def test_argument_errors(cuda):
    x = torch.zeros((2, 1, 40, 2, 8), device=cuda)
    with pytest.raises(ValueError):
        ops.reduce_metrics(x, None, kind="logits")           # C > 32
    x = torch.zeros((2, 1, 20, 2, 8), device=cuda)
    with pytest.raises(ValueError):
        ops.reduce_metrics(x, None, kind="alpha")            # alpha needs T == 1
    with pytest.raises(ValueError):
        ops.reduce_metrics(x, None, kind="logits", confmat=ops.new_confmat(20, cuda))    # histogram without labels


@pytest.mark.parametrize("shape", [(2, 20, 8, 256), (3, 20, 4, 192), (1, 7, 3, 37), (2, 13, 5, 100), (1, 32, 2, 128), (1, 2, 2, 64), (16, 20, 2, 130)])
def test_single_sample_kernel_vs_oracle_and_staged(cuda, shape):
    """T == 1 takes reduce_single_kernel (one thread per pixel, entropy from the softmax sums); it must agree with the
    oracle to 1e-5 and, on margin-enforced inputs, bit for bit in pred / counts with the multi-sample kernel."""
    from semanticlidarunc_b200 import _lib
    B, C, H, W = shape
    x, lab = synth.synth_mc_logits(100 + B + C, 1, B, C, H, W)
    x, lab = enforce_margins(x, lab)
    out, cm, bins, ref = run_both(x, lab, C, cuda)
    check(out, cm, bins, ref)
    assert float(out["MI_norm"].abs().max()) == 0.0                  # T == 1: MI is exactly 0 in the reference
    _lib.lib().slu_debug_reduce_no_single(1)
    try:
        out2, cm2, bins2, _ = run_both(x, lab, C, cuda)
    finally:
        _lib.lib().slu_debug_reduce_no_single(0)
    assert torch.equal(out["pred"], out2["pred"]) and torch.equal(cm, cm2) and torch.equal(bins[:2], bins2[:2])
    assert torch.allclose(out["H_norm"], out2["H_norm"], rtol=1e-5, atol=1e-6)
    assert torch.equal(out["p_bar"], out2["p_bar"]) and torch.equal(out["conf"], out2["conf"])


def test_single_sample_probs_alpha_large_eps_and_neg_inf(cuda):
    g = torch.Generator().manual_seed(77)
    x = torch.randn((1, 2, 20, 4, 100), generator=g) * 4.0
    # probabilities and concentrations (ECE / AUROC modes) through the same kernel
    probs = torch.softmax(x, dim=2)
    ref = ou.mc_reduce(x)
    out = ops.reduce_metrics(probs[0].to(cuda), None, kind="probs", want=("H_norm", "pred", "conf"))
    ok, aerr, rerr = rel_close(out["H_norm"].cpu().numpy(), ref["H_norm"].numpy(), RTOL, ATOL)
    assert ok, (aerr, rerr)
    alpha = torch.rand((2, 20, 4, 100), generator=g) * 30 + 1.0
    out = ops.reduce_metrics(alpha.to(cuda), None, kind="alpha", want=("conf", "pred"))
    pa = om.to_probs(alpha, "alpha")
    assert torch.equal(out["pred"].cpu(), pa.argmax(1))
    ok, aerr, rerr = rel_close(out["conf"].cpu().numpy(), pa.max(1).values.numpy(), RTOL, 0.0)
    assert ok, (aerr, rerr)
    # eps large enough to matter: the literal clamp path
    out = ops.reduce_metrics(x.to(cuda), None, kind="logits", eps=1e-3, want=("H_norm",))
    ok, aerr, rerr = rel_close(out["H_norm"].cpu().numpy(), ou.mc_reduce(x, eps=1e-3)["H_norm"].numpy(), RTOL, ATOL)
    assert ok, (aerr, rerr)
    # a masked class (-inf logit) and a pixel whose logits are all huge
    x2 = x.clone()
    x2[:, :, 5] = float("-inf")
    x2[:, 0, :, 0, :3] += 80.0
    out = ops.reduce_metrics(x2.to(cuda), None, kind="logits", want=("H_norm", "pred"))
    got = out["H_norm"].cpu().numpy()
    assert np.isfinite(got).all()
    ok, aerr, rerr = rel_close(got, ou.mc_reduce(x2)["H_norm"].numpy(), RTOL, ATOL)
    assert ok, (aerr, rerr)


def test_full_size_batch_properties(cuda):
    """BASELINE.json configs[1] at full size ([T=20,B=16,C=20,64,2048] fp32 = 3.36 GB; the oracle cannot run this in
    seconds): size-independent properties, the two load paths against each other, and an oracle check on a slice."""
    T, B, C, H, W = 20, 16, 20, 64, 2048
    g = torch.Generator(device=cuda).manual_seed(2024)
    x = torch.randn((T, B, C, H, W), generator=g, device=cuda) * 3.0
    lab = torch.randint(-1, C + 1, (B, H, W), generator=g, device=cuda)            # includes out-of-range labels
    cm, bins = ops.new_confmat(C, cuda), ops.new_ece_bins(15, cuda)
    out = ops.reduce_metrics(x, lab, kind="logits", conf_mode=ops.CONF_RENORM, ignore_index=0, confmat=cm, ece_bins=bins,
                             want=("p_bar", "pred", "conf", "H_norm", "MI_norm"))
    # histogram invariants (a checksum of the counts)
    assert int(cm.sum()) == int(((lab >= 0) & (lab < C)).sum())
    assert int(bins[0].sum()) == int((lab != 0).sum()) and bool((bins[1] <= bins[0]).all())
    assert torch.equal(cm.sum(dim=1), torch.bincount(lab[(lab >= 0) & (lab < C)].reshape(-1), minlength=C))
    # the maps are consistent with the kernel's own mean distribution
    pb = out["p_bar"]
    assert torch.equal(out["pred"], pb.argmax(dim=1))
    assert float((pb.sum(dim=1) - 1.0).abs().max()) < 2e-6
    assert torch.allclose(out["conf"], pb.max(dim=1).values / pb.clamp_min(0).sum(dim=1), rtol=1e-6, atol=0)
    assert float(out["H_norm"].min()) >= 0.0 and float(out["H_norm"].max()) <= 1.0 + 1e-6
    assert float(out["MI_norm"].min()) >= 0.0 and bool((out["MI_norm"] <= out["H_norm"] + 1e-6).all())
    # shard additivity: the two halves of the batch add up to the whole (what the multi-GPU all-reduce relies on)
    cm2, bins2 = ops.new_confmat(C, cuda), ops.new_ece_bins(15, cuda)
    for sl in (slice(0, 8), slice(8, 16)):
        ops.reduce_metrics(x[:, sl].contiguous(), lab[sl], kind="logits", conf_mode=ops.CONF_RENORM, ignore_index=0, confmat=cm2,
                           ece_bins=bins2, want=())
    assert torch.equal(cm, cm2) and torch.equal(bins, bins2)
    # TMA-staged and direct-load kernels run the same arithmetic: bit-identical maps and counts
    cm3, bins3 = ops.new_confmat(C, cuda), ops.new_ece_bins(15, cuda)
    out3 = ops.reduce_metrics(x, lab, kind="logits", conf_mode=ops.CONF_RENORM, ignore_index=0, confmat=cm3, ece_bins=bins3,
                              want=("pred", "conf", "H_norm", "MI_norm"), direct=True)
    assert torch.equal(cm, cm3) and torch.equal(bins, bins3)
    for k in ("pred", "conf", "H_norm", "MI_norm"):
        assert torch.equal(out[k], out3[k]), k
    # oracle on the last scan's last two image rows (offsets at full-size strides)
    xs = x[:, 15:16, :, 62:64, :].cpu()
    ref = ou.mc_reduce(xs)
    for k in ("H_norm", "MI_norm"):
        ok, aerr, rerr = rel_close(out[k][15:16, 62:64].cpu().numpy(), ref[k].numpy(), RTOL, ATOL)
        assert ok, f"{k}: abs {aerr:.3e} rel {rerr:.3e}"
    top = ref["p_bar"].topk(2, dim=1).values
    safe = (top[:, 0] - top[:, 1]) > 1e-5
    assert torch.equal(out["pred"][15:16, 62:64].cpu()[safe], ref["pred"][safe])


def test_periodic_histogram_flush_in_long_sweeps(cuda):
    """The single-sample and evidential kernels keep sum(conf * 2^32) as two 32-bit halves per CTA and flush them to the
    int64 accumulators before they can overflow.  160 scans in one launch make every CTA flush in mid-loop; the two
    80-scan halves stay below the threshold.  All counters, including the fixed-point confidence sums, must be equal."""
    B, C, H, W = 160, 20, 64, 2048
    g = torch.Generator(device=cuda).manual_seed(7)
    x = torch.randn((B, C + 1, H, W), generator=g, device=cuda) * 3.0
    lab = torch.randint(0, C, (B, H, W), generator=g, device=cuda)

    def single(xs, ls, cm, bins):
        ops.reduce_metrics(xs[:, :C].contiguous(), ls, kind="logits", conf_mode=ops.CONF_RENORM, ignore_index=0, confmat=cm, ece_bins=bins, want=())

    def evid(xs, ls, cm, bins):
        ops.evidential_reduce(xs, ls, from_outputs=True, ignore_index=0, confmat=cm, ece_bins=bins, want=())

    for fn in (single, evid):
        cm, bins = ops.new_confmat(C, cuda), ops.new_ece_bins(15, cuda)
        fn(x, lab, cm, bins)
        cm2, bins2 = ops.new_confmat(C, cuda), ops.new_ece_bins(15, cuda)
        fn(x[:80], lab[:80], cm2, bins2)
        fn(x[80:], lab[80:], cm2, bins2)
        assert int(cm.sum()) == B * H * W
        assert torch.equal(cm, cm2) and torch.equal(bins, bins2), fn.__name__
        assert int(bins[2].max()) > 2 ** 40                     # far beyond what a 32-bit half could hold


def test_single_sample_argmax_with_nan_matches_torch(cuda):
    """torch.argmax treats NaN as maximal (first NaN wins); the single-sample kernel only runs its NaN-aware comparison for
    pixels whose distribution contains one."""
    g = torch.Generator().manual_seed(5)
    probs = torch.softmax(torch.randn((2, 20, 4, 64), generator=g) * 3.0, dim=1)
    probs[0, 7, 1, :10] = float("nan")
    probs[1, 3, 2, 5:9] = float("nan")
    probs[1, 12, 2, 5:9] = float("nan")                     # two NaNs: the first index wins
    probs[0, 0, 3, :4] = float("inf")
    out = ops.reduce_metrics(probs.to(cuda), None, kind="probs", want=("pred",))
    assert torch.equal(out["pred"].cpu(), probs.argmax(dim=1))
    logits = torch.randn((1, 20, 2, 64), generator=g)
    logits[0, 4, 0, :8] = float("nan")
    out = ops.reduce_metrics(logits.to(cuda), None, kind="logits", want=("pred",))
    assert torch.equal(out["pred"].cpu(), torch.softmax(logits, dim=1).argmax(dim=1))
