"""The reference's own evaluation loop, Tester.test_epoch (src/models/tester.py:272-720), run twice on the GPU box with a
stub model that replays identical head outputs: once with the stock reference classes, once with the INTEGRATION.md
section A import swap (sys.modules pre-seeded with the semanticlidarunc_b200 mirrors; not a line of the reference
changed).  Both branches of the loop: MC-dropout (:405-471) and single-pass Dirichlet (:484-512)."""
import numpy as np
import pytest
import torch

from oracle import ref_arm
from tests import tester_cases as tc

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_arm.available(), reason="oracle/_ref/reference_src.zip missing (oracle/make_ref.sh)")]


@pytest.mark.parametrize("branch", ["mc", "dirichlet"])
def test_tester_loop_stock_vs_import_swap(cuda, tmp_path, branch):
    from oracle import tester_harness as th
    from semanticlidarunc_b200 import _lib
    batches, outs = tc.make_case(branch)
    ref = th.run_tester(branch, False, batches, outs, str(tmp_path / "stock"), tc.C, tc.T)
    n0 = _lib.launch_count()
    ours = th.run_tester(branch, True, batches, outs, str(tmp_path / "swap"), tc.C, tc.T)
    assert _lib.launch_count() - n0 >= 5 * tc.STEPS, "the swapped loop must run on libslu kernels"
    assert ref["classes"]["iou"] == "models.evaluator"
    assert all(v.startswith("semanticlidarunc_b200.") for v in ours["classes"].values()), ours["classes"]
    assert ours["model_calls"] == ref["model_calls"]
    # integer results: exact
    assert torch.equal(ours["confmat"], ref["confmat"])
    assert ours["mIoU"] == pytest.approx(ref["mIoU"], abs=1e-12)
    for k, v in ref["iou"].items():
        if v is None:
            assert ours["iou"][k] is None
        else:
            assert ours["iou"][k] == pytest.approx(v, abs=1e-12)
    assert ours["unc_seen"] == ref["unc_seen"]
    # reliability bins: counts may differ only by pixels whose confidence sits on a bin edge (none expected here)
    assert sum(ours["ece_bin_n"]) == sum(ref["ece_bin_n"])
    assert sum(abs(a - b) for a, b in zip(ours["ece_bin_n"], ref["ece_bin_n"])) <= 2
    assert ours["ece"] == pytest.approx(ref["ece"], rel=1e-4, abs=2e-6)
    assert ours["mce"] == pytest.approx(ref["mce"], rel=1e-3, abs=1e-3)
    # ranking metrics from the 60000-bin histograms
    assert ours["auroc"] == pytest.approx(ref["auroc"], abs=2e-4)
    assert ours["auroc_mi"] == pytest.approx(ref["auroc_mi"], abs=2e-4)
    # accuracy-vs-uncertainty bins (0.05 wide): counts equal up to pixels on a coarse edge, accuracies follow
    assert sum(ours["ua_n"]) == sum(ref["ua_n"])
    assert sum(abs(a - b) for a, b in zip(ours["ua_n"], ref["ua_n"])) <= 4
    for a, b, n in zip(ours["ua_acc"], ref["ua_acc"], ref["ua_n"]):
        if n > 200:
            assert a == pytest.approx(b, abs=5e-3)
    assert ours["summary_saved"]


def test_adaptive_binning_vs_reference(cuda):
    """binning='adaptive' (equal-mass edges, src/metrics/ece.py:118-126): the fine-histogram version against the stock class."""
    ref_arm.install()
    from metrics.ece import ECEAggregator as RefECE
    from semanticlidarunc_b200.metrics.ece import ECEAggregator
    g = torch.Generator().manual_seed(4)
    logits = torch.randn(2, 12, 32, 128, generator=g) * 2.0
    labels = torch.randint(0, 12, (2, 32, 128), generator=g)
    logits.scatter_add_(1, labels[:, None], torch.full((2, 1, 32, 128), 1.5))
    a = RefECE(n_bins=10, mode="logits", ignore_index=0, binning="adaptive")
    a.update(logits, labels)
    (e_ref, m_ref), s_ref, _ = a.compute(save_plot_path="/tmp/_ece_adaptive_ref.png")
    b = ECEAggregator(n_bins=10, mode="logits", ignore_index=0, binning="adaptive")
    b.update(logits.to(cuda), labels.to(cuda))
    (e, m), s, _ = b.compute()
    assert len(s) == len(s_ref)
    # np.quantile interpolates between neighbouring samples (7 700 of them here, ~1e-4 apart in confidence); the
    # histogram reads the edge off its CDF: they agree to the local sample spacing
    np.testing.assert_allclose(s["low"].to_numpy(), s_ref["low"].to_numpy(), atol=4e-4)
    assert abs(int(s["n"].sum()) - int(s_ref["n"].sum())) == 0
    assert np.abs(s["n"].to_numpy() - s_ref["n"].to_numpy()).max() <= 0.02 * s_ref["n"].max()
    assert e == pytest.approx(e_ref, abs=2e-3)
