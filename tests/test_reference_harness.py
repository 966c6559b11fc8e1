"""CPU checks of the test / bench infrastructure that executes the UNMODIFIED reference (oracle/ref_arm.py,
oracle/tester_harness.py).  Skipped where neither /root/reference nor oracle/_ref/reference_src.zip exists."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_arm
from tests import tester_cases as tc

pytestmark = pytest.mark.skipif(not ref_arm.available(), reason="no reference (run oracle/make_ref.sh in the build container)")


def test_archive_holds_unmodified_reference_files():
    """Every file in oracle/_ref/reference_src.zip is byte-identical to /root/reference (when both are present)."""
    import hashlib
    import json
    import zipfile
    if not os.path.isfile(ref_arm.REF_ZIP):
        pytest.skip("archive not built")
    with zipfile.ZipFile(ref_arm.REF_ZIP) as z:
        man = json.loads(z.read("MANIFEST.json"))["sha256"]
        assert len(man) >= 25
        for name, digest in man.items():
            assert hashlib.sha256(z.read(name)).hexdigest() == digest
            live = os.path.join(os.path.dirname(ref_arm.REF_LIVE), name)
            if os.path.isfile(live):
                with open(live, "rb") as f:
                    assert hashlib.sha256(f.read()).hexdigest() == digest, name


@pytest.mark.parametrize("branch", ["mc", "dirichlet"])
def test_stock_tester_loop_runs(tmp_path, branch):
    """The reference's Tester.test_epoch runs end to end on the stub model, and its counts are consistent."""
    from oracle import tester_harness as th
    batches, outs = tc.make_case(branch)
    r = th.run_tester(branch, False, batches, outs, str(tmp_path), tc.C, tc.T)
    assert r["classes"]["iou"] == "models.evaluator" and r["classes"]["ece"] == "metrics.ece"
    assert int(r["confmat"].sum()) == tc.STEPS * tc.H * tc.W
    assert r["model_calls"] == tc.STEPS * (tc.T if branch == "mc" else 1)
    labels = torch.cat([b[4] for b in batches]).reshape(-1)
    assert sum(r["ece_bin_n"]) == int((labels != 0).sum())
    assert 0.0 < r["mIoU"] < 1.0 and 0.5 < r["auroc"] < 1.0
    assert r["summary_saved"]


def test_reference_arm_matches_oracle_port(tmp_path):
    """bench.py's reference arm (stock loader + MC block + IoU / ECE classes) against the oracle port on the same scan."""
    from oracle import metrics as om, projection as oproj, uncertainty as ou
    from semanticlidarunc_b200 import synth
    from semanticlidarunc_b200.dataset.definitions import build_id_lut
    Hh, Ww, Cc, Tt = 16, 256, 20, 3
    scans = [synth.synth_scan(5, "tiny")]
    logits = [synth.synth_mc_logits(3, Tt, 1, Cc, Hh, Ww)[0]]
    arm = ref_arm.ReferenceArm(scans, logits, H=Hh, W=Ww, C=Cc, workers=0, total_items=1)
    try:
        out = arm.scan_pass()
        fr = oproj.kitti_frame(scans[0][0], scans[0][1], Hh, Ww, build_id_lut())
        assert np.array_equal(out["labels"][0].numpy(), fr["semantics"][0])
        r = ou.mc_reduce(logits[0])
        assert torch.equal(out["pred"], r["pred"])
        np.testing.assert_allclose(out["H_norm"].numpy(), r["H_norm"].numpy(), rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(out["MI_norm"].numpy(), r["MI_norm"].numpy(), rtol=1e-5, atol=1e-6)
        assert np.array_equal(out["point_labels"], r["pred"][0].numpy().reshape(-1)[fr["pix"]])
        res = arm.finish()
        assert res["confmat_sum"] == Hh * Ww
        cm = om.confusion_counts(r["pred"], torch.from_numpy(fr["semantics"]), Cc)
        assert torch.equal(arm.iou.confmat, cm)
    finally:
        arm.close()
