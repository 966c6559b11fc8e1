"""The cases of tests/golden/loss_terms.npz (oracle/gen_golden.py::gen_loss_terms), shared by the CPU test of the
oracle and the GPU test of the CUDA path.  Each entry: name -> (input key, oracle callable, module factory)."""
import functools

from oracle import losses as ol


def cases(target, keep):
    from semanticlidarunc_b200.losses import dirichlet_losses as dl
    from semanticlidarunc_b200.losses import regularizers as rg
    P = functools.partial
    return {
        "comp": ("alpha", P(ol.complement_kl_uniform, target=target, ignore_index=0),
                 lambda: (dl.ComplementKLUniform(ignore_index=0), dict(target=target), True)),
        "comp_trainer": ("alpha", P(ol.complement_kl_uniform, target=target, ignore_index=0, gamma=1.25, tau=0.65, sigma=0.15),
                         lambda: (dl.ComplementKLUniform(ignore_index=0, gamma=1.25, tau=0.65, sigma=0.15), dict(target=target), True)),
        "comp_evid_gate": ("alpha", P(ol.complement_kl_uniform, target=target, ignore_index=0, s_target=40.0, normalize=False),
                           lambda: (dl.ComplementKLUniform(ignore_index=0, s_target=40.0, normalize=False), dict(target=target), True)),
        "comp_attached": ("alpha", P(ol.complement_kl_uniform, target=target, ignore_index=0, detach_uncert=False),
                          lambda: (dl.ComplementKLUniform(ignore_index=0, detach_uncert=False), dict(target=target), True)),
        "wle": ("alpha", P(ol.wrong_low_evidence, target=target, ignore_index=0),
                lambda: (rg.WrongLowEvidence(ignore_index=0, s_low=0.0, margin=0.05, soft_margin_k=0.08), dict(target=target), True)),
        "wle_hard": ("alpha", P(ol.wrong_low_evidence, target=target, ignore_index=0, s_low=2.0, margin=0.1, soft_margin_k=0.0),
                     lambda: (rg.WrongLowEvidence(ignore_index=0, s_low=2.0, margin=0.1, soft_margin_k=0.0), dict(target=target), True)),
        "wle_nomargin": ("alpha", P(ol.wrong_low_evidence, target=target, ignore_index=None, margin=0.0),
                         lambda: (rg.WrongLowEvidence(ignore_index=None, margin=0.0), dict(target=target), True)),
        "band": ("alpha", P(ol.evidence_reg_band, s_target=60.0, band=0.10, ignore_index=0, target=target),
                 lambda: (rg.EvidenceRegBand(60.0, band=0.10, ignore_index=0), dict(target=target), False)),
        "band_mask": ("alpha", P(ol.evidence_reg_band, s_target=200.0, band=0.25, mask=keep),
                      lambda: (rg.EvidenceRegBand(200.0, band=0.25), dict(mask=keep), False)),
        "band_nomask": ("alpha", P(ol.evidence_reg_band, s_target=30.0),
                        lambda: (rg.EvidenceRegBand(30.0), dict(), False)),
        "ereg_log": ("alpha", P(ol.evidence_reg, s_target=60.0, ignore_index=0, target=target),
                     lambda: (rg.EvidenceReg(60.0, ignore_index=0), dict(target=target), False)),
        "ereg_log_sc": ("alpha", P(ol.evidence_reg, s_target=60.0, scale_correct=True, mask=keep),
                        lambda: (rg.EvidenceReg(60.0, scale_correct=True), dict(mask=keep), False)),
        "ereg_one_sided": ("alpha", P(ol.evidence_reg, s_target=60.0, mode="one_sided", margin=0.2, ignore_index=(0, 3), target=target),
                           lambda: (rg.EvidenceReg(60.0, mode="one_sided", margin=0.2, ignore_index=(0, 3)), dict(target=target), False)),
        "ereg_l2": ("alpha", P(ol.evidence_reg, s_target=60.0, mode="l2"),
                    lambda: (rg.EvidenceReg(60.0, mode="l2"), dict(), False)),
        "klw": ("alpha", P(ol.kl_offclasses_conf_weighted, target=target, ignore_index=0, gamma=1.0),
                lambda: (rg.KL_offClasses_to_uniform(ignore_index=0, with_conf_weighting=True, gamma=1.0), dict(target=target), True)),
        "klw_g2": ("alpha", P(ol.kl_offclasses_conf_weighted, target=target, ignore_index=0, gamma=2.0),
                   lambda: (rg.KL_offClasses_to_uniform(ignore_index=0, with_conf_weighting=True, gamma=2.0), dict(target=target), True)),
        "logit": ("logits", P(ol.logit_regularizer),
                  lambda: (rg.LogitRegularizer(), dict(), False)),
        "logit_thr_target": ("logits", P(ol.logit_regularizer, threshold=3.0, ignore_index=0, target=target),
                             lambda: (rg.LogitRegularizer(threshold=3.0, ignore_index=0), dict(target=target), False)),
        "logit_mask": ("logits", P(ol.logit_regularizer, mask=keep),
                       lambda: (rg.LogitRegularizer(threshold=None), dict(mask=keep), False)),
        "logit_target_noignore": ("logits", P(ol.logit_regularizer, threshold=1.0, target=target),
                                  lambda: (rg.LogitRegularizer(threshold=1.0), dict(target=target), False)),
    }


NAMES = ["comp", "comp_trainer", "comp_evid_gate", "comp_attached", "wle", "wle_hard", "wle_nomargin", "band", "band_mask",
         "band_nomask", "ereg_log", "ereg_log_sc", "ereg_one_sided", "ereg_l2", "klw", "klw_g2", "logit", "logit_thr_target",
         "logit_mask", "logit_target_noignore"]
