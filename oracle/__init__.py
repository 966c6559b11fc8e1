"""oracle/ -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in numpy / CPU torch, the algorithms of
kav-institute/SemanticLiDARUnc that the CUDA library replaces.  It exists so
the CUDA path can be checked on a GPU box where /root/reference is absent.

Rules (enforced by tests/test_layout.py):
  * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
    `--impl reference` legs may import it -- as the checker or the timed CPU
    baseline, never as the product;
  * the product package `semanticlidarunc_b200` never imports it and has no
    CPU fallback: it raises if libslu.so is missing.

Pinning.  The reference ships no tests, golden vectors or fixtures for this
path, so the oracle is pinned against OUTPUTS OF THE
REFERENCE ITSELF: oracle/gen_golden.py imports the unmodified reference from
/root/reference (through oracle/_refshim.py), runs it on seeded inputs and
commits inputs' seeds + outputs under tests/golden/.  tests/test_oracle_golden.py
checks every oracle function against those vectors on CPU; when the reference
tree is present the same test also re-runs the reference live.

Third-party arithmetic the reference relies on (not under /root/reference):
numpy 2.3.5 (sqrt, arctan2, linspace, digitize, argsort, histogram) and
torch 2.11.0 CPU (softmax, log, digamma, lgamma, softplus, bincount).  The
reference pins torch==2.4.1 / numpy==1.26.4 (docker/Dockerfile:195,
docker/requirements.txt:8); the vectors were produced with the versions
installed in this image, recorded in tests/golden/MANIFEST.json.
"""
