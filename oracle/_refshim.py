"""Import shim for the UNMODIFIED reference (test infrastructure only).

Only usable inside the build container, where /root/reference exists.  It is
used by oracle/gen_golden.py (to write tests/golden/*) and by the optional
`-m "not gpu"` cross-checks that skip when the reference is absent.  Nothing in
the product package, bench.py's timed legs or the `-m gpu` tests imports it.

The reference imports matplotlib / seaborn / open3d at module top for plotting
only (src/dataset/utils.py:69-74, src/metrics/ece.py:5-11,
src/models/evaluator.py:188-189); they are absent here, so they are stubbed.
"""
import os
import sys
from unittest.mock import MagicMock

REF_ROOT = os.environ.get("SLU_REFERENCE_ROOT", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "src")


def available() -> bool:
    return os.path.isdir(REF_SRC)


def install():
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_SRC)
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker",
              "matplotlib.patheffects", "matplotlib.colors", "matplotlib.cm",
              "matplotlib.patches", "matplotlib.lines", "matplotlib.gridspec",
              "seaborn", "open3d"):
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = MagicMock()
    plt = sys.modules["matplotlib.pyplot"]
    if isinstance(plt, MagicMock):
        plt.subplots.return_value = (MagicMock(), MagicMock())
        sys.modules["matplotlib"].pyplot = plt
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
