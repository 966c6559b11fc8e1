"""The exactness argument of the projection's cross-product decision (csrc/slu_project.cu, SLU_PROJECT_CROSS), checked on
the CPU: outside the 1e-14 band the sign of the float64 cross product gives the bin count np.digitize(arctan2) gives, for
points from 1e-17 to 1e-5 rad around column and row edges, float64 and float32 coordinates (tools/emulate_cross_decision.py)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import emulate_cross_decision as emu  # noqa: E402


@pytest.mark.parametrize("f32", [False, True])
def test_cross_product_decision_agrees_with_digitize_outside_the_band(f32):
    n = 300_000
    bad_c, band_c = emu.columns(n, f32=f32, seed=7)
    bad_r, band_r = emu.rows(n, f32=f32, seed=8)
    assert bad_c == 0 and bad_r == 0
    if not f32:                                   # arbitrary float64 coordinates reach the band: ~3/12 of the log-uniform offsets
        assert 0.15 * n < band_c < 0.35 * n and 0.15 * n < band_r < 0.35 * n
    else:                                         # float32 coordinates are never that close to an edge
        assert band_c < 50 and band_r < 50
