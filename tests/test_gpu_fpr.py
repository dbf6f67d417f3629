"""GPU parity: calculate_fpr through the C ABI on FIXED label / prediction vectors, bit-exact
(==) against the values the reference's own three textual forms produced (tests/golden/fpr.json, made by
oracle/make_golden.py from nsga_penalty.py:351-364, sa_nsga_local.py:138-141, init_sa_nsga_local.py:137-143)."""
import numpy as np
import pytest

from oracle import nsga_ref

pytestmark = pytest.mark.gpu


def test_fpr_three_forms_bit_exact_against_reference_golden(golden):
    from cmoop_audio_processing_b200.problem import calculate_fpr
    cases = golden("fpr")["cases"]
    assert len(cases) >= 20
    for case in cases:
        c, yt, yp = case["classes"], case["y_true"], case["y_pred"]
        got_all, cm = calculate_fpr(yt, yp, c, "all", return_confusion=True)
        assert got_all == case["fpr_all"]
        assert calculate_fpr(yt, yp, c, "filtered") == case["fpr_filtered"]
        assert calculate_fpr(yt, yp, c, "vectorised") == case["fpr_vectorised"]
        np.testing.assert_array_equal(cm, nsga_ref.confusion(yt, yp, c))


def test_fpr_zero_label_quirk_and_edges():
    """nsga_penalty.py:387 feeds argmax(y_val, axis=1) == 0 as y_true: the value collapses to (1 - frac_pred0) / C."""
    from cmoop_audio_processing_b200.problem import calculate_fpr
    rng = np.random.default_rng(7)
    for c, n in ((10, 640), (12, 768), (397, 5000)):
        pred = rng.integers(0, c, n)
        zeros = np.zeros(n, np.int64)
        got = calculate_fpr(zeros, pred, c, "all")
        assert got == nsga_ref.fpr_macro(nsga_ref.confusion(zeros, pred, c), "all")
        assert got == pytest.approx((1 - np.mean(pred == 0)) / c, rel=1e-12)
    # empty input, out-of-range labels (dropped like sklearn's labels=range(C)), single class
    assert calculate_fpr([], [], 5) == 0.0
    assert calculate_fpr([], [], 5, "filtered") == 0.0
    yt, yp = [0, 1, 7, -1, 2], [1, 1, 0, 0, 9]
    assert calculate_fpr(yt, yp, 3) == nsga_ref.fpr_macro(nsga_ref.confusion(yt, yp, 3), "all")
    assert calculate_fpr([0, 0], [0, 0], 1) == 0.0
    # a large random case against the oracle, every mode (numpy pairwise summation order above 128 classes)
    yt, yp = rng.integers(0, 300, 20000), rng.integers(0, 300, 20000)
    cm = nsga_ref.confusion(yt, yp, 300)
    assert calculate_fpr(yt, yp, 300, "all") == nsga_ref.fpr_macro(cm, "all")
    assert calculate_fpr(yt, yp, 300, "filtered") == nsga_ref.fpr_macro(cm, "filtered")
