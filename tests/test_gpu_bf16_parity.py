"""GPU parity of the precision the benchmark uses (bf16 tensor-core path, activations stored in bf16) against the exact
fp32 path at BASELINE configs[0] SIZE (12 x 256 training / 12 x 64 validation clips): 16 genotypes -- six of them with
filters = 64 -- x 3 seeds, early stopping as the reference (VERDICT r1 "weak" #1).

Fixed-seed training is a chaotic map: after a few hundred Adam steps two arithmetics that differ in the last bits follow
different trajectories, so the statement is statistical and is given next to the noise floor of the SAME arithmetic under
different seeds.  Measured on B200 (tools/study_bf16.py, profiles/r02_bf16_study_*.json):

    regime                       |d acc| median / p90 / max     |d FPR| median / p95 / max     fp32 seed-to-seed spread
    10 dB SNR, 6 ep  0 / 0 / 0  (all candidates reach 1.0 in both precisions; fronts identical)
    -12 dB, 6 epochs             0.005 / 0.039 / 0.066          0.0004 / 0.0041 / 0.0061       0.123
    -15 dB, 8 epochs             0.012 / 0.050 / 0.370          0.0011 / 0.0100 / 0.0339       0.229
    -18 dB, 10 epochs            0.021 / 0.066 / 0.190          0.0019 / 0.0064 / 0.0174       0.086
(final kernels of round 2, profiles/r02b_bf16_study_*.json; the first set of the round -- other summation orders in the BN
backward sums and the 1x1 projection -- gave medians 0.009 / 0.012 / 0.020, p90 0.041 / 0.054 / 0.062 and maxima 0.148 / 0.181 /
0.195, profiles/r02_bf16_study_*.json: median and p90 are stable across builds, the MAXIMUM over 48 trainings is not -- it is one
candidate that learns an epoch earlier in one arithmetic -- and is of the size of the fp32 seed-to-seed spread.)

Stated tolerance (asserted below with margin): |d acc| median <= 0.03, p90 <= 0.10, max <= 2 x the fp32 seed-to-seed spread of
the same run (and <= 0.5); |d FPR| median <= 0.004, p95 <= 0.015, max <= 0.06; the p90 accuracy gap between precisions stays
below the fp32 seed-to-seed spread; size exact.  Front membership: every pair whose fp32 objectives are separated by more than
the run's maximal gaps keeps its dominance relation in bf16 (a consistency check of the bookkeeping: it follows from the
gaps); at the survey's 10 dB SNR objectives and fronts are equal.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu


def _run(snr_db, epochs):
    import study_bf16 as st
    from cmoop_audio_processing_b200.problem import FitnessProblem, TrainConfig
    xt, yt, xv, yv = st.config0_data(snr_db)
    assert xt.shape == (12 * 256, 49, 40, 1) and xv.shape == (12 * 64, 49, 40, 1)
    rows = {}
    hps = [hp for _ in st.SEEDS for hp in st.GENOTYPES]
    seeds = [s + 7 * i for s in st.SEEDS for i in range(len(st.GENOTYPES))]
    assert sum(hp["filters"] == 64 for hp in st.GENOTYPES) >= 6 and len(st.GENOTYPES) >= 16
    for prec in ("bf16", "fp32"):
        cfg = TrainConfig(variant="B", epochs=epochs, patience=5, restore_best_weights=True, acc_from="evaluate",
                          fpr_mode="filtered", precision=prec)
        prob = FitnessProblem.sa_nsga_local(xt, yt, xv, yv, classes=12, config=cfg)
        rows[prec], _ = prob.train_eval(hps, seeds)
        prob.data.close()
    return st, rows


def test_bf16_vs_fp32_objectives_in_the_noisy_regime():
    st, rows = _run(-15.0, 8)
    n_g, n_s = len(st.GENOTYPES), len(st.SEEDS)
    np.testing.assert_array_equal(rows["bf16"][:, 1], rows["fp32"][:, 1])                  # size objective exact
    acc = {p: rows[p][:, 0].reshape(n_s, n_g) for p in rows}
    fpr = {p: rows[p][:, 2].reshape(n_s, n_g) for p in rows}
    assert 0.25 < acc["fp32"].mean() < 0.97                                                # not a saturated task
    d_acc, d_fpr = np.abs(acc["bf16"] - acc["fp32"]).ravel(), np.abs(fpr["bf16"] - fpr["fp32"]).ravel()
    seed_spread = np.abs(acc["fp32"] - acc["fp32"].mean(axis=0)).max()                     # same arithmetic, other seeds
    assert np.median(d_acc) <= 0.03 and np.percentile(d_acc, 90) <= 0.10 and d_acc.max() <= min(0.5, 2.0 * seed_spread)
    assert np.median(d_fpr) <= 0.004 and np.percentile(d_fpr, 95) <= 0.015 and d_fpr.max() <= 0.06
    assert np.percentile(d_acc, 90) <= seed_spread
    # dominance relations separated by more than the stated tolerance survive the change of precision
    for s in range(n_s):
        a = np.stack([-acc["fp32"][s], rows["fp32"][s * n_g:(s + 1) * n_g, 1], fpr["fp32"][s]], axis=1)
        b = np.stack([-acc["bf16"][s], rows["bf16"][s * n_g:(s + 1) * n_g, 1], fpr["bf16"][s]], axis=1)
        robust, broken = st.robust_dominance_agreement(a, b, (d_acc.max() + 1e-9, 0.0, d_fpr.max() + 1e-9))
        assert broken == 0, (s, robust, broken)


def test_bf16_and_fp32_fronts_are_identical_on_the_survey_task():
    """SURVEY.md section 8(d) config 1 data (10 dB SNR): both precisions drive every genotype to the same accuracy / FPR
    within 0.01 and the penalised non-dominated fronts (reference order included) are equal."""
    from cmoop_audio_processing_b200 import nsga
    st, rows = _run(10.0, 6)
    n_g = len(st.GENOTYPES)
    assert np.abs(rows["bf16"][:, 0] - rows["fp32"][:, 0]).max() <= 0.01
    assert np.abs(rows["bf16"][:, 2] - rows["fp32"][:, 2]).max() <= 0.002
    for s in range(len(st.SEEDS)):
        recs = {}
        for p in rows:
            r = rows[p][s * n_g:(s + 1) * n_g]
            # objectives rounded at the stated tolerance, so that equal-within-tolerance objectives compare equal
            recs[p] = [{"hparams": hp, "objs": [-round(float(r[i, 0]), 2), float(r[i, 1]), round(float(r[i, 2]), 2)],
                        "CV": max(0.0, 0.9 - round(float(r[i, 0]), 2)) + max(0.0, float(r[i, 1]) - 2.5)}
                       for i, hp in enumerate(st.GENOTYPES)]
        for lam in (1.0, 25.5, 50.0):
            assert nsga.fast_non_dominated_sort(recs["bf16"], lam) == nsga.fast_non_dominated_sort(recs["fp32"], lam)
