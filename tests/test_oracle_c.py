"""The plain-C oracle (CPU baseline) against the NumPy oracle and the reference-made goldens.  CPU only."""
import numpy as np

from conftest import records
from oracle import build_c, mfcc_ref, nsga_ref
from cmoop_audio_processing_b200 import synth


def test_c_mfcc_matches_numpy_oracle():
    wave, _ = synth.make_clips(9, 12, seed=11)
    for n_mfcc in (40, 0, 13):
        got = build_c.mfcc(wave, n_mfcc=n_mfcc)
        want = mfcc_ref.mfcc(wave, mfcc_ref.MfccSpec(n_mfcc=n_mfcc))
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-8)
    got = build_c.mfcc(wave, frame_length=400, hop=160, n_mfcc=13)
    want = mfcc_ref.mfcc(wave, mfcc_ref.MfccSpec(frame_length=400, hop=160, n_mfcc=13))
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-8)


def test_c_nds_and_crowding_match_reference_golden(golden):
    for case in golden("nsga")["cases"]:
        objs, cv = np.array(case["objs"]), np.array(case["cv"])
        fronts, _ = build_c.nds(objs, cv, case["lam"])
        assert fronts == case["fronts"]
        for f, gt, lt in zip(fronts, case["crowd_gt"], case["crowd_lt"]):
            assert build_c.crowding(objs, f, mode=0).tolist() == gt
            assert build_c.crowding(objs, f, mode=1).tolist() == lt


def test_c_nds_matches_python_oracle_large():
    rng = np.random.default_rng(3)
    objs = rng.random((300, 3))
    cv = rng.random(300) * (rng.random(300) < 0.5)
    fronts, _ = build_c.nds(objs, cv, 9.5)
    assert fronts == nsga_ref.fast_non_dominated_sort(records(objs.tolist(), cv.tolist()), 9.5)
