"""GPU integration: the whole hot path chained the way the reference drivers use it --
synthetic GSC-shaped waveforms -> CUDA log-mel front-end -> per-feature standardisation -> population CNN
train/score -> GP surrogate + local search -> NDS / crowding truncation -> HV / IGD / Spread
(BASELINE configs[0] and [2] at toy size)."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kws_features():
    from cmoop_audio_processing_b200 import synth
    from cmoop_audio_processing_b200.features import MfccConfig, MfccFrontEnd
    from oracle import mfcc_ref
    n_cls = 12
    wt, yt = synth.make_clips(n_cls * 32, n_cls, seed=1234)
    wv, yv = synth.make_clips(n_cls * 16, n_cls, seed=4321)
    fe = MfccFrontEnd(MfccConfig(n_mfcc=0))                        # log-mel, 49 x 40 (KWS_10_log_mel_3000 in the reference)
    xt = fe(wt)
    mean = xt.reshape(-1, 40).mean(axis=0)
    scale = xt.reshape(-1, 40).std(axis=0)
    fe.set_standardise(mean, scale)                                # StandardScaler fitted on train (sa_nsga_local.py:52-60)
    xt_s, xv_s = fe(wt), fe(wv)
    ref, _, _ = mfcc_ref.standardise(mfcc_ref.log_mel(wt[:8], mfcc_ref.MfccSpec(n_mfcc=0)), mean.astype(np.float64),
                                     scale.astype(np.float64))
    assert np.abs(xt_s[:8] - ref).max() < 5e-4
    return xt_s[..., None], yt, xv_s[..., None], yv, n_cls


def test_nsga2_generation_loop(kws_features):
    from cmoop_audio_processing_b200 import drivers
    from cmoop_audio_processing_b200.problem import FitnessProblem, TrainConfig
    xt, yt, xv, yv, n_cls = kws_features
    cfg = TrainConfig(variant="A", epochs=3, patience=3, acc_from="history", y_true_mode="argmax_quirk", precision="bf16")
    prob = FitnessProblem.nsga_penalty(xt, yt, xv, yv, classes=n_cls, config=cfg)
    random.seed(0)
    pareto, history, timings = drivers.nsga2(8, 2, drivers.default_ops(prob, surrogate=False))
    assert len(history) == 2 and all(len(h) == 8 for h in history)
    assert prob.evaluations == 8 * 3
    accs = [-r["objs"][0] for r in history[-1]]
    assert all(0.0 <= a <= 1.0 for a in accs)
    # training really optimises: the epoch-mean training loss falls for every candidate (validation accuracy is not
    # asserted: with Keras' BN momentum 0.99 the moving statistics lag far behind after a few dozen steps)
    _, hist = prob.train_eval([r["hparams"] for r in history[-1][:4]], [1, 2, 3, 4], want_history=True)
    assert np.all(hist[:, 2, 0] < hist[:, 0, 0])
    for r in pareto:
        assert r["CV"] == 0
    ind = drivers.front_indicators(history[-1])
    assert ind["hv"] >= 0.0


def test_sa_nsga2_with_surrogate_and_local_search(kws_features):
    from cmoop_audio_processing_b200 import drivers, quality
    from cmoop_audio_processing_b200.problem import FitnessProblem, TrainConfig
    xt, yt, xv, yv, n_cls = kws_features
    cfg = TrainConfig(variant="B", epochs=3, patience=3, restore_best_weights=True, acc_from="evaluate",
                      fpr_mode="filtered", precision="bf16")
    prob = FitnessProblem.sa_nsga_local(xt, yt, xv, yv, classes=n_cls, config=cfg)
    prob.min_accuracy = 0.5
    random.seed(1)
    np.random.seed(1)
    pareto, history, timings = drivers.sa_nsga2(8, 2, 0.334, drivers.default_ops(prob), local_search=True)
    assert prob.evaluations == 8 + 2 * max(1, int(8 * 0.334))
    assert all(t["true_evals"] == 2 for t in timings)
    pts = np.array([r["objs"] for r in history[-1]])
    front = pts[quality.nondominated_mask(pts)]
    ref = quality.reference_point(pts)
    hv = quality.hypervolume(front, ref)
    assert hv > 0.0
    m = quality.front_metrics(front, front)
    assert m["gd"] == 0.0 and m["igd"] == 0.0


def test_run_mobo_loop(kws_features):
    """mobo_penalty.py flow: 4 GPs (Matern 2.5, normalize_y) on the GPU posterior, 500 random candidates, penalised sum."""
    from cmoop_audio_processing_b200 import drivers
    from cmoop_audio_processing_b200.problem import FitnessProblem, TrainConfig
    xt, yt, xv, yv, n_cls = kws_features
    cfg = TrainConfig(variant="A", epochs=2, patience=2, restore_best_weights=True, acc_from="history", precision="bf16")
    prob = FitnessProblem.mobo_penalty(xt, yt, xv, yv, classes=n_cls, config=cfg)
    random.seed(2)
    np.random.seed(2)
    pareto, (x_vec, y_objs, y_cv) = drivers.run_mobo(5, 3, 500, prob)
    assert x_vec.shape == (8, 6) and y_objs.shape == (8, 3) and y_cv.shape == (8, 1)
    assert prob.evaluations == 8
    assert np.all((x_vec >= 0) & (x_vec <= 1))
    for hp, objs, cv in pareto:
        assert cv <= 1e-8


def test_device_resident_dataset_staging():
    """waveforms -> MFCC -> StandardScaler statistics -> standardised features -> CNN dataset without a
    host round trip.  Statistics vs NumPy float64, scaled features vs the oracle for the three scaler policies of the
    reference scripts, and a population evaluated from the device-resident dataset equals the host-staged one bit for bit."""
    import torch
    from cmoop_audio_processing_b200 import synth
    from cmoop_audio_processing_b200.features import MfccFrontEnd, feature_stats, prepare_dataset_device
    from cmoop_audio_processing_b200.problem import FitnessProblem, TrainConfig
    from oracle import mfcc_ref

    wave, labels = synth.make_clips(96 + 48, 12, seed=21)
    w_tr, w_va = torch.from_numpy(wave[:96]).cuda(), torch.from_numpy(wave[96:]).cuda()
    fe = MfccFrontEnd()
    raw = fe(w_tr)
    mean, var = feature_stats(raw)
    flat = raw.cpu().numpy().astype(np.float64).reshape(-1, 40)
    np.testing.assert_allclose(mean, flat.mean(axis=0), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(var, flat.var(axis=0), rtol=1e-11)
    base_tr, base_va = mfcc_ref.mfcc(wave[:96]), mfcc_ref.mfcc(wave[96:])
    want_tr, m, s = mfcc_ref.standardise(base_tr)
    for policy, want_va in (("fit_train", mfcc_ref.standardise(base_va, m, s)[0]), ("fit_each", mfcc_ref.standardise(base_va)[0]),
                            ("none", base_va)):
        x_tr, x_va = prepare_dataset_device(fe, [w_tr, w_va], policy)
        assert x_tr.is_cuda and x_tr.shape == (96, 49, 40)
        assert np.abs(x_va.cpu().numpy() - want_va).max() < (2e-4 if policy != "none" else 2e-3)
        if policy != "none":
            assert np.abs(x_tr.cpu().numpy() - want_tr).max() < 2e-4
    x_tr, x_va = prepare_dataset_device(fe, [w_tr, w_va], "fit_train")
    hps = [dict(filters=16, kernel_size=3, use_bn=True, residual_blocks=1, fc_layers=1, use_dropout=False),
           dict(filters=32, kernel_size=5, use_bn=False, residual_blocks=2, fc_layers=2, use_dropout=True)]
    cfg = TrainConfig(variant="B", epochs=2, patience=2, precision="bf16")
    dev = FitnessProblem(x_tr, labels[:96], x_va, labels[96:], classes=12, config=cfg).train_eval(hps, [3, 4])[0]
    host = FitnessProblem(x_tr.cpu().numpy()[..., None], labels[:96], x_va.cpu().numpy()[..., None], labels[96:], classes=12,
                          config=cfg).train_eval(hps, [3, 4])[0]
    np.testing.assert_array_equal(dev, host)
