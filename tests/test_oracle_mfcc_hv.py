"""Pins for the two 'parity unpinned' oracles: MFCC spec vs torch/torchaudio/scipy, HV vs
inclusion-exclusion and hand-computed cases.  CPU only."""
import numpy as np
import pytest

from oracle import hv_ref, mfcc_ref
from cmoop_audio_processing_b200 import synth


def test_frame_count_and_shapes():
    spec = mfcc_ref.MfccSpec()
    assert spec.n_frames(16000) == 49
    assert spec.n_frames(639) == 0 and spec.n_frames(640) == 1
    wave, _ = synth.make_clips(3, 12, seed=5)
    assert mfcc_ref.mfcc(wave).shape == (3, 49, 40)
    assert mfcc_ref.log_mel(wave).shape == (3, 49, 40)


def test_mel_filterbank_matches_torchaudio_slaney():
    torchaudio = pytest.importorskip("torchaudio")
    spec = mfcc_ref.MfccSpec()
    fb = mfcc_ref.mel_filterbank(spec)
    ta = torchaudio.functional.melscale_fbanks(n_freqs=513, f_min=0.0, f_max=8000.0, n_mels=40, sample_rate=16000,
                                               norm="slaney", mel_scale="slaney").numpy().T
    np.testing.assert_allclose(fb, ta, rtol=2e-5, atol=1e-7)      # torchaudio builds it in fp32 (weights ~1e-2)


def test_power_spectrum_matches_torch_fft():
    import torch
    spec = mfcc_ref.MfccSpec()
    wave, _ = synth.make_clips(2, 12, seed=9)
    ours = mfcc_ref.power_spectrogram(wave, spec)
    w = torch.from_numpy(wave.astype(np.float64))
    frames = w.unfold(1, spec.frame_length, spec.hop) * torch.hann_window(spec.frame_length, periodic=True,
                                                                         dtype=torch.float64)
    ref = torch.fft.rfft(frames, n=spec.n_fft).abs().pow(2).numpy()
    np.testing.assert_allclose(ours, ref, rtol=1e-10, atol=1e-12)


def test_dct_matches_scipy_ortho():
    from scipy.fft import dct
    rng = np.random.default_rng(1)
    x = rng.standard_normal((7, 40))
    np.testing.assert_allclose(x @ mfcc_ref.dct_matrix(40, 40).T, dct(x, type=2, norm="ortho", axis=-1), atol=1e-12)
    np.testing.assert_allclose(x @ mfcc_ref.dct_matrix(40, 13).T, dct(x, type=2, norm="ortho", axis=-1)[:, :13],
                               atol=1e-12)


def test_log_floor_and_silence():
    spec = mfcc_ref.MfccSpec()
    out = mfcc_ref.log_mel(np.zeros((1, 16000)), spec)
    assert np.all(out == -100.0)


def test_standardise_matches_sklearn():
    from sklearn.preprocessing import StandardScaler
    rng = np.random.default_rng(3)
    feats = rng.standard_normal((5, 49, 40)) * 3 + 1
    ours, mean, scale = mfcc_ref.standardise(feats)
    ref = StandardScaler().fit_transform(feats.reshape(-1, 40)).reshape(feats.shape)
    np.testing.assert_allclose(ours, ref, atol=1e-12)


def test_hv_hand_cases():
    assert hv_ref.hypervolume([[0, 0, 0]], [1, 1, 1]) == 1.0
    assert hv_ref.hypervolume([[0.5, 0.5, 0.5]], [1, 1, 1]) == 0.125
    # two boxes overlapping in a 0.25 cube corner
    assert hv_ref.hypervolume([[0, 0.5, 0.5], [0.5, 0, 0.5]], [1, 1, 1]) == pytest.approx(0.25 + 0.25 - 0.125)
    assert hv_ref.hypervolume([[2, 0, 0]], [1, 1, 1]) == 0.0          # outside the reference box
    assert hv_ref.hypervolume([[0.2, 0.8], [0.5, 0.5], [0.8, 0.2]], [1, 1]) == pytest.approx(0.16 + 0.15 + 0.06)
    assert hv_ref.hypervolume(np.zeros((0, 3)), [1, 1, 1]) == 0.0


def test_hv_matches_inclusion_exclusion_and_monte_carlo():
    rng = np.random.default_rng(4)
    for n in (1, 2, 5, 9, 12):
        pts = rng.random((n, 3))
        if n > 4:
            pts[1] = pts[0]                        # duplicate
            pts[2, 2] = pts[3, 2]                  # tie in z
        ref = hv_ref.reference_point(pts)
        assert hv_ref.hypervolume(pts, ref) == pytest.approx(hv_ref.hypervolume_inclusion_exclusion(pts, ref), rel=1e-12)
    pts = rng.random((60, 3))
    ref = np.array([1.0, 1.0, 1.0])
    samples = rng.random((400000, 3))
    dominated = np.zeros(len(samples), bool)
    for p in pts:
        dominated |= np.all(samples >= p, axis=1)
    assert hv_ref.hypervolume(pts, ref) == pytest.approx(dominated.mean(), abs=4e-3)
