"""fp64 NumPy definition of the log-mel / MFCC front-end.  PARITY UNPINNED.

TEST INFRASTRUCTURE ONLY.  The reference contains *no* feature-extraction code:
every script loads pre-computed features (nsga_penalty.py:64-71 ``.npy``,
sa_nsga_penalty.py:42-63 ``.h5``) and ``librosa==0.11.0`` (requirements.txt:80) is
never imported.  The only constraints are BASELINE.json's "1 s 16 kHz clips,
40 mel x 49 frames" and the loaders' layout ``(N, T, F)`` with features last
(nsga_penalty.py:104-114).  This file therefore *defines* the spec the CUDA
kernel is held to (1e-4 relative, BASELINE.json north_star):

  frames      t = 0..T-1, samples [t*hop, t*hop + frame_length), no centring / padding
              (T = 1 + (n - frame_length)//hop = 49 for 16000/640/320)
  window      periodic Hann of frame_length
  spectrum    |rfft(frame * window, n_fft)|^2  (frame zero-padded at the END to n_fft;
              this is the tf.signal.stft / speech-commands convention)
  mel         n_mels triangular filters, Slaney scale + Slaney area normalisation,
              f_min..f_max  (librosa.filters.mel defaults, restated below)
  log         10*log10(max(S, 1e-10))     (librosa.power_to_db, ref=1, no top_db)
  mfcc        orthonormal DCT-II over the mel axis, first n_mfcc coefficients
              (librosa.feature.mfcc / scipy.fft.dct(type=2, norm='ortho'))

Cross-checked on CPU against torch.fft + torchaudio.functional.melscale_fbanks
(mel_scale='slaney', norm='slaney') and scipy.fft.dct in tests/test_oracle_mfcc.py.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class MfccSpec:
    sample_rate: int = 16000
    frame_length: int = 640
    hop: int = 320
    n_fft: int = 1024
    n_mels: int = 40
    n_mfcc: int = 40          # 0 -> return log-mel only
    f_min: float = 0.0
    f_max: float = 8000.0
    log_floor: float = 1e-10
    center: bool = False      # librosa stft(center=True): reflect-pad n_fft//2 on both sides (frame_length == n_fft)

    def n_frames(self, n_samples: int) -> int:
        if self.center:
            return 1 + n_samples // self.hop if n_samples > self.n_fft // 2 else 0
        if n_samples < self.frame_length:
            return 0
        return 1 + (n_samples - self.frame_length) // self.hop


def hz_to_mel_slaney(f):
    f = np.asarray(f, np.float64)
    f_sp = 200.0 / 3.0
    mel = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore", invalid="ignore"):
        hi = min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep
    return np.where(f >= min_log_hz, hi, mel)


def mel_to_hz_slaney(m):
    m = np.asarray(m, np.float64)
    f_sp = 200.0 / 3.0
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(spec: MfccSpec) -> np.ndarray:
    """(n_mels, n_fft//2 + 1) fp64 weights."""
    n_bins = spec.n_fft // 2 + 1
    fft_freqs = np.linspace(0.0, spec.sample_rate / 2.0, n_bins)
    mel_pts = np.linspace(hz_to_mel_slaney(spec.f_min), hz_to_mel_slaney(spec.f_max), spec.n_mels + 2)
    hz_pts = mel_to_hz_slaney(mel_pts)
    width = np.diff(hz_pts)
    ramps = hz_pts[:, None] - fft_freqs[None, :]
    fb = np.zeros((spec.n_mels, n_bins))
    for i in range(spec.n_mels):
        rising = -ramps[i] / width[i]
        falling = ramps[i + 2] / width[i + 1]
        fb[i] = np.maximum(0.0, np.minimum(rising, falling))
    fb *= (2.0 / (hz_pts[2:spec.n_mels + 2] - hz_pts[:spec.n_mels]))[:, None]
    return fb


def hann_periodic(n: int) -> np.ndarray:
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def dct_matrix(n_mels: int, n_mfcc: int) -> np.ndarray:
    """(n_mfcc, n_mels) orthonormal DCT-II."""
    b = np.arange(n_mels)
    c = np.arange(n_mfcc)[:, None]
    mat = np.cos(np.pi * c * (2 * b + 1) / (2.0 * n_mels)) * np.sqrt(2.0 / n_mels)
    mat[0] *= np.sqrt(0.5)
    return mat


def power_spectrogram(wave: np.ndarray, spec: MfccSpec) -> np.ndarray:
    """(B, T, n_fft//2+1) fp64."""
    wave = np.atleast_2d(np.asarray(wave, np.float64))
    t = spec.n_frames(wave.shape[1])
    if spec.center:
        assert spec.frame_length == spec.n_fft
        wave = np.pad(wave, ((0, 0), (spec.n_fft // 2, spec.n_fft // 2)), mode="reflect")
    idx = np.arange(spec.frame_length)[None, :] + spec.hop * np.arange(t)[:, None]
    frames = wave[:, idx] * hann_periodic(spec.frame_length)
    spectrum = np.fft.rfft(frames, n=spec.n_fft, axis=-1)
    return spectrum.real ** 2 + spectrum.imag ** 2


def log_mel(wave: np.ndarray, spec: MfccSpec = MfccSpec()) -> np.ndarray:
    """(B, T, n_mels) fp64, dB."""
    mel = power_spectrogram(wave, spec) @ mel_filterbank(spec).T
    return 10.0 * np.log10(np.maximum(mel, spec.log_floor))


def mfcc(wave: np.ndarray, spec: MfccSpec = MfccSpec()) -> np.ndarray:
    """(B, T, n_mfcc) fp64; returns log-mel when spec.n_mfcc == 0."""
    lm = log_mel(wave, spec)
    if spec.n_mfcc == 0:
        return lm
    return lm @ dct_matrix(spec.n_mels, spec.n_mfcc).T


def standardise(features: np.ndarray, mean: np.ndarray | None = None, scale: np.ndarray | None = None):
    """Per-feature StandardScaler over the last axis (nsga_penalty.py:102-114): fit on
    features.reshape(-1, F) when mean/scale are None.  Returns (scaled, mean, scale)."""
    f = features.shape[-1]
    flat = features.reshape(-1, f)
    if mean is None:
        mean = flat.mean(axis=0)
        var = flat.var(axis=0)
        scale = np.sqrt(var)
        scale = np.where(scale == 0.0, 1.0, scale)   # sklearn _handle_zeros_in_scale
    return ((flat - mean) / scale).reshape(features.shape), mean, scale
