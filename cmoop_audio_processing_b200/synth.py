"""Synthetic GSC-/BirdCLEF-shaped audio (no dataset is reachable offline).

SURVEY.md section 8(d): class c = three sinusoids at class-specific mel-centre frequencies
with random phase and onset (uniform 0-0.5 s), plus white noise at 10 dB SNR;
``numpy.random.default_rng(seed)``.  Used by tests, bench.py and the CPU baseline alike,
so both sides always see identical inputs.
"""
from __future__ import annotations

import numpy as np


def _mel_to_hz(m):
    f_sp = 200.0 / 3.0
    min_log_mel = 15.0
    logstep = np.log(6.4) / 27.0
    m = np.asarray(m, np.float64)
    return np.where(m >= min_log_mel, 1000.0 * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def class_frequencies(n_classes: int, sample_rate: int = 16000, tones: int = 3, seed: int = 1234) -> np.ndarray:
    """(n_classes, tones) Hz; tone frequencies sit on mel-spaced centres so classes are
    separable in a 40-band log-mel picture."""
    rng = np.random.default_rng(seed)
    top_mel = 15.0 + np.log(0.45 * sample_rate / 1000.0) / (np.log(6.4) / 27.0)
    centres = _mel_to_hz(np.linspace(2.0, top_mel, 4 * max(n_classes, 8)))
    return np.stack([np.sort(rng.choice(centres, size=tones, replace=False)) for _ in range(n_classes)])


def make_clips(n_clips: int, n_classes: int = 12, *, sample_rate: int = 16000, seconds: float = 1.0,
               snr_db: float = 10.0, seed: int = 1234, labels: np.ndarray | None = None,
               dtype=np.float32):
    """Returns (wave [n_clips, n_samples] in [-1,1], labels [n_clips] int64)."""
    rng = np.random.default_rng(seed)
    n = int(round(sample_rate * seconds))
    freqs = class_frequencies(n_classes, sample_rate, seed=seed)
    if labels is None:
        labels = np.arange(n_clips) % n_classes
        rng.shuffle(labels)
    labels = np.asarray(labels, np.int64)
    t = np.arange(n) / sample_rate
    wave = np.empty((n_clips, n), dtype)
    for start in range(0, n_clips, 1024):                      # chunked to bound memory
        sl = slice(start, min(n_clips, start + 1024))
        lab = labels[sl]
        b = len(lab)
        phase = rng.uniform(0, 2 * np.pi, size=(b, freqs.shape[1]))
        onset = rng.uniform(0.0, 0.5, size=(b, 1)) * seconds
        sig = np.zeros((b, n))
        for j in range(freqs.shape[1]):
            sig += np.sin(2 * np.pi * freqs[lab, j][:, None] * t[None, :] + phase[:, j][:, None])
        sig *= (t[None, :] >= onset)
        power = np.maximum((sig ** 2).mean(axis=1, keepdims=True), 1e-12)
        noise = rng.standard_normal((b, n)) * np.sqrt(power / (10 ** (snr_db / 10)))
        x = sig + noise
        x /= np.maximum(np.abs(x).max(axis=1, keepdims=True), 1e-12)
        wave[sl] = x.astype(dtype)
    return wave, labels


def uniform_clips(n_clips: int, n_samples: int = 16000, seed: int = 2, dtype=np.float32) -> np.ndarray:
    """Config 2 throughput input: U(-1,1) samples (value distribution is irrelevant to timing)."""
    rng = np.random.default_rng(seed)
    return rng.uniform(-1.0, 1.0, size=(n_clips, n_samples)).astype(dtype)
