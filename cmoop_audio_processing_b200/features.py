"""Log-mel / MFCC front-end (host mirror of cmoop_mfcc_*).

The reference loads pre-computed features shaped (N, T, F) and standardises them per
feature (prepare_dataset, nsga_penalty.py:85-155); this module produces that tensor on
the GPU from raw waveforms.  Spec = oracle/mfcc_ref.py (frames of 640 @ hop 320, no
padding, periodic Hann, 1024-point power spectrum, 40 Slaney mel bands, 10*log10,
orthonormal DCT-II).  Torch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib


@dataclass(frozen=True)
class MfccConfig:
    sample_rate: int = 16000
    frame_length: int = 640
    hop: int = 320
    n_fft: int = 1024
    n_mels: int = 40
    n_mfcc: int = 40          # 0 -> log-mel output
    f_min: float = 0.0
    f_max: float = 8000.0
    log_floor: float = 1e-10
    center: bool = False      # librosa-style reflect-padded centring (frame_length must equal n_fft)


class MfccFrontEnd:
    def __init__(self, config: MfccConfig = MfccConfig()):
        self._lib = _lib.load()
        _lib.bind_device()
        self.config = config
        c = _lib.MfccConfig(config.sample_rate, config.frame_length, config.hop, config.n_fft, config.n_mels,
                            config.n_mfcc, config.f_min, config.f_max, config.log_floor, int(config.center))
        handle = C.c_void_p()
        _lib.check(self._lib.cmoop_mfcc_create(C.byref(c), C.byref(handle)), "cmoop_mfcc_create")
        self._handle = handle
        self.n_out = int(self._lib.cmoop_mfcc_n_out(handle))

    def n_frames(self, n_samples: int) -> int:
        return int(self._lib.cmoop_mfcc_n_frames(self._handle, int(n_samples)))

    def set_standardise(self, mean=None, scale=None):
        """Fuse (x - mean[f]) / scale[f] into the kernel epilogue (StandardScaler statistics)."""
        if mean is None:
            _lib.check(self._lib.cmoop_mfcc_set_standardise(self._handle, None, None), "cmoop_mfcc_set_standardise")
            return
        m = np.ascontiguousarray(mean, np.float32)
        s = np.ascontiguousarray(scale, np.float32)
        if m.shape != (self.n_out,) or s.shape != (self.n_out,):
            raise ValueError(f"mean/scale must have shape ({self.n_out},)")
        _lib.check(self._lib.cmoop_mfcc_set_standardise(self._handle, _lib.ptr(m), _lib.ptr(s)),
                   "cmoop_mfcc_set_standardise")

    def __call__(self, wave, out=None):
        """wave: (n_clips, n_samples) float32, or int16 PCM (x = sample / 32768) -- a CUDA torch tensor (stream-ordered, no sync,
        returns a CUDA tensor) or a host numpy array / CPU tensor (chunked overlapped copies,
        returns numpy)."""
        try:
            import torch
        except ImportError:  # pragma: no cover
            torch = None
        if torch is not None and isinstance(wave, torch.Tensor) and wave.is_cuda:
            if wave.dtype not in (torch.float32, torch.int16) or wave.dim() != 2:
                raise ValueError("wave must be a 2-D float32 (or int16 PCM) tensor")
            wave = wave.contiguous()
            n_clips, n_samples = wave.shape
            frames = self.n_frames(n_samples)
            if out is None:
                out = torch.empty((n_clips, frames, self.n_out), dtype=torch.float32, device=wave.device)
            stream = torch.cuda.current_stream(wave.device).cuda_stream
            fwd = self._lib.cmoop_mfcc_fwd_dev if wave.dtype == torch.float32 else self._lib.cmoop_mfcc_fwd_dev_i16
            with torch.cuda.device(wave.device):
                _lib.check(fwd(self._handle, C.c_void_p(wave.data_ptr()), n_clips, n_samples,
                               C.c_void_p(out.data_ptr()), C.c_void_p(stream)), "cmoop_mfcc_fwd_dev")
            return out
        if torch is not None and isinstance(wave, torch.Tensor):
            wave = wave.numpy()
        pcm = np.asarray(wave).dtype == np.int16                   # 16-bit PCM: x = sample / 32768, half the PCIe bytes
        wave = np.ascontiguousarray(wave, np.int16 if pcm else np.float32)
        if wave.ndim != 2:
            raise ValueError("wave must be 2-D (n_clips, n_samples)")
        n_clips, n_samples = wave.shape
        frames = self.n_frames(n_samples)
        if out is None:
            out = np.empty((n_clips, frames, self.n_out), np.float32)
        fwd = self._lib.cmoop_mfcc_fwd_host_i16 if pcm else self._lib.cmoop_mfcc_fwd_host
        _lib.check(fwd(self._handle, _lib.ptr(wave), n_clips, n_samples, _lib.ptr(out)), "cmoop_mfcc_fwd_host")
        return out

    def close(self):
        if getattr(self, "_handle", None):
            self._lib.cmoop_mfcc_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def feature_stats(feats):
    """StandardScaler statistics of a CUDA feature tensor (..., F): per-feature mean and population variance over all
    leading axes (prepare_dataset, nsga_penalty.py:102-114), computed on the device in fp64.  Returns numpy float64."""
    import torch
    lib = _lib.load()
    if not (isinstance(feats, torch.Tensor) and feats.is_cuda and feats.dtype == torch.float32):
        raise ValueError("feats must be a float32 CUDA tensor")
    flat = feats.contiguous().reshape(-1, feats.shape[-1])
    mean, var = np.empty(flat.shape[1], np.float64), np.empty(flat.shape[1], np.float64)
    with torch.cuda.device(flat.device):
        _lib.check(lib.cmoop_feature_stats_dev(C.c_void_p(flat.data_ptr()), flat.shape[0], flat.shape[1], _lib.ptr(mean),
                                               _lib.ptr(var), C.c_void_p(torch.cuda.current_stream(flat.device).cuda_stream)),
                   "cmoop_feature_stats_dev")
    return mean, var


def prepare_dataset_device(front_end: "MfccFrontEnd", splits, policy: str = "fit_train"):
    """Waveforms -> standardised features without leaving the GPU.

    ``splits``: CUDA waveform tensors (float32 or int16 PCM), training split first.  ``policy`` is the reference's scaler
    handling: ``"fit_train"`` -- fit on the training split, transform the others
    (mobo_penalty.py:69-79, sa_nsga_local.py:52-60); ``"fit_each"`` -- fit_transform every split separately
    (nsga_penalty.py:111,124,137); ``"none"`` (sa_nsga_penalty.py).  The statistics come from one un-scaled pass
    (cmoop_feature_stats_dev), the scaled features from a second pass with the scaler fused into the kernel epilogue.
    Returns the list of (N, T, F) CUDA tensors, ready for ``FitnessProblem`` / ``CnnDataset``.
    """
    if policy not in ("fit_train", "fit_each", "none"):
        raise ValueError("policy must be 'fit_train', 'fit_each' or 'none'")
    front_end.set_standardise(None, None)
    if policy == "none":
        return [front_end(w) for w in splits]
    out, stats = [], None
    for i, w in enumerate(splits):
        if policy == "fit_each" or i == 0:
            raw = front_end(w)
            mean, var = feature_stats(raw)
            del raw
            scale = np.sqrt(var)
            scale[scale == 0.0] = 1.0                                   # sklearn _handle_zeros_in_scale
            stats = (mean.astype(np.float32), scale.astype(np.float32))
        front_end.set_standardise(*stats)
        out.append(front_end(w))
        front_end.set_standardise(None, None)
    return out


def mfcc(wave, config: MfccConfig = MfccConfig()):
    """One-shot convenience wrapper."""
    fe = MfccFrontEnd(config)
    try:
        return fe(wave)
    finally:
        fe.close()
