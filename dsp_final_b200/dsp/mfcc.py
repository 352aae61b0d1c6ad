"""MfccConfig / mel_filterbank / dct_type_2 / log_mel_spectrogram / mfcc with the
reference's signatures (src/dsp/mfcc.py).

log_mel_spectrogram() and mfcc() run the whole chain -- pre-emphasis, framing,
window, FFT, |X|^2, mel projection, log, DCT-II -- in one fused CUDA kernel
(dspx_features) and return float64 ndarrays like the reference; the values are
the GPU's float32 results widened, so FeatureCache's .astype(np.float32)
(src/features/cache.py:74) stores them unchanged.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from functools import lru_cache

import numpy as np

from ..batch import dct2_rows, features_batch


@dataclass
class MfccConfig:
    """Same ten fields, order and defaults as src/dsp/mfcc.py:10-21 (asdict() feeds the cache hash)."""

    sample_rate: int
    frame_length: int
    hop_length: int
    n_fft: int | None = None
    n_mels: int = 40
    n_mfcc: int = 13
    f_min: float = 0.0
    f_max: float | None = None
    pre_emphasis: float = 0.97
    window: str = "hann"


def _hz_to_mel(hz: float) -> float:
    return 2595.0 * math.log10(1.0 + hz / 700.0)


def _mel_to_hz(mel: float) -> float:
    return 700.0 * (10 ** (mel / 2595.0) - 1.0)


def mel_filterbank(n_mels: int, n_fft: int, sample_rate: int, f_min: float = 0.0,
                   f_max: float | None = None) -> np.ndarray:
    """HTK-mel, floor-binned, unnormalised triangles in float64 (src/dsp/mfcc.py:32-58).

    Host-side constant table; the plan inside libdspx builds the same table in C++
    (csrc/tables.cuh) and tests compare the two.
    """
    if f_max is None:
        f_max = sample_rate / 2
    mels = np.linspace(_hz_to_mel(f_min), _hz_to_mel(f_max), n_mels + 2)
    edges = np.floor((n_fft + 1) * np.array([_mel_to_hz(m) for m in mels]) / sample_rate).astype(int)
    n_bins = n_fft // 2 + 1
    bank = np.zeros((n_mels, n_bins), dtype=np.float64)
    bins = np.arange(n_bins)
    for row, (lo, mid, hi) in enumerate(zip(edges[:-2], edges[1:-1], edges[2:])):
        if hi <= lo:
            continue
        rise = (bins >= lo) & (bins < mid)
        fall = (bins >= mid) & (bins < hi)
        bank[row, rise] = (bins[rise] - lo) / max(1, mid - lo)
        bank[row, fall] = (hi - bins[fall]) / max(1, hi - mid)
    return bank


@lru_cache(maxsize=128)
def _mel_filterbank_cached(n_mels: int, n_fft: int, sample_rate: int, f_min: float, f_max: float | None) -> np.ndarray:
    return mel_filterbank(n_mels, n_fft, sample_rate, f_min, f_max)


@lru_cache(maxsize=64)
def _dct_basis(n_mfcc: int, n: int) -> np.ndarray:
    """cos(pi/n (j + 0.5) k), float64 (src/dsp/mfcc.py:79-83)."""
    k = np.arange(n_mfcc)[:, None]
    j = np.arange(n)[None, :]
    return np.cos(math.pi / n * (j + 0.5) * k)


def dct_type_2(x: np.ndarray, n_mfcc: int) -> np.ndarray:
    """Un-normalised DCT-II times two over the last axis (src/dsp/mfcc.py:73-76), on the GPU."""
    x = np.asarray(x)
    return dct2_rows(x.astype(np.float32), n_mfcc).astype(np.float64)


def _clip(signal) -> np.ndarray:
    sig = np.asarray(signal).reshape(-1)
    if sig.dtype != np.float32:
        # the reference evaluates pre-emphasis in the input dtype; the kernel is float32
        sig = sig.astype(np.float32)
    return sig


def log_mel_spectrogram(signal: np.ndarray, cfg: MfccConfig) -> np.ndarray:
    """[n_frames, n_mels] float64 (src/dsp/mfcc.py:86-103)."""
    return features_batch(_clip(signal)[None, :], cfg, ("log_mel",))["log_mel"][0].astype(np.float64)


def mfcc(signal: np.ndarray, cfg: MfccConfig) -> np.ndarray:
    """[n_frames, n_mfcc] float64 (src/dsp/mfcc.py:106-109)."""
    return features_batch(_clip(signal)[None, :], cfg, ("mfcc",))["mfcc"][0].astype(np.float64)
