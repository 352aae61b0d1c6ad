"""stft / frame_signal / _get_window with the reference's signatures (src/dsp/stft.py).

stft() sends the whole clip to the fused GPU kernel (framing, window, zero-pad or
truncation to n_fft, FFT) and returns complex128 [n_frames, n_fft_pow2 // 2 + 1]
like the reference.  The one deliberate difference: a signal shorter than one
frame raises ValueError, where the reference's as_strided view reads past the
buffer.
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import Literal

import numpy as np

from ..batch import stft_batch

WindowType = Literal["hann", "hamming", "rect"]


@lru_cache(maxsize=32)
def _get_window(window: WindowType, frame_length: int) -> np.ndarray:
    """Periodic window table in float64 (src/dsp/stft.py:12-24). Host-side constant."""
    if frame_length <= 0:
        raise ValueError("frame_length must be positive")
    phase = 2.0 * math.pi * np.arange(frame_length) / frame_length
    if window == "hann":
        return 0.5 - 0.5 * np.cos(phase)
    if window == "hamming":
        return 0.54 - 0.46 * np.cos(phase)
    if window == "rect":
        return np.ones(frame_length)
    raise ValueError(f"Unsupported window: {window}")


def frame_signal(signal: np.ndarray, frame_length: int, hop_length: int) -> np.ndarray:
    """[n_frames, frame_length] float64 copy of the overlapping frames (src/dsp/stft.py:27-40).

    Pure data movement (no arithmetic); kept for callers that want the frames themselves.
    """
    if frame_length <= 0 or hop_length <= 0:
        raise ValueError("frame_length and hop_length must be positive")
    sig = np.asarray(signal).reshape(-1)
    if sig.shape[0] < frame_length:
        raise ValueError(f"signal of {sig.shape[0]} samples is shorter than one frame ({frame_length})")
    view = np.lib.stride_tricks.sliding_window_view(sig, frame_length)[::hop_length]
    return np.array(view, dtype=np.float64, copy=True)


def stft(signal: np.ndarray, frame_length: int, hop_length: int, window: WindowType = "hann",
         n_fft: int | None = None) -> np.ndarray:
    """Short-time Fourier transform, no centring, no tail padding (src/dsp/stft.py:43-56)."""
    if frame_length <= 0 or hop_length <= 0:
        raise ValueError("frame_length and hop_length must be positive")
    if window not in ("hann", "hamming", "rect"):
        raise ValueError(f"Unsupported window: {window}")
    sig = np.asarray(signal).reshape(-1)
    out = stft_batch(sig[None, :], frame_length, hop_length, window=window, n_fft=n_fft)
    return out[0].astype(np.complex128)
