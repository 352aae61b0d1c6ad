"""fft / ifft / rfft with the reference's signatures (src/dsp/fft.py:27-77).

Same contract as the reference: 1-D input, optional n (truncate or zero-pad to n,
then zero-pad to the next power of two, fft.py:32-42), complex128 ndarray out.
The transform runs on the GPU in complex64 (dspx_fft_c2c, a Stockham autosort
FFT; no bit-reversal table is ever built) and is widened on return.
"""
from __future__ import annotations

from typing import Iterable

import numpy as np

from ..batch import fft_batch


def _next_pow_two(n: int) -> int:
    """src/dsp/fft.py:7-10."""
    return 1 if n <= 1 else 1 << (int(n) - 1).bit_length()


def _as_1d(x, dtype) -> np.ndarray:
    a = np.asarray(x if isinstance(x, np.ndarray) else list(x), dtype=dtype)
    return a.reshape(-1)


def fft(x: Iterable[complex], n: int | None = None) -> np.ndarray:
    """Forward DFT, length rounded up to a power of two (src/dsp/fft.py:27-61)."""
    a = _as_1d(x, np.complex128)
    if n is None:
        n = a.shape[0]
    return fft_batch(a, n=n).astype(np.complex128)


def ifft(x: Iterable[complex], n: int | None = None) -> np.ndarray:
    """conj(fft(conj x)) / N (src/dsp/fft.py:64-70)."""
    a = _as_1d(x, np.complex128)
    if n is None:
        n = a.shape[0]
    return fft_batch(a, n=n, inverse=True).astype(np.complex128)


def rfft(x: Iterable[float], n: int | None = None) -> np.ndarray:
    """First N/2+1 bins of the complex transform of a real signal (src/dsp/fft.py:73-77)."""
    a = _as_1d(x, np.float64)
    if n is None:
        n = a.shape[0]
    y = fft_batch(a, n=n).astype(np.complex128)
    return y[: y.shape[0] // 2 + 1]
