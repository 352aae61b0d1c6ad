"""Synthetic ESC-50-shaped clips.

ESC-50 itself is not shipped with the reference (data/ESC-50-master holds only
.gitkeep) and there is no network, so every parity and throughput run uses
clips made here: 5 s mono at 44.1 kHz (220 500 float32 samples), 50 classes,
5 folds x 400 clips, peak-normalised like src/features/cache.py:67 does.

Two generators with the same recipe:
  * host_clips()    - NumPy, seeded, bit-reproducible: parity sets and goldens;
  * device_clips()  - torch on the GPU, for throughput sets that never touch
                      the host (parity is then checked on a copied-back sample).
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 44_100
CLIP_SECONDS = 5.0
CLIP_LEN = 220_500
N_CLASSES = 50


def class_of(i: int) -> int:
    return i % N_CLASSES


def fold_of(i: int, n_clips: int = 2000) -> int:
    """Folds 1..5 in equal contiguous blocks (ESC-50: 400 clips per fold)."""
    per = max(1, n_clips // 5)
    return min(5, 1 + i // per)


def _class_freqs(c: int) -> np.ndarray:
    # three partials per class, log-spaced over 80 Hz .. 12 kHz
    base = 80.0 * (12_000.0 / 80.0) ** (c / (N_CLASSES - 1) * 0.6)
    return np.array([base, base * 2.0 ** (7.0 / 12.0) * 1.5, base * 3.7])


def host_clip(i: int, seed: int, length: int = CLIP_LEN, sample_rate: int = SAMPLE_RATE) -> np.ndarray:
    """One float32 clip; depends only on (i, seed, length, sample_rate)."""
    rng = np.random.default_rng([seed, i])
    c = class_of(i)
    t = np.arange(length, dtype=np.float64) / sample_rate
    x = np.zeros(length, np.float64)
    freqs = np.minimum(_class_freqs(c), 0.45 * sample_rate)
    amps = np.array([1.0, 0.5, 0.25]) * rng.uniform(0.5, 1.0, 3)
    for f, a in zip(freqs, amps):
        x += a * np.sin(2.0 * np.pi * f * (1.0 + 0.01 * rng.standard_normal()) * t + rng.uniform(0, 2 * np.pi))
    # coloured noise: one-pole tilt that depends on the class
    white = rng.standard_normal(length)
    pole = 0.2 + 0.75 * (c / (N_CLASSES - 1))
    # one-pole response truncated to a 64-tap geometric FIR (deterministic, no scipy)
    taps = pole ** np.arange(64)
    col = np.convolve(white, taps)[:length]
    x += (0.1 + 0.4 * ((c * 7) % N_CLASSES) / N_CLASSES) * col / np.sqrt(np.sum(taps ** 2))
    x += 1e-3 * rng.standard_normal(length)
    if rng.uniform() < 0.25:                      # digital-silence tail (real ESC-50 padding)
        cut = int(rng.uniform(0.3, 0.95) * length)
        x[cut:] = 0.0
    peak = np.max(np.abs(x))
    if peak > 0:
        x = x / peak
    return x.astype(np.float32)


def host_clips(n: int, seed: int, length: int = CLIP_LEN, sample_rate: int = SAMPLE_RATE,
               first: int = 0) -> np.ndarray:
    """float32 [n, length]; clip j is host_clip(first + j, seed)."""
    out = np.empty((n, length), np.float32)
    for j in range(n):
        out[j] = host_clip(first + j, seed, length, sample_rate)
    return out


def labels(n: int, first: int = 0) -> np.ndarray:
    return np.array([class_of(first + j) for j in range(n)], dtype=np.int32)


def device_clips(n: int, seed: int, device, length: int = CLIP_LEN, sample_rate: int = SAMPLE_RATE,
                 first: int = 0, chunk: int = 256):
    """float32 [n, length] generated on `device` (same recipe family, torch RNG).

    Not bit-identical to host_clips(); throughput sets only.  Parity on such a
    set is checked by copying a sample of clips back and running the oracle on
    exactly those samples.
    """
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed) * 1_000_003 + int(first))
    out = torch.empty((n, length), dtype=torch.float32, device=device)
    t = torch.arange(length, device=device, dtype=torch.float32) / float(sample_rate)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        m = e - s
        cls = (torch.arange(first + s, first + e, device=device) % N_CLASSES).to(torch.float32)
        base = 80.0 * (12_000.0 / 80.0) ** (cls / (N_CLASSES - 1) * 0.6)
        x = torch.zeros((m, length), device=device)
        for mult, a in ((1.0, 1.0), (2.0 ** (7.0 / 12.0) * 1.5, 0.5), (3.7, 0.25)):
            f = torch.clamp(base * mult, max=0.45 * sample_rate)[:, None]
            ph = torch.rand((m, 1), generator=gen, device=device) * (2 * np.pi)
            amp = a * (0.5 + 0.5 * torch.rand((m, 1), generator=gen, device=device))
            x += amp * torch.sin(2 * np.pi * f * t[None, :] + ph)
        noise = torch.randn((m, length), generator=gen, device=device)
        # cheap colouring: first-order difference/sum mix that depends on the class
        tilt = (0.2 + 0.75 * cls / (N_CLASSES - 1))[:, None]
        col = noise.clone()
        col[:, 1:] += tilt * noise[:, :-1]
        x += (0.1 + 0.4 * ((cls * 7) % N_CLASSES) / N_CLASSES)[:, None] * col
        x += 1e-3 * torch.randn((m, length), generator=gen, device=device)
        tail = torch.rand((m,), generator=gen, device=device) < 0.25
        cut = (torch.rand((m,), generator=gen, device=device) * 0.65 + 0.3) * length
        idx = torch.arange(length, device=device)[None, :]
        x = torch.where(tail[:, None] & (idx >= cut[:, None]), torch.zeros_like(x), x)
        peak = x.abs().amax(dim=1, keepdim=True).clamp_min(1e-30)
        out[s:e] = x / peak
    return out
