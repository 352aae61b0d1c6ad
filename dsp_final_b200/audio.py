"""Audio file ingest with the reference's semantics (src/utils/audio.py:19-38).

    load_audio(path, target_sr=None) -> (float32 mono samples, sample_rate)
    normalize_audio(audio)           -> audio / max|audio|   (unchanged when the clip is all zero)

ESC-50 ships 16-bit PCM WAV files.  The reference decodes them with `soundfile.read(dtype="float32")`,
which yields int16 / 32768 exactly; `read_wav` below does the same with the standard library's RIFF
parsing (PCM 8/16/24/32-bit and IEEE float), so the package has no hard dependency on soundfile.  When
soundfile is importable it is used, like in the reference.  Decoding is host work (file I/O); the batched
GPU path takes the raw int16 samples instead (`load_pcm16`, dspx_features_host_pcm16: the conversion and
the peak normalisation then run on the device and PCIe carries half the bytes).
"""
from __future__ import annotations

import struct
from pathlib import Path
from typing import Tuple

import numpy as np

try:                                    # the reference's decoder, when present
    import soundfile as _sf
except ImportError:                     # pragma: no cover - depends on the image
    _sf = None


def _riff_chunks(blob: bytes):
    if len(blob) < 12 or blob[:4] != b"RIFF" or blob[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file")
    pos = 12
    while pos + 8 <= len(blob):
        tag, size = blob[pos:pos + 4], struct.unpack_from("<I", blob, pos + 4)[0]
        yield tag, blob[pos + 8: pos + 8 + size]
        pos += 8 + size + (size & 1)


def read_wav_raw(path: str | Path) -> Tuple[np.ndarray, int, int]:
    """(samples [n, channels] in the file's own dtype, sample_rate, format tag) of a WAV file."""
    fmt = data = None
    for tag, body in _riff_chunks(Path(path).read_bytes()):
        if tag == b"fmt ":
            fmt = body
        elif tag == b"data":
            data = body
    if fmt is None or data is None:
        raise ValueError(f"{path}: missing fmt or data chunk")
    code, channels, rate, _bps, _align, bits = struct.unpack_from("<HHIIHH", fmt, 0)
    if code == 0xFFFE and len(fmt) >= 26:                       # WAVE_FORMAT_EXTENSIBLE: real tag in the GUID
        code = struct.unpack_from("<H", fmt, 24)[0]
    if code == 1 and bits == 16:
        x = np.frombuffer(data, dtype="<i2")
    elif code == 1 and bits == 8:
        x = np.frombuffer(data, dtype=np.uint8)
    elif code == 1 and bits == 32:
        x = np.frombuffer(data, dtype="<i4")
    elif code == 1 and bits == 24:
        b = np.frombuffer(data[: len(data) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        x = (b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16))
        x = np.where(x & 0x800000, x - (1 << 24), x).astype(np.int32)
    elif code == 3 and bits in (32, 64):
        x = np.frombuffer(data, dtype="<f4" if bits == 32 else "<f8")
    else:
        raise ValueError(f"{path}: unsupported WAV encoding (format {code}, {bits} bits)")
    n = x.size // channels
    return x[: n * channels].reshape(n, channels), int(rate), (code << 8) | bits


def read_wav(path: str | Path) -> Tuple[np.ndarray, int]:
    """float32 samples in [-1, 1) like soundfile.read(dtype="float32"): integer PCM divided by 2^(bits-1)."""
    x, rate, tag = read_wav_raw(path)
    code, bits = tag >> 8, tag & 0xFF
    if code == 3:
        y = x.astype(np.float32)
    elif bits == 8:
        y = (x.astype(np.float32) - 128.0) / 128.0
    else:
        y = (x.astype(np.float64) / float(1 << (bits - 1))).astype(np.float32)
    return (y[:, 0] if y.shape[1] == 1 else y), rate


def load_audio(path: str | Path, target_sr: int | None = None) -> Tuple[np.ndarray, int]:
    """src/utils/audio.py:19-31: decode, average channels, resample when the rates differ."""
    path = Path(path)
    if _sf is not None:
        audio, sr = _sf.read(path, dtype="float32")
    else:
        audio, sr = read_wav(path)
    if audio.ndim > 1:
        audio = np.mean(audio, axis=1)
    if target_sr is not None and sr != target_sr:
        from scipy.signal import resample_poly                  # same resampler as the reference

        audio = resample_poly(audio, target_sr, sr)
        sr = target_sr
    return audio.astype(np.float32), sr


def normalize_audio(audio: np.ndarray) -> np.ndarray:
    """src/utils/audio.py:34-38."""
    peak = np.max(np.abs(audio)) if audio.size else 0.0
    return audio / peak if peak > 0 else audio


def load_pcm16(path: str | Path, target_sr: int | None = None) -> np.ndarray | None:
    """The file's int16 samples when it is mono 16-bit PCM at `target_sr` (the ESC-50 case), else None.

    Such clips can skip the host-side float conversion altogether: features_batch() takes int16 and
    converts / peak-normalises on the GPU, bit-identically to load_audio + normalize_audio.
    """
    try:
        x, rate, tag = read_wav_raw(path)
    except ValueError:
        return None
    if tag != ((1 << 8) | 16) or x.shape[1] != 1 or (target_sr is not None and rate != target_sr):
        return None
    return np.ascontiguousarray(x[:, 0])
