"""ctypes front end of the CPU oracle (oracle/dsp_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; nothing under
dsp_final_b200/ may import this module (tests/test_no_oracle_in_product.py
enforces that).

Function names mirror the reference entry points they restate
(src/dsp/fft.py, src/dsp/stft.py, src/dsp/mfcc.py, src/retrieval/retrieval.py);
the file:line of each is given in dsp_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "liborc.so"
_lib = None

WINDOWS = {"hann": 0, "hamming": 1, "rect": 2}


class _Cfg(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_int),
        ("frame_length", C.c_int64),
        ("hop_length", C.c_int64),
        ("n_fft", C.c_int64),
        ("n_mels", C.c_int),
        ("n_mfcc", C.c_int),
        ("f_min", C.c_double),
        ("f_max", C.c_double),
        ("pre_emphasis", C.c_double),
        ("window", C.c_int),
    ]


@dataclass
class OracleConfig:
    """Field-for-field twin of the reference MfccConfig (src/dsp/mfcc.py:10-21)."""

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


def build(force: bool = False) -> Path:
    """Compile oracle/dsp_oracle.c -> oracle/_build/liborc.so (make)."""
    src = _HERE / "dsp_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.orc_fft.restype = C.c_int64
        _lib.orc_ifft.restype = C.c_int64
        _lib.orc_num_frames.restype = C.c_int64
        _lib.orc_next_pow_two.restype = C.c_int64
        _lib.orc_hits_at_k.restype = C.c_int64
    return _lib


def _cfg(cfg) -> _Cfg:
    return _Cfg(
        int(cfg.sample_rate), int(cfg.frame_length), int(cfg.hop_length),
        int(cfg.n_fft) if cfg.n_fft else 0, int(cfg.n_mels), int(cfg.n_mfcc),
        float(cfg.f_min), -1.0 if cfg.f_max is None else float(cfg.f_max),
        float(cfg.pre_emphasis), WINDOWS[cfg.window],
    )


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def _check(rc: int, what: str) -> None:
    if rc < 0:
        raise ValueError(f"oracle {what} failed with code {rc}")


def next_pow_two(n: int) -> int:
    return int(lib().orc_next_pow_two(C.c_int64(n)))


def num_frames(length: int, frame_length: int, hop_length: int) -> int:
    t = int(lib().orc_num_frames(C.c_int64(length), C.c_int64(frame_length), C.c_int64(hop_length)))
    _check(t, "num_frames")
    return t


def fft(x, n: int | None = None, inverse: bool = False) -> np.ndarray:
    x = np.asarray(x, dtype=np.complex128).reshape(-1)
    nn = x.shape[0] if n is None else int(n)
    p = next_pow_two(nn)
    re = np.ascontiguousarray(x.real)
    im = np.ascontiguousarray(x.imag)
    ore = np.empty(p, np.float64)
    oim = np.empty(p, np.float64)
    fn = lib().orc_ifft if inverse else lib().orc_fft
    rc = fn(_p(re, C.c_double), _p(im, C.c_double), C.c_int64(x.shape[0]), C.c_int64(nn),
            _p(ore, C.c_double), _p(oim, C.c_double))
    _check(rc, "fft")
    return ore + 1j * oim


def ifft(x, n: int | None = None) -> np.ndarray:
    return fft(x, n, inverse=True)


def rfft(x, n: int | None = None) -> np.ndarray:
    y = fft(np.asarray(x, dtype=np.float64), n)
    return y[: y.shape[0] // 2 + 1]


def get_window(window: str, frame_length: int) -> np.ndarray:
    if window not in WINDOWS:
        raise ValueError(f"Unsupported window: {window}")
    out = np.empty(max(frame_length, 0), np.float64)
    _check(lib().orc_window(WINDOWS[window], C.c_int64(frame_length), _p(out, C.c_double)), "window")
    return out


def stft(signal, frame_length: int, hop_length: int, window: str = "hann",
         n_fft: int | None = None) -> np.ndarray:
    sig = np.ascontiguousarray(np.asarray(signal, dtype=np.float64).reshape(-1))
    t = num_frames(sig.shape[0], frame_length, hop_length)
    nf = frame_length if n_fft is None else int(n_fft)
    bins = next_pow_two(nf) // 2 + 1
    ore = np.empty((t, bins), np.float64)
    oim = np.empty((t, bins), np.float64)
    rc = lib().orc_stft(_p(sig, C.c_double), C.c_int64(sig.shape[0]), C.c_int64(frame_length),
                        C.c_int64(hop_length), WINDOWS[window], C.c_int64(nf),
                        _p(ore, C.c_double), _p(oim, C.c_double))
    _check(rc, "stft")
    return ore + 1j * oim


def mel_filterbank(n_mels: int, n_fft: int, sample_rate: int, f_min: float = 0.0,
                   f_max: float | None = None) -> np.ndarray:
    out = np.empty((n_mels, n_fft // 2 + 1), np.float64)
    rc = lib().orc_mel_filterbank(n_mels, C.c_int64(n_fft), sample_rate, C.c_double(f_min),
                                  C.c_double(-1.0 if f_max is None else f_max), _p(out, C.c_double))
    _check(rc, "mel_filterbank")
    return out


def dct_basis(n_mfcc: int, n: int) -> np.ndarray:
    out = np.empty((n_mfcc, n), np.float64)
    _check(lib().orc_dct_basis(n_mfcc, n, _p(out, C.c_double)), "dct_basis")
    return out


def _signal_ptrs(signal):
    sig = np.asarray(signal).reshape(-1)
    if sig.dtype == np.float32:
        s32 = np.ascontiguousarray(sig)
        return s32, _p(s32, C.c_float), None
    s64 = np.ascontiguousarray(sig, dtype=np.float64)
    return s64, None, _p(s64, C.c_double)


def log_mel_spectrogram(signal, cfg) -> np.ndarray:
    keep, p32, p64 = _signal_ptrs(signal)
    t = num_frames(keep.shape[0], cfg.frame_length, cfg.hop_length)
    out = np.empty((t, cfg.n_mels), np.float64)
    c = _cfg(cfg)
    rc = lib().orc_log_mel(p32, p64, C.c_int64(keep.shape[0]), C.byref(c), _p(out, C.c_double))
    _check(rc, "log_mel")
    return out


def mfcc(signal, cfg) -> np.ndarray:
    keep, p32, p64 = _signal_ptrs(signal)
    t = num_frames(keep.shape[0], cfg.frame_length, cfg.hop_length)
    out = np.empty((t, cfg.n_mfcc), np.float64)
    c = _cfg(cfg)
    rc = lib().orc_mfcc(p32, p64, C.c_int64(keep.shape[0]), C.byref(c), _p(out, C.c_double), None)
    _check(rc, "mfcc")
    return out


def embedding(feats) -> np.ndarray:
    f = np.ascontiguousarray(np.asarray(feats, dtype=np.float64))
    out = np.empty(2 * f.shape[1], np.float64)
    _check(lib().orc_embedding(_p(f, C.c_double), C.c_int64(f.shape[0]), f.shape[1],
                               _p(out, C.c_double)), "embedding")
    return out


def cmvn(feats, eps: float = 1e-10) -> np.ndarray:
    """(x - mean_t) / (std_t + eps) per coefficient of one clip's [T, C] features, float64 (option, see orc_cmvn)."""
    f = np.array(feats, dtype=np.float64, order="C", copy=True)
    _check(lib().orc_cmvn(_p(f, C.c_double), C.c_int64(f.shape[0]), f.shape[1], C.c_double(eps)), "cmvn")
    return f


def features_batch(clips, cfg, want=("mfcc", "log_mel", "embed"), n_threads: int = 0):
    """float32 [B, L] -> dict of float32 arrays; all host threads by default."""
    x = np.ascontiguousarray(np.asarray(clips, dtype=np.float32))
    b, length = x.shape
    t = num_frames(length, cfg.frame_length, cfg.hop_length)
    out = {}
    if "mfcc" in want:
        out["mfcc"] = np.empty((b, t, cfg.n_mfcc), np.float32)
    if "log_mel" in want:
        out["log_mel"] = np.empty((b, t, cfg.n_mels), np.float32)
    if "embed" in want:
        out["embed"] = np.empty((b, 2 * cfg.n_mfcc), np.float32)
    c = _cfg(cfg)

    def ptr(name):
        return _p(out[name], C.c_float) if name in out else None

    rc = lib().orc_features_batch(_p(x, C.c_float), C.c_int64(b), C.c_int64(length),
                                  C.c_int64(length), C.byref(c), ptr("mfcc"), ptr("log_mel"),
                                  ptr("embed"), int(n_threads))
    _check(rc, "features_batch")
    return out


def cosine_topk(q, db, k: int, n_threads: int = 0, return_scores: bool = False):
    """Stable descending top-k of the reference cosine similarity, float64."""
    qa = np.ascontiguousarray(np.asarray(q, dtype=np.float64))
    da = np.ascontiguousarray(np.asarray(db, dtype=np.float64))
    idx = np.empty((qa.shape[0], k), np.int32)
    sc = np.empty((qa.shape[0], k), np.float64)
    rc = lib().orc_cosine_topk(_p(qa, C.c_double), C.c_int64(qa.shape[0]), _p(da, C.c_double),
                               C.c_int64(da.shape[0]), int(qa.shape[1]), int(k),
                               _p(idx, C.c_int32), _p(sc, C.c_double), int(n_threads))
    _check(rc, "cosine_topk")
    return (idx, sc) if return_scores else idx


def hits_at_k(topk_idx, k: int, targets_db, targets_q) -> int:
    idx = np.ascontiguousarray(np.asarray(topk_idx, dtype=np.int32))
    tdb = np.ascontiguousarray(np.asarray(targets_db, dtype=np.int32))
    tq = np.ascontiguousarray(np.asarray(targets_q, dtype=np.int32))
    return int(lib().orc_hits_at_k(_p(idx, C.c_int32), C.c_int64(idx.shape[0]), int(idx.shape[1]),
                                   int(k), _p(tdb, C.c_int32), _p(tq, C.c_int32)))


def evaluate_retrieval(targets_db, targets_q, db_embeddings, query_embeddings, k_list):
    """hit@k 'precision' per k, as src/retrieval/retrieval.py:52-72."""
    res = []
    for k in k_list:
        idx = cosine_topk(query_embeddings, db_embeddings, k)
        res.append((int(k), hits_at_k(idx, k, targets_db, targets_q) / len(targets_q)))
    return res


def max_threads() -> int:
    return int(lib().orc_max_threads())


def relative_error(a, b) -> float:
    """The reference's own parity metric, scripts/tools/compare_librosa.py:37-38."""
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-8))
