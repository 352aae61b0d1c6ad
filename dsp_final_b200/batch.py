"""Batched entry points: many clips per call, on the device or from host buffers.

The reference computes one clip per Python call (src/features/cache.py:65-74,
scripts/tools/precompute_features.py:96-107 fans clips out to processes).  One
call per 5 s of audio cannot feed a B200, so the batched forms below are the
throughput API; the per-clip functions in dsp_final_b200.dsp are thin wrappers
over them and keep the reference signatures.

  * torch CUDA tensor in  -> torch CUDA tensors out, enqueued on torch's current
    stream through the device-pointer C ABI (dspx_features / dspx_stft);
  * NumPy array (or CPU tensor) in -> NumPy arrays out through the host-buffer
    C ABI (dspx_features_host / dspx_stft_host: pinned staging + copy/compute
    overlap inside the library).
"""
from __future__ import annotations

from typing import Any, Iterable

import numpy as np

from . import _lib
from .plan import Plan, get_plan

FEATURES = ("log_mel", "mfcc", "embed")


def _is_cuda_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


def _as_host_clips(clips) -> np.ndarray:
    if type(clips).__module__.startswith("torch"):
        clips = clips.detach().cpu().numpy()
    a = np.asarray(clips)
    if a.ndim == 1:
        a = a[None, :]
    if a.ndim != 2:
        raise ValueError("clips must be [n_clips, n_samples]")
    if a.dtype != np.float32 or a.strides[1] != 4 or a.strides[0] % 4 or a.strides[0] < 4 * a.shape[1]:
        a = np.ascontiguousarray(a, dtype=np.float32)      # also normalises the 0 stride of a[None, :]
    return a


def _want(want: Iterable[str]) -> tuple[str, ...]:
    w = tuple(want)
    for name in w:
        if name not in FEATURES:
            raise ValueError(f"Unsupported feature_type: {name}")      # cache.py:73
    if not w:
        raise ValueError("no output requested")
    return w


def _is_pcm16(clips) -> bool:
    dt = getattr(clips, "dtype", None)
    return dt is not None and str(dt) in ("int16", "torch.int16")


def pcm16_to_float(pcm, normalize: bool = True):
    """int16 PCM [B, L] on the GPU -> float32 [B, L]: x/32768, then x/max|x| (src/utils/audio.py:19-38)."""
    import torch

    _lib.require_device()
    x = pcm if pcm.dim() == 2 else pcm[None, :]
    if x.stride(1) != 1:
        x = x.contiguous()
    dev = x.device.index
    with torch.cuda.device(dev):
        out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().dspx_pcm16_to_float(x.data_ptr(), x.shape[0], x.shape[1], x.stride(0),
                                                   1 if normalize else 0, out.data_ptr(), out.stride(0),
                                                   torch.cuda.current_stream(dev).cuda_stream), "dspx_pcm16_to_float")
    return out


def features_batch(clips, cfg: Any, want: Iterable[str] = ("mfcc",), device: int | None = None,
                   kernel: str = "auto", normalize: bool = True) -> dict:
    """log-mel / MFCC / clip embeddings of a batch of equal-length clips.

    Follows src/dsp/mfcc.py:86-109 (+ src/retrieval/retrieval.py:19-23 for "embed").
    Returns {"log_mel": [B,T,n_mels], "mfcc": [B,T,n_mfcc], "embed": [B,2*n_mfcc]} (float32)
    restricted to `want`.  int16 input is taken as PCM16 and converted on the GPU like
    load_audio (+ normalize_audio when `normalize`), src/utils/audio.py:19-38, cache.py:66-67.
    """
    want = _want(want)
    lib = _lib.load()
    if _is_cuda_tensor(clips):
        import torch

        if _is_pcm16(clips):
            pcm = clips if clips.dim() == 2 else clips[None, :]
            pcm = pcm if pcm.stride(1) == 1 else pcm.contiguous()
            dev = pcm.device.index if device is None else device
            plan = get_plan(cfg, dev, kernel)
            if plan.kernel == "warp8" and plan.n_fft != 4096:
                # conversion and peak normalisation inside the feature kernel's loads: no float32 copy of the clips
                b, length = pcm.shape
                t = plan.num_frames(length)
                need_mfcc = "mfcc" in want or "embed" in want
                with torch.cuda.device(dev):
                    lm = torch.empty((b, t, plan.n_mels), dtype=torch.float32, device=pcm.device) if "log_mel" in want else None
                    mf = torch.empty((b, t, plan.n_mfcc), dtype=torch.float32, device=pcm.device) if need_mfcc else None
                    em = torch.empty((b, 2 * plan.n_mfcc), dtype=torch.float32, device=pcm.device) if "embed" in want else None
                    ws_bytes = int(lib.dspx_features_pcm16_workspace(b))
                    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=pcm.device)
                    _lib.check(lib.dspx_features_pcm16(plan.handle, pcm.data_ptr(), b, length, pcm.stride(0), 1 if normalize else 0,
                                                       lm.data_ptr() if lm is not None else None,
                                                       mf.data_ptr() if mf is not None else None,
                                                       em.data_ptr() if em is not None else None, ws.data_ptr(), ws_bytes,
                                                       torch.cuda.current_stream(dev).cuda_stream), "dspx_features_pcm16")
                out = {}
                if lm is not None:
                    out["log_mel"] = lm
                if "mfcc" in want:
                    out["mfcc"] = mf
                if em is not None:
                    out["embed"] = em
                return out
            clips = pcm16_to_float(clips, normalize)
        x = clips if clips.dim() == 2 else clips[None, :]
        if x.dtype != torch.float32 or x.stride(1) != 1:
            x = x.to(torch.float32).contiguous()
        dev = x.device.index if device is None else device
        plan = get_plan(cfg, dev, kernel)
        b, length = x.shape
        t = plan.num_frames(length)
        # embeddings without MFCCs: accumulated inside the feature kernel (no [B, T, n_mfcc] tensor at all)
        fused = ("embed" in want and "mfcc" not in want and plan.kernel == "warp8" and plan.n_fft != 4096
                 and x.data_ptr() % 8 == 0 and x.stride(0) % 2 == 0 and plan.hop_length % 2 == 0)
        need_mfcc = "mfcc" in want or ("embed" in want and not fused)
        with torch.cuda.device(dev):
            out = {}
            lm = torch.empty((b, t, plan.n_mels), dtype=torch.float32, device=x.device) if "log_mel" in want else None
            mf = torch.empty((b, t, plan.n_mfcc), dtype=torch.float32, device=x.device) if need_mfcc else None
            em = torch.empty((b, 2 * plan.n_mfcc), dtype=torch.float32, device=x.device) if "embed" in want else None
            stream = torch.cuda.current_stream(dev).cuda_stream
            if fused:
                ws_bytes = int(lib.dspx_embeddings_workspace(plan.handle, b))
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
                _lib.check(lib.dspx_embeddings(plan.handle, x.data_ptr(), b, length, x.stride(0), em.data_ptr(),
                                               lm.data_ptr() if lm is not None else None, ws.data_ptr(), ws_bytes, stream),
                           "dspx_embeddings")
            else:
                _lib.check(lib.dspx_features(plan.handle, x.data_ptr(), b, length, x.stride(0),
                                             lm.data_ptr() if lm is not None else None,
                                             mf.data_ptr() if mf is not None else None,
                                             em.data_ptr() if em is not None else None, stream), "dspx_features")
        if lm is not None:
            out["log_mel"] = lm
        if "mfcc" in want:
            out["mfcc"] = mf
        if em is not None:
            out["embed"] = em
        return out

    pcm = _is_pcm16(clips)
    if pcm:
        x = np.asarray(clips.numpy() if type(clips).__module__.startswith("torch") else clips)
        x = x[None, :] if x.ndim == 1 else x
        if x.strides[1] != 2 or x.strides[0] < 2 * x.shape[1] or x.strides[0] % 2:
            x = np.ascontiguousarray(x)
    else:
        x = _as_host_clips(clips)
    plan = get_plan(cfg, device, kernel)
    b, length = x.shape
    t = plan.num_frames(length)
    lm = np.empty((b, t, plan.n_mels), np.float32) if "log_mel" in want else None
    mf = np.empty((b, t, plan.n_mfcc), np.float32) if "mfcc" in want else None
    em = np.empty((b, 2 * plan.n_mfcc), np.float32) if "embed" in want else None
    outs = (lm.ctypes.data if lm is not None else None, mf.ctypes.data if mf is not None else None,
            em.ctypes.data if em is not None else None)
    if pcm:
        _lib.check(lib.dspx_features_host_pcm16(plan.handle, x.ctypes.data, b, length, max(x.strides[0] // 2, length),
                                                1 if normalize else 0, *outs), "dspx_features_host_pcm16")
    else:
        _lib.check(lib.dspx_features_host(plan.handle, x.ctypes.data, b, length, max(x.strides[0] // 4, length),
                                          *outs), "dspx_features_host")
    out = {}
    if lm is not None:
        out["log_mel"] = lm
    if mf is not None:
        out["mfcc"] = mf
    if em is not None:
        out["embed"] = em
    return out


def mfcc_batch(clips, cfg, **kw):
    return features_batch(clips, cfg, ("mfcc",), **kw)["mfcc"]


def log_mel_batch(clips, cfg, **kw):
    return features_batch(clips, cfg, ("log_mel",), **kw)["log_mel"]


def mfcc_embed_batch(clips, cfg, **kw):
    """[B, 2*n_mfcc] clip embeddings, src/retrieval/retrieval.py:19-23 batched."""
    return features_batch(clips, cfg, ("embed",), **kw)["embed"]


def log_mel_nchw(clips, cfg, device: int | None = None, kernel: str = "auto"):
    """log-mel as the CNN input tensor [B, 1, n_mels, n_frames] (float32, torch CUDA in and out).

    What LogMelTransform + DataLoader batching hand to ResNetAudio (src/train/transforms.py:16-18,
    scripts/models/train_cnn.py:46-55), written in that layout by the kernel's store epilogue.
    """
    import torch

    if not _is_cuda_tensor(clips):
        clips = torch.as_tensor(_as_host_clips(clips)).cuda()
    if _is_pcm16(clips):
        clips = pcm16_to_float(clips)
    x = clips if clips.dim() == 2 else clips[None, :]
    if x.dtype != torch.float32 or x.stride(1) != 1:
        x = x.to(torch.float32).contiguous()
    dev = x.device.index if device is None else device
    plan = get_plan(cfg, dev, kernel)
    b, length = x.shape
    t = plan.num_frames(length)
    with torch.cuda.device(dev):
        out = torch.empty((b, 1, plan.n_mels, t), dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().dspx_log_mel_nchw(plan.handle, x.data_ptr(), b, length, x.stride(0), out.data_ptr(),
                                                 torch.cuda.current_stream(dev).cuda_stream), "dspx_log_mel_nchw")
    return out


def stft_batch(clips, frame_length: int, hop_length: int, window: str = "hann", n_fft: int | None = None,
               device: int | None = None, pre_emphasis: float = 0.0, kernel: str = "auto"):
    """[B, T, n_fft_pow2//2+1] complex64 STFT, src/dsp/stft.py:43-56 batched.

    pre_emphasis > 0 applies the MFCC path's filter first (the reference stft() has none).
    """
    lib = _lib.load()
    cfg = dict(sample_rate=1, frame_length=frame_length, hop_length=hop_length, n_fft=n_fft, n_mels=1, n_mfcc=1,
               f_min=0.0, f_max=None, pre_emphasis=float(pre_emphasis), window=window)
    pre = 1 if pre_emphasis > 0 else 0
    if _is_cuda_tensor(clips):
        import torch

        x = clips if clips.dim() == 2 else clips[None, :]
        if x.dtype != torch.float32 or x.stride(1) != 1:
            x = x.to(torch.float32).contiguous()
        dev = x.device.index if device is None else device
        plan = get_plan(cfg, dev, kernel)
        b, length = x.shape
        t = plan.num_frames(length)
        with torch.cuda.device(dev):
            out = torch.empty((b, t, plan.n_bins), dtype=torch.complex64, device=x.device)
            _lib.check(lib.dspx_stft(plan.handle, x.data_ptr(), b, length, x.stride(0), pre, out.data_ptr(),
                                     torch.cuda.current_stream(dev).cuda_stream), "dspx_stft")
        return out
    x = _as_host_clips(clips)
    plan = get_plan(cfg, device, kernel)
    b, length = x.shape
    t = plan.num_frames(length)
    out = np.empty((b, t, plan.n_bins), np.complex64)
    _lib.check(lib.dspx_stft_host(plan.handle, x.ctypes.data, b, length, max(x.strides[0] // 4, length), pre,
                                  out.ctypes.data),
               "dspx_stft_host")
    return out


def embed_stats(feats):
    """concat(mean_t, std_t) of [B, T, C] float32 features (retrieval.py:38-41). torch CUDA in/out."""
    import torch

    _lib.require_device()
    if not _is_cuda_tensor(feats):
        feats = torch.as_tensor(np.asarray(feats, dtype=np.float32)).cuda()
        to_host = True
    else:
        to_host = False
    f = feats.to(torch.float32).contiguous()
    if f.dim() == 2:
        f = f[None]
    b, t, c = f.shape
    dev = f.device.index
    with torch.cuda.device(dev):
        out = torch.empty((b, 2 * c), dtype=torch.float32, device=f.device)
        _lib.check(_lib.load().dspx_embed_stats(f.data_ptr(), b, t, c, out.data_ptr(),
                                                torch.cuda.current_stream(dev).cuda_stream), "dspx_embed_stats")
    return out.cpu().numpy() if to_host else out


def cmvn(feats, eps: float = 1e-10):
    """Cepstral mean / variance normalisation per clip: (x - mean_t) / (std_t + eps) over [B, T, C] features.

    Not part of the reference pipeline (src/dsp/mfcc.py:102-109 end at log and DCT); offered because
    BASELINE.json's north_star names a log/CMVN epilogue.  torch CUDA tensors are normalised in place
    and returned; NumPy input returns a new array.
    """
    import torch

    _lib.require_device()
    host = not _is_cuda_tensor(feats)
    f = torch.as_tensor(np.ascontiguousarray(np.asarray(feats, dtype=np.float32))).cuda() if host else feats
    if f.dtype != torch.float32 or not f.is_contiguous():
        raise ValueError("cmvn needs a contiguous float32 tensor")
    v = f if f.dim() == 3 else f[None]
    b, t, c = v.shape
    dev = f.device.index
    with torch.cuda.device(dev):
        _lib.check(_lib.load().dspx_cmvn(v.data_ptr(), b, t, c, float(eps), torch.cuda.current_stream(dev).cuda_stream),
                   "dspx_cmvn")
    return f.cpu().numpy() if host else f


def fft_batch(x, n: int | None = None, inverse: bool = False):
    """Batched complex FFT over the last axis with the reference's length rule (fft.py:27-42).

    x: [..., n_in] real or complex (NumPy or torch CUDA).  Returns complex64 [..., next_pow2(n)],
    same kind of container as the input.
    """
    import torch

    _lib.require_device()
    lib = _lib.load()
    host = not _is_cuda_tensor(x)
    if host:
        xt = torch.as_tensor(np.ascontiguousarray(np.asarray(x).astype(np.complex64))).cuda()
    else:
        xt = x.to(torch.complex64)
    xt = xt.contiguous()
    lead = xt.shape[:-1]
    n_in = xt.shape[-1]
    nn = n_in if n is None else int(n)
    if nn < 1:
        raise ValueError("fft length must be positive")
    p = int(lib.dspx_next_pow_two(nn))
    batch = int(np.prod(lead)) if lead else 1
    dev = xt.device.index
    with torch.cuda.device(dev):
        out = torch.empty((batch, p), dtype=torch.complex64, device=xt.device)
        work = torch.empty((batch, p), dtype=torch.complex64, device=xt.device)
        _lib.check(lib.dspx_fft_c2c(xt.data_ptr(), batch, n_in, nn, 1 if inverse else 0, out.data_ptr(),
                                    work.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "dspx_fft_c2c")
    out = out.reshape(*lead, p)
    return out.cpu().numpy() if host else out


def dct2_rows(x, n_mfcc: int):
    """dct_type_2 over the last axis on the GPU (src/dsp/mfcc.py:73-83). NumPy or torch CUDA."""
    import torch

    _lib.require_device()
    host = not _is_cuda_tensor(x)
    xt = torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float32))).cuda() if host else x.to(torch.float32)
    xt = xt.contiguous()
    lead, n = xt.shape[:-1], xt.shape[-1]
    rows = int(np.prod(lead)) if lead else 1
    dev = xt.device.index
    with torch.cuda.device(dev):
        out = torch.empty((rows, int(n_mfcc)), dtype=torch.float32, device=xt.device)
        _lib.check(_lib.load().dspx_dct2(xt.data_ptr(), rows, int(n), int(n_mfcc), out.data_ptr(),
                                         torch.cuda.current_stream(dev).cuda_stream), "dspx_dct2")
    out = out.reshape(*lead, int(n_mfcc))
    return out.cpu().numpy() if host else out
