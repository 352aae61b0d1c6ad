#!/usr/bin/env python
"""Generate the committed golden vectors by running the REFERENCE itself.

Run in the dev container only (needs /root/reference; the GPU box has no
reference checkout, which is why the outputs are committed):

    python tests/golden/make_golden.py

Every array below is an output of the unmodified reference code imported from
/root/reference (src/dsp, src/features/cache.py, src/retrieval/retrieval.py) on
seeded synthetic inputs (dsp_final_b200/synth.py) or on an excerpt of a real
ESC-50 clip the reference ships under outputs/.../errors/audio/.  Nothing from
the reference's *sources* is copied; only its numerical outputs are stored.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
from pathlib import Path
from types import SimpleNamespace

sys.dont_write_bytecode = True
REF = Path(os.environ.get("DSP_REF_PATH", "/root/reference"))
REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REF))
sys.path.insert(0, str(REPO))

import numpy as np  # noqa: E402

from src.dsp.fft import fft, ifft, rfft  # noqa: E402  (reference)
from src.dsp.mfcc import MfccConfig, _dct_basis, log_mel_spectrogram, mel_filterbank, mfcc  # noqa: E402
from src.dsp.stft import _get_window, frame_signal, stft  # noqa: E402
from src.features.cache import FeatureCache  # noqa: E402
from src.retrieval.retrieval import cosine_similarity, evaluate_retrieval  # noqa: E402

from dsp_final_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402  (only to mass-produce retrieval embeddings)

OUT = Path(__file__).resolve().parent
SR = 44_100


def sha12(a: np.ndarray) -> str:
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:12]


def cfg_dict(cfg: MfccConfig) -> dict:
    return dict(sample_rate=cfg.sample_rate, frame_length=cfg.frame_length, hop_length=cfg.hop_length,
                n_fft=cfg.n_fft, n_mels=cfg.n_mels, n_mfcc=cfg.n_mfcc, f_min=cfg.f_min, f_max=cfg.f_max,
                pre_emphasis=cfg.pre_emphasis, window=cfg.window)


def small_feature_cases() -> None:
    """Short clips x (the 3x3 sweep of configs/experiments.yaml + edge settings)."""
    cases = []
    for fl in (512, 1024, 2048):
        for hop in (256, 512, 1024):
            cases.append(dict(frame_length=fl, hop_length=hop))
    cases += [
        dict(frame_length=4096, hop_length=1024),                       # extended grid
        dict(frame_length=1024, hop_length=2048),                       # hop > frame (extended grid)
        dict(frame_length=1000, hop_length=300),                        # non-pow2 frame -> zero-pad to 1024
        dict(frame_length=400, hop_length=160, n_fft=512),              # n_fft > frame
        dict(frame_length=1024, hop_length=512, n_fft=512),             # n_fft < frame -> truncation
        dict(frame_length=1024, hop_length=512, n_mels=128),            # train_cnn --n-mels 128
        dict(frame_length=512, hop_length=256, n_mels=128),             # many all-zero filters
        dict(frame_length=1024, hop_length=512, window="hamming"),
        dict(frame_length=1024, hop_length=512, window="rect", pre_emphasis=0.0),
        dict(frame_length=1024, hop_length=512, f_min=300.0, f_max=8000.0, n_mels=26, n_mfcc=20),
        dict(frame_length=256, hop_length=128, n_mels=20, n_mfcc=12),
        dict(frame_length=64, hop_length=32, n_mels=10, n_mfcc=5),
    ]
    length = 12_000
    clips = np.stack([synth.host_clip(i, 1234, length=length) for i in (0, 7, 21)])
    clips[2, 9000:] = 0.0                                               # exact-zero tail -> 1e-10 floor
    blob = {"clips": clips, "n_cases": np.int64(len(cases))}
    meta = []
    for ci, kw in enumerate(cases):
        cfg = MfccConfig(sample_rate=SR, **kw)
        meta.append(cfg_dict(cfg))
        for b in range(clips.shape[0]):
            blob[f"c{ci}_mfcc_{b}"] = mfcc(clips[b], cfg)
            blob[f"c{ci}_logmel_{b}"] = log_mel_spectrogram(clips[b], cfg)
        # plain stft (no pre-emphasis) of clip 0, float64 like the reference computes it
        nf = cfg.n_fft
        blob[f"c{ci}_stft_0"] = stft(clips[0], cfg.frame_length, cfg.hop_length, window=cfg.window, n_fft=nf)
        print("case", ci, kw, blob[f"c{ci}_mfcc_0"].shape, flush=True)
    blob["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT / "features_small.npz", **blob)


def config1_full_clip() -> None:
    """BASELINE.json configs[0]: one 5 s clip, frame 1024 / hop 512."""
    x = synth.host_clip(0, 0)
    cfg = MfccConfig(sample_rate=SR, frame_length=1024, hop_length=512)
    s = stft(x, 1024, 512)
    sel = np.array([0, 1, 2, 100, 214, 300, 427, 428])
    np.savez_compressed(
        OUT / "config1_clip.npz",
        clip=x, clip_sha12=np.array(sha12(x)),
        mfcc=mfcc(x, cfg), logmel=log_mel_spectrogram(x, cfg),
        stft_frames=sel, stft_sel=s[sel], stft_abs_fro=np.float64(np.linalg.norm(s)),
        stft_shape=np.array(s.shape),
    )
    print("config1", s.shape, flush=True)


def real_clip_excerpt() -> None:
    """Excerpt of the real ESC-50 clip used by the reference's librosa check."""
    from scipy.io import wavfile

    wav = REF / "outputs/results/run_20251221_003516/errors/audio/cnn_fold5/5-103415-A-2__gt_pig__pred_cow.wav"
    sr, pcm = wavfile.read(wav)
    assert sr == SR and pcm.dtype == np.int16 and pcm.ndim == 1
    full = (pcm.astype(np.float32) / np.float32(32768.0))               # soundfile float32 convention
    # known answer: reference stft vs float64 np.fft rounded to complex64 (librosa's dtype)
    s_ref = stft(full, 1024, 512)
    frames = frame_signal(full, 1024, 512) * _get_window("hann", 1024)
    s_np = np.fft.rfft(frames, n=1024, axis=1).astype(np.complex64)
    rel = float(np.linalg.norm(s_ref - s_np) / (np.linalg.norm(s_np) + 1e-8))
    start = 40_000
    ex = pcm[start:start + 49_152].copy()
    x = ex.astype(np.float32) / np.float32(32768.0)
    peak = np.max(np.abs(x))
    xn = x / peak                                                       # normalize_audio (float32 divide)
    cfg = MfccConfig(sample_rate=SR, frame_length=1024, hop_length=512)
    np.savez_compressed(
        OUT / "real_clip.npz", pcm16=ex, start=np.int64(start),
        mfcc=mfcc(xn, cfg), logmel=log_mel_spectrogram(xn, cfg), stft=stft(x, 1024, 512),
        librosa_style_stft_rel=np.float64(rel),
    )
    print("real clip", rel, flush=True)


def fft_cases() -> None:
    rng = np.random.default_rng(42)
    blob = {}
    specs = [(1, None), (2, None), (3, 3), (8, None), (100, None), (100, 64), (100, 300), (1024, None),
             (5000, None), (70000, None)]
    for i, (ln, n) in enumerate(specs):
        z = rng.standard_normal(ln) + 1j * rng.standard_normal(ln)
        blob[f"in_{i}"] = z
        blob[f"n_{i}"] = np.int64(-1 if n is None else n)
        outs = {"fft": fft(z, n=n), "ifft": ifft(z, n=n), "rfft": rfft(z.real, n=n)}
        for name, y in outs.items():
            # long transforms: keep 512 seeded sample positions plus the norm
            sel = np.arange(y.shape[0]) if y.shape[0] <= 2048 else np.sort(
                np.random.default_rng(i).choice(y.shape[0], 512, replace=False))
            blob[f"{name}_{i}"] = y[sel]
            blob[f"{name}_sel_{i}"] = sel
            blob[f"{name}_len_{i}"] = np.int64(y.shape[0])
            blob[f"{name}_norm_{i}"] = np.float64(np.linalg.norm(y))
    blob["n_cases"] = np.int64(len(specs))
    np.savez_compressed(OUT / "fft_cases.npz", **blob)
    print("fft cases", len(specs), flush=True)


def retrieval_case() -> None:
    """ESC-50-shaped retrieval: 1600 DB (folds 1-4) x 400 queries (fold 5), k = 10, 20."""
    n, length = 2000, 22_050                                            # 0.5 s clips keep this quick
    cfg = O.OracleConfig(SR, 1024, 512)
    clips = synth.host_clips(n, 1234, length=length)
    emb = O.features_batch(clips, cfg, want=("embed",))["embed"]        # float32 [2000, 26] (cached-path dtype)
    folds = np.array([synth.fold_of(i, n) for i in range(n)])
    targets = synth.labels(n)
    db_sel, q_sel = np.where(folds <= 4)[0], np.where(folds == 5)[0]
    mk = lambda i: SimpleNamespace(target=int(targets[i]))              # noqa: E731  (evaluate_retrieval reads .target)
    db_items, q_items = [mk(i) for i in db_sel], [mk(i) for i in q_sel]
    out = {"emb": emb, "targets": targets, "folds": folds}
    for tag, dt in (("f64", np.float64), ("f32", np.float32)):
        db, q = emb[db_sel].astype(dt), emb[q_sel].astype(dt)
        sims = cosine_similarity(q, db)
        out[f"top20_{tag}"] = np.argsort(-sims, axis=1, kind="stable")[:, :20].astype(np.int32)
        res = evaluate_retrieval(db_items, q_items, db, q, (10, 20))
        out[f"prec_{tag}"] = np.array([r.precision for r in res])
        srt = -np.sort(-sims, axis=1)[:, :21]
        out[f"min_gap_{tag}"] = np.float64(np.min(srt[:, :-1] - srt[:, 1:]))
    # a tie case: duplicated database rows must come back lowest-index-first
    db = emb[db_sel].astype(np.float64).copy()
    db[900] = db[10]
    db[1500] = db[10]
    sims = cosine_similarity(emb[q_sel].astype(np.float64), db)
    out["tie_db_rows"] = np.array([10, 900, 1500])
    out["top20_ties_f64"] = np.argsort(-sims, axis=1, kind="stable")[:, :20].astype(np.int32)
    np.savez_compressed(OUT / "retrieval.npz", **out)
    print("retrieval", out["prec_f64"], out["prec_f32"], out["min_gap_f64"], flush=True)


def known_answers() -> None:
    ka = {"cache_digests": {}, "table_sha12": {}}
    fc = FeatureCache("unused")
    for ft in ("mfcc", "log_mel"):
        for fl in (512, 1024, 2048, 4096):
            for hop in (256, 512, 1024, 2048):
                cfg = MfccConfig(sample_rate=SR, frame_length=fl, hop_length=hop)
                ka["cache_digests"][f"{ft}/{fl}/{hop}"] = fc.params_hash(ft, cfg)[0]
    cfg128 = MfccConfig(sample_rate=SR, frame_length=1024, hop_length=512, n_mels=128)
    ka["cache_digests"]["log_mel/1024/512/n_mels128"] = fc.params_hash("log_mel", cfg128)[0]
    _, params = fc.params_hash("mfcc", MfccConfig(sample_rate=SR, frame_length=1024, hop_length=512))
    ka["cache_params_mfcc_1024_512"] = params
    # digests printed by the reference's own precompute logs
    ka["published_digests"] = {"mfcc/1024/512": "e637fe1e8db0", "log_mel/1024/512": "e6e97bbcdc42",
                               "mfcc/512/256": "f92d8f09a668", "mfcc/2048/1024": "5534e76d78ec"}
    for n in (512, 1024, 2048):
        for m in (40, 128):
            ka["table_sha12"][f"fbank/{m}/{n}"] = sha12(mel_filterbank(m, n, SR))
    ka["table_sha12"]["dct/13/40"] = sha12(_dct_basis(13, 40))
    ka["table_sha12"]["hann/1024"] = sha12(_get_window("hann", 1024))
    ka["table_sha12"]["hamming/400"] = sha12(_get_window("hamming", 400))
    ka["librosa_compare_published_stft_complex_rel_error"] = 2.5356998233773264e-08
    ka["mel_edges_1024_40"] = [0, 1, 3, 4, 6, 8, 10, 13, 15, 18, 21, 25, 28, 32, 37, 41, 47, 52, 58, 65, 72, 80, 89,
                               98, 108, 119, 131, 144, 159, 174, 191, 209, 229, 251, 275, 301, 329, 360, 393, 429,
                               469, 512]
    (OUT / "known_answers.json").write_text(json.dumps(ka, indent=1, sort_keys=True))
    print("known answers written", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["known", "fft", "small", "config1", "real", "retrieval"]
    if "known" in which:
        known_answers()
    if "fft" in which:
        fft_cases()
    if "small" in which:
        small_feature_cases()
    if "config1" in which:
        config1_full_clip()
    if "real" in which:
        real_clip_excerpt()
    if "retrieval" in which:
        retrieval_case()
