"""Drop-in API costs: per-clip mfcc() latency (what FeatureCache.compute_feature pays) and the
host-buffer batched path with pageable vs pinned NumPy input."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from dsp_final_b200 import synth
from dsp_final_b200.batch import features_batch
from dsp_final_b200.dsp.mfcc import MfccConfig, mfcc, log_mel_spectrogram
from dsp_final_b200.dsp.stft import stft

cfg = MfccConfig(44100, 1024, 512)
clip = synth.host_clip(0, 1)
for fn, name in ((lambda: mfcc(clip, cfg), "mfcc"), (lambda: log_mel_spectrogram(clip, cfg), "log_mel"), (lambda: stft(clip, 1024, 512), "stft")):
    for _ in range(5): fn()
    t0 = time.perf_counter()
    n = 200
    for _ in range(n): fn()
    dt = (time.perf_counter() - t0) / n
    print(json.dumps({"api": name + "(one 5 s clip, numpy in, float64 out)", "ms_per_call": dt * 1e3, "audio_s_per_s": 5.0 / dt}), flush=True)
clips = synth.host_clips(64, seed=2)
big = np.concatenate([clips] * 16)                      # 1024 clips, pageable
for label, arr in (("pageable numpy", big), ("pinned numpy", torch.as_tensor(big).pin_memory().numpy())):
    features_batch(arr, cfg, ("mfcc", "log_mel"))
    t0 = time.perf_counter()
    for _ in range(3): features_batch(arr, cfg, ("mfcc", "log_mel"))
    dt = (time.perf_counter() - t0) / 3
    print(json.dumps({"api": f"features_batch(1024 clips, {label})", "ms": dt * 1e3, "audio_s_per_s": 1024 * 5.0 / dt,
                      "GBps_in": big.nbytes / dt / 1e9}), flush=True)
