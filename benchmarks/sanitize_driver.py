"""Tiny run of every kernel in libdspx for compute-sanitizer (memcheck / racecheck)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from dsp_final_b200 import retrieval as R
from dsp_final_b200 import synth
from dsp_final_b200.batch import embed_stats, features_batch, fft_batch, log_mel_nchw, pcm16_to_float, stft_batch
from dsp_final_b200.dsp.mfcc import MfccConfig, dct_type_2

clips = torch.as_tensor(synth.host_clips(6, seed=1, length=9001 + 1)).cuda()       # odd frame counts
for fl, hop in ((512, 256), (1024, 512), (2048, 1024), (1000, 300), (256, 128)):
    for kernel in ("auto", "generic"):
        cfg = MfccConfig(44100, fl, hop)
        out = features_batch(clips, cfg, ("log_mel", "mfcc", "embed"), kernel=kernel)
        assert torch.isfinite(out["mfcc"]).all()
log_mel_nchw(clips, MfccConfig(44100, 1024, 512, n_mels=128))
stft_batch(clips, 1024, 512)
stft_batch(clips, 400, 160, n_fft=512)
fft_batch(torch.randn(3, 1000, device="cuda"), n=1000)
fft_batch(np.random.randn(5), n=None, inverse=True)
embed_stats(torch.randn(4, 57, 13, device="cuda"))
dct_type_2(np.random.randn(7, 40), 13)
pcm16_to_float(torch.randint(-3000, 3000, (3, 5000), dtype=torch.int16, device="cuda"))
features_batch((np.random.randn(3, 5000) * 2000).astype(np.int16), MfccConfig(44100, 1024, 512), ("mfcc",))
features_batch(synth.host_clips(5, seed=2, length=7000), MfccConfig(44100, 1024, 512), ("mfcc", "log_mel"))   # host pipeline
g = torch.Generator(device="cuda"); g.manual_seed(0)
for nq, ndb, dim, k in ((70, 900, 26, 20), (5, 3000, 26, 10), (33, 500, 80, 7), (300, 70000, 26, 20)):
    q = torch.randn((nq, dim), generator=g, device="cuda"); db = torch.randn((ndb, dim), generator=g, device="cuda")
    idx = R.cosine_topk(q, db, k)
    R.hits_at_k(idx, k, torch.randint(0, 50, (ndb,), device="cuda").cpu().numpy(), torch.randint(0, 50, (nq,), device="cuda").cpu().numpy())
R.cosine_similarity(torch.randn(9, 26, device="cuda"), torch.randn(40, 26, device="cuda"))
torch.cuda.synchronize()
print("sanitize driver done")
