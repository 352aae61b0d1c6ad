#!/usr/bin/env python
"""Golden outputs of the reference's whole retrieval pipeline on a tiny ESC-50-shaped directory.

Dev container only (needs /root/reference):

    python tests/golden/make_pipeline_golden.py        ->  tests/golden/pipeline.npz

It builds <tmp>/meta/esc50.csv + <tmp>/audio/*.wav (40 one-second 16-bit PCM clips, 5 folds x 8 clips,
5 classes; samples come from dsp_final_b200/synth.py, so the GPU test can rebuild the same files) and
runs the UNMODIFIED reference on it:

    src.datasets.esc50.Esc50Meta                 the index
    src.features.cache.FeatureCache.get_feature  -> .npy files (mfcc, frame 1024 / hop 512)
    src.retrieval.retrieval.compute_embeddings   with the cache (float32 path) and without (float64 path)
    src.tasks.retrieval.run_mfcc_retrieval       Top-3 / Top-5 hit rates, both paths

The reference decodes audio with `soundfile`, which this image lacks, and src.tasks imports
`panns_inference` through retrieval_ml.py; both are stubbed *outside* the reference tree (the attribute
src.utils.audio.sf, which the reference itself sets to None when the import fails, is pointed at the stub):
soundfile.read is backed by scipy.io.wavfile (int16 / 32768, what libsndfile returns for
dtype="float32"), panns_inference is an empty module (the MFCC task never touches it).
Only numerical outputs of the reference are stored, none of its source.
"""
from __future__ import annotations

import hashlib
import os
import sys
import tempfile
import types
from pathlib import Path

sys.dont_write_bytecode = True
REF = Path(os.environ.get("DSP_REF_PATH", "/root/reference"))
REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REF))
sys.path.insert(0, str(REPO))

import numpy as np  # noqa: E402
from scipy.io import wavfile  # noqa: E402

from dsp_final_b200 import synth  # noqa: E402

N_ITEMS, CLIP_LEN, SR, SEED = 40, 44_100, 44_100, 31
K_LIST = (3, 5)


def item_table():
    """(filename, fold, target, synth index) of the 40 items; shared with tests/test_pipeline_gpu.py."""
    rows = []
    for n in range(N_ITEMS):
        fold, target = 1 + n // 8, n % 5
        rows.append((f"{fold}-{100000 + n}-A-{target}.wav", fold, target, target + 50 * n))
    return rows


def pcm_of(synth_index: int) -> np.ndarray:
    x = synth.host_clip(synth_index, SEED, length=CLIP_LEN)
    return np.round(x.astype(np.float64) * 32767.0).astype(np.int16)


def build_dataset(root: Path) -> np.ndarray:
    (root / "meta").mkdir(parents=True)
    (root / "audio").mkdir()
    pcm = np.stack([pcm_of(r[3]) for r in item_table()])
    with (root / "meta" / "esc50.csv").open("w", encoding="utf-8") as f:
        f.write("filename,fold,target,category,esc10,src_file,take\n")
        for (name, fold, target, _), x in zip(item_table(), pcm):
            f.write(f"{name},{fold},{target},class{target},False,{100000},A\n")
            wavfile.write(root / "audio" / name, SR, x)
    return pcm


def install_stubs():
    sf = types.ModuleType("soundfile")

    def read(path, dtype="float32"):
        sr, x = wavfile.read(str(path))
        assert x.dtype == np.int16 and dtype == "float32"
        return (x.astype(np.float32) / np.float32(32768.0)), sr

    sf.read = read
    import src.utils.audio as ref_audio          # the reference module holds `sf = None` when soundfile is missing

    ref_audio.sf = sf
    panns = types.ModuleType("panns_inference")
    panns.AudioTagging = object
    sys.modules["panns_inference"] = panns


def main():
    install_stubs()
    from src.datasets.esc50 import Esc50Meta
    from src.dsp.mfcc import MfccConfig
    from src.features.cache import FeatureCache
    from src.retrieval.retrieval import compute_embeddings, cosine_similarity
    from src.tasks.retrieval import run_mfcc_retrieval

    cfg = MfccConfig(sample_rate=SR, frame_length=1024, hop_length=512)
    with tempfile.TemporaryDirectory() as tmp:
        root = Path(tmp)
        pcm = build_dataset(root / "ESC-50")
        meta = Esc50Meta(root / "ESC-50")
        assert [(i.filename, i.fold, i.target) for i in meta.items] == [r[:3] for r in item_table()]
        cache = FeatureCache(root / "features")
        res_cached = run_mfcc_retrieval(meta, cfg, k_list=K_LIST, feature_cache=cache)      # fills the cache
        res_raw = run_mfcc_retrieval(meta, cfg, k_list=K_LIST, feature_cache=None)
        feats = np.stack([cache.load_feature(it, "mfcc", cfg) for it in meta.items])
        first = cache.feature_path(meta.items[0], "mfcc", cfg)
        blob = first.read_bytes()
        header_len = len(blob) - feats[0].nbytes
        emb_cached = compute_embeddings(meta.items, cfg, feature_cache=cache)
        emb_raw = compute_embeddings(meta.items, cfg, feature_cache=None)
        db = [i for i, it in enumerate(meta.items) if it.fold != 5]
        qs = [i for i, it in enumerate(meta.items) if it.fold == 5]
        top = {}
        gaps = {}
        for name, emb in (("cached", emb_cached), ("raw", emb_raw)):
            sims = cosine_similarity(emb[qs].astype(np.float64), emb[db].astype(np.float64))
            order = np.argsort(-sims, axis=1, kind="stable")
            top[name] = order[:, :max(K_LIST)].astype(np.int32)
            s = np.take_along_axis(sims, order, axis=1)[:, :max(K_LIST) + 1]
            gaps[name] = float(np.min(s[:, :-1] - s[:, 1:]))
        rel = str(first.relative_to(root / "features"))
    np.savez_compressed(
        Path(__file__).resolve().parent / "pipeline.npz",
        pcm_sha1=np.frombuffer(hashlib.sha1(pcm.tobytes()).hexdigest().encode(), dtype=np.uint8),
        mfcc_f32=feats.astype(np.float32),
        npy_header=np.frombuffer(blob[:header_len], dtype=np.uint8),
        first_path=np.frombuffer(rel.encode(), dtype=np.uint8),
        emb_cached=emb_cached, emb_raw=emb_raw,
        top_cached=top["cached"], top_raw=top["raw"],
        min_gap=np.array([gaps["cached"], gaps["raw"]]),
        k_list=np.array(K_LIST),
        prec_cached=np.array([r.precision for r in res_cached]),
        prec_raw=np.array([r.precision for r in res_raw]),
    )
    print("emb dtypes", emb_cached.dtype, emb_raw.dtype, "min gaps", gaps, "precisions", [r.precision for r in res_cached],
          [r.precision for r in res_raw], "path", rel)


if __name__ == "__main__":
    main()
