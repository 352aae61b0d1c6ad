"""Rows a16 / a19: the reference-named pipeline (Esc50Meta -> compute_embeddings -> evaluate_retrieval,
run_mfcc_retrieval, FeatureCache-compatible get_feature) against golden outputs of the reference itself.

tests/golden/pipeline.npz was produced by tests/golden/make_pipeline_golden.py, which runs the unmodified
reference (src/datasets, src/features/cache.py, src/retrieval/retrieval.py, src/tasks/retrieval.py) on a
40-clip ESC-50-shaped directory of 16-bit PCM WAV files.  The GPU test rebuilds the same directory
(the samples come from the seeded generator; their SHA-1 is in the fixture) and runs this package's
functions of the same names on it.
"""
from __future__ import annotations

import hashlib
import importlib.util
import struct
import sys

import numpy as np
import pytest

from conftest import GOLDEN, REPO, rel_err

_spec = importlib.util.spec_from_file_location("make_pipeline_golden_tables", GOLDEN / "make_pipeline_golden.py")


def _tables():
    """item_table / pcm_of / constants of the generator script, without importing the reference."""
    src = (GOLDEN / "make_pipeline_golden.py").read_text()
    head = src[: src.index("def build_dataset")]
    head = head.replace('sys.path.insert(0, str(REF))', '')
    ns: dict = {"__file__": str(GOLDEN / "make_pipeline_golden.py"), "__name__": "pipeline_tables"}
    exec(compile(head, "make_pipeline_golden.py", "exec"), ns)
    return ns


def _write_wav16(path, rate, x):
    """Minimal canonical RIFF writer (the GPU box needs no scipy for this)."""
    data = np.ascontiguousarray(x, dtype="<i2").tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, rate, rate * 2, 2, 16)
    path.write_bytes(hdr + b"data" + struct.pack("<I", len(data)) + data)


def _build(root, ns):
    (root / "meta").mkdir(parents=True)
    (root / "audio").mkdir()
    rows = ns["item_table"]()
    pcm = np.stack([ns["pcm_of"](r[3]) for r in rows])
    lines = ["filename,fold,target,category,esc10,src_file,take"]
    for (name, fold, target, _), x in zip(rows, pcm):
        lines.append(f"{name},{fold},{target},class{target},False,100000,A")
        _write_wav16(root / "audio" / name, ns["SR"], x)
    (root / "meta" / "esc50.csv").write_text("\n".join(lines) + "\n", encoding="utf-8")
    return pcm


@pytest.fixture(scope="module")
def golden_pipeline():
    return np.load(GOLDEN / "pipeline.npz")


def test_fixture_inputs_are_reproducible(golden_pipeline, tmp_path):
    """The seeded generator rebuilds, bit for bit, the samples the reference saw; our WAV reader and index
    parse what we write (host logic, no GPU)."""
    from dsp_final_b200.audio import load_audio, load_pcm16, normalize_audio, read_wav
    from dsp_final_b200.datasets import Esc50Meta, get_fold_splits

    ns = _tables()
    pcm = _build(tmp_path / "ESC-50", ns)
    assert hashlib.sha1(pcm.tobytes()).hexdigest() == bytes(golden_pipeline["pcm_sha1"]).decode()
    meta = Esc50Meta(tmp_path / "ESC-50")
    assert [(i.filename, i.fold, i.target) for i in meta.items] == [r[:3] for r in ns["item_table"]()]
    db, qs = get_fold_splits(meta)
    assert len(db) == 32 and len(qs) == 8 and all(i.fold == 5 for i in qs)
    it = meta.items[7]
    x, sr = read_wav(it.path)
    assert sr == ns["SR"] and x.dtype == np.float32 and np.array_equal(x, pcm[7].astype(np.float32) / np.float32(32768.0))
    a, sr2 = load_audio(it.path, target_sr=ns["SR"])
    assert sr2 == ns["SR"] and np.array_equal(a, x)
    assert np.array_equal(load_pcm16(it.path, ns["SR"]), pcm[7]) and load_pcm16(it.path, 22050) is None
    n = normalize_audio(a)
    assert np.max(np.abs(n)) == 1.0 and normalize_audio(np.zeros(4, np.float32)).tolist() == [0, 0, 0, 0]
    # resampling branch follows scipy.signal.resample_poly like the reference (audio.py:27-30)
    from scipy.signal import resample_poly

    half, sr3 = load_audio(it.path, target_sr=22050)
    assert sr3 == 22050 and np.array_equal(half, resample_poly(x, 22050, ns["SR"]).astype(np.float32))


def test_wav_reader_variants(tmp_path):
    from dsp_final_b200.audio import load_audio, read_wav

    rate = 8000
    x = (np.arange(-50, 50) * 300).astype(np.int16)
    stereo = np.stack([x, -x], axis=1)
    data = stereo.astype("<i2").tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 2, rate, rate * 4, 4, 16)
    p = tmp_path / "st.wav"
    p.write_bytes(hdr + b"data" + struct.pack("<I", len(data)) + data)
    y, sr = read_wav(p)
    assert y.shape == (100, 2) and sr == rate
    mono, _ = load_audio(p)
    assert mono.shape == (100,) and np.allclose(mono, 0.0)                     # channel mean (audio.py:24-25)
    f = np.linspace(-1, 1, 64, dtype=np.float32)
    fd = f.astype("<f4").tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(fd)) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 3, 1, rate, rate * 4, 4, 32)
    q = tmp_path / "f.wav"
    q.write_bytes(hdr + b"LIST" + struct.pack("<I", 4) + b"abcd" + b"data" + struct.pack("<I", len(fd)) + fd)
    z, _ = read_wav(q)
    assert np.array_equal(z, f)
    with pytest.raises(ValueError):
        (tmp_path / "bad.wav").write_bytes(b"nope")
        read_wav(tmp_path / "bad.wav")


@pytest.mark.gpu
def test_reference_named_pipeline_on_gpu(golden_pipeline, tmp_path):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dsp_final_b200 import retrieval as R
    from dsp_final_b200.cache import BatchFeatureCache
    from dsp_final_b200.datasets import Esc50Meta, get_fold_splits
    from dsp_final_b200.dsp.mfcc import MfccConfig
    from dsp_final_b200.tasks import run_mfcc_retrieval
    from oracle import oracle as O

    g = golden_pipeline
    ns = _tables()
    pcm = _build(tmp_path / "ESC-50", ns)
    assert hashlib.sha1(pcm.tobytes()).hexdigest() == bytes(g["pcm_sha1"]).decode()
    meta = Esc50Meta(tmp_path / "ESC-50")
    cfg = MfccConfig(sample_rate=ns["SR"], frame_length=1024, hop_length=512)
    k_list = tuple(int(k) for k in g["k_list"])

    # --- run_mfcc_retrieval with a cache: fills <root>/mfcc/<digest>/fold*/x.wav.npy like FeatureCache does ---
    cache = BatchFeatureCache(tmp_path / "features")
    res = run_mfcc_retrieval(meta, cfg, k_list=k_list, feature_cache=cache)
    assert [r.k for r in res] == list(k_list)
    assert [r.precision for r in res] == [float(p) for p in g["prec_cached"]]
    first = cache.feature_path(meta.items[0], "mfcc", cfg)
    assert str(first.relative_to(tmp_path / "features")) == bytes(g["first_path"]).decode()
    blob = first.read_bytes()
    hdr = bytes(g["npy_header"])
    assert blob[: len(hdr)] == hdr and len(blob) == len(hdr) + g["mfcc_f32"][0].nbytes       # same .npy header, same size
    ours = np.stack([cache.load_feature(it, "mfcc", cfg) for it in meta.items])
    assert ours.dtype == np.float32 and ours.shape == g["mfcc_f32"].shape
    assert rel_err(ours, g["mfcc_f32"]) < 1e-4                                                 # tolerance of the north star
    for i in range(len(meta.items)):
        assert rel_err(ours[i], g["mfcc_f32"][i]) < 1e-4
    # per-item reference-named access: a hit returns the file's content, compute_feature recomputes it
    assert np.array_equal(cache.get_feature(meta.items[3], "mfcc", cfg), ours[3])
    assert np.array_equal(cache.compute_feature(meta.items[3], "mfcc", cfg), ours[3])
    with pytest.raises(ValueError):
        cache.get_feature(meta.items[3], "chroma", cfg)

    # --- compute_embeddings, both reference paths ---
    emb_c = R.compute_embeddings(meta.items, cfg, feature_cache=cache)
    emb_r = R.compute_embeddings(meta.items, cfg)                                              # reads item.path
    assert emb_c.shape == g["emb_cached"].shape == (40, 26)
    assert rel_err(emb_c, g["emb_cached"]) < 1e-4 and rel_err(emb_r, g["emb_raw"]) < 1e-4
    assert np.allclose(emb_c, emb_r, rtol=0, atol=1e-5 * np.max(np.abs(emb_r)))
    # a reference-style cache object (only get_feature) works as well
    class OnlyGet:
        def __init__(self, inner):
            self.inner, self.calls = inner, 0

        def get_feature(self, item, ft, c):
            self.calls += 1
            return self.inner.get_feature(item, ft, c)

    og = OnlyGet(cache)
    assert np.array_equal(R.compute_embeddings(meta.items[:5], cfg, feature_cache=og), emb_c[:5]) and og.calls == 5
    # injected loader (pre-decoded audio) == files
    dec = {it.filename: (pcm[i].astype(np.float32) / np.float32(32768.0)) for i, it in enumerate(meta.items)}
    dec = {k: v / np.max(np.abs(v)) for k, v in dec.items()}
    emb_l = R.compute_embeddings(meta.items, cfg, loader=lambda it: dec[it.filename])
    # files go through the fused PCM16 loads (general kernel variant), pre-decoded float32 through the aligned variant:
    # same samples bit for bit (tests/test_gpu_parity.py::test_pcm16_ingest...), float32 round-off apart in the features
    assert rel_err(emb_l, emb_r) < 1e-6

    # --- run_mfcc_retrieval without a cache ---
    res2 = run_mfcc_retrieval(meta, cfg, k_list=k_list)
    assert [r.precision for r in res2] == [float(p) for p in g["prec_raw"]]

    # --- ranking: identical inputs -> identical indices (reference embeddings in, stable argsort order out) ---
    db_items, q_items = get_fold_splits(meta)
    db_i = [i for i, it in enumerate(meta.items) if it.fold != 5]
    q_i = [i for i, it in enumerate(meta.items) if it.fold == 5]
    for name in ("cached", "raw"):
        e = g[f"emb_{name}"]
        idx = R.cosine_topk(e[q_i], e[db_i], max(k_list))
        assert np.array_equal(idx, g[f"top_{name}"]), name
        got = R.evaluate_retrieval(db_items, q_items, e[db_i], e[q_i], k_list)
        assert [r.precision for r in got] == [float(p) for p in g[f"prec_{name}"]]
    # and on our own embeddings the kernel agrees with the oracle's stable ranking
    assert np.array_equal(R.cosine_topk(emb_c[q_i], emb_c[db_i], 5), O.cosine_topk(emb_c[q_i], emb_c[db_i], 5))
