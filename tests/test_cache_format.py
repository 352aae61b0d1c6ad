"""Row f1: the batched cache writer produces the reference's cache, byte for byte."""
from __future__ import annotations

import hashlib
import importlib.util
import json
import sys
from types import SimpleNamespace

import numpy as np
import pytest

from conftest import REFERENCE


def _items(n):
    return [SimpleNamespace(filename=f"{1 + i // 400}-{100000 + i}-A-{i % 50}.wav", fold=1 + i // 400, target=i % 50)
            for i in range(n)]


def test_digests_and_paths_are_the_published_ones(tmp_path, known_answers):
    from dsp_final_b200.cache import BatchFeatureCache
    from dsp_final_b200.dsp.mfcc import MfccConfig

    bc = BatchFeatureCache(tmp_path)
    for key, want in known_answers["cache_digests"].items():
        ft, fl, hop, *rest = key.split("/")
        cfg = MfccConfig(44100, int(fl), int(hop), **({"n_mels": 128} if rest else {}))
        assert bc.params_hash(ft, cfg)[0] == want, key
    cfg = MfccConfig(44100, 1024, 512)
    assert bc.params_hash("mfcc", cfg)[1] == known_answers["cache_params_mfcc_1024_512"]
    it = _items(1)[0]
    assert bc.feature_path(it, "mfcc", cfg) == tmp_path / "mfcc" / "e637fe1e8db0" / "fold1" / f"{it.filename}.npy"
    # dict configs are hashed verbatim, no n_fft / f_max defaulting (cache.py:35-38; how retrieval_ml.py keys its
    # embedding_* caches): digests below were produced by the reference's FeatureCache.params_hash
    assert bc.params_hash("mfcc", {"sample_rate": 44100, "frame_length": 1024, "hop_length": 512, "n_fft": None,
                                   "n_mels": 40, "n_mfcc": 13, "f_min": 0.0, "f_max": None, "pre_emphasis": 0.97,
                                   "window": "hann"})[0] == "6f264003971b"
    assert bc.params_hash("embedding_panns", {"model": "cnn14", "sr": 32000})[0] == "b2b4aab840f8"


def test_save_load_manifest_roundtrip(tmp_path):
    from dsp_final_b200.cache import BatchFeatureCache
    from dsp_final_b200.dsp.mfcc import MfccConfig

    bc = BatchFeatureCache(tmp_path)
    cfg = MfccConfig(44100, 1024, 512)
    items = _items(5)
    feats = np.random.default_rng(0).standard_normal((5, 429, 13))            # float64 in -> float32 on disk
    recs = bc.save_features(items, feats, "mfcc", cfg, workers=3)
    for i, it in enumerate(items):
        got = bc.load_feature(it, "mfcc", cfg)
        assert got.dtype == np.float32 and got.flags["C_CONTIGUOUS"]
        assert np.array_equal(got, feats[i].astype(np.float32))
    m = json.loads(bc.write_manifest("mfcc", cfg, recs).read_text())
    assert list(m) == ["feature_type", "hash", "params", "created_at", "num_files", "files"]
    assert m["num_files"] == 5 and m["hash"] == "e637fe1e8db0" and m["files"][0]["shape"] == [429, 13]
    assert set(m["files"][0]) == {"filename", "fold", "path", "shape"}
    # corrupt file self-heals to a miss (cache.py:56-63)
    p = bc.feature_path(items[0], "mfcc", cfg)
    p.write_bytes(b"not an npy")
    assert bc.load_feature(items[0], "mfcc", cfg) is None and not p.exists()
    with pytest.raises(ValueError):
        bc.save_features(items, feats, "chroma", cfg)
    assert not list(tmp_path.rglob("*.tmp"))


@pytest.mark.needs_reference
def test_reference_cache_reads_our_files_and_bytes_match(tmp_path):
    """The reference's FeatureCache (imported from its checkout) and ours are interchangeable on disk."""
    from dsp_final_b200.cache import BatchFeatureCache
    from dsp_final_b200.dsp.mfcc import MfccConfig

    sys.dont_write_bytecode = True
    sys.path.insert(0, str(REFERENCE))
    try:
        saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "src" or k.startswith("src.")}
        ref_cache = importlib.import_module("src.features.cache")
        ref_cfg_cls = importlib.import_module("src.dsp.mfcc").MfccConfig
    finally:
        sys.path.remove(str(REFERENCE))
    try:
        items = _items(3)
        feats = np.random.default_rng(1).standard_normal((3, 215, 40)).astype(np.float32)
        ours, theirs = BatchFeatureCache(tmp_path / "ours"), ref_cache.FeatureCache(tmp_path / "theirs")
        cfg, rcfg = MfccConfig(44100, 1024, 1024), ref_cfg_cls(44100, 1024, 1024)
        recs = ours.save_features(items, feats, "log_mel", cfg)
        reader = ref_cache.FeatureCache(tmp_path / "ours")                    # reference code reads OUR files
        for i, it in enumerate(items):
            assert np.array_equal(reader.load_feature(it, "log_mel", rcfg), feats[i])
            theirs.save_feature(theirs.feature_path(it, "log_mel", rcfg), feats[i])
            a = ours.feature_path(it, "log_mel", cfg).read_bytes()
            b = theirs.feature_path(it, "log_mel", rcfg).read_bytes()
            assert hashlib.sha1(a).hexdigest() == hashlib.sha1(b).hexdigest()  # identical .npy bytes
            assert ours.feature_path(it, "log_mel", cfg).relative_to(tmp_path / "ours") == \
                theirs.feature_path(it, "log_mel", rcfg).relative_to(tmp_path / "theirs")
        m_ours = json.loads(ours.write_manifest("log_mel", cfg, recs).read_text())
        m_ref = json.loads(theirs.write_manifest("log_mel", rcfg, recs).read_text())
        assert list(m_ours) == list(m_ref)
        for k in ("feature_type", "hash", "params", "num_files"):
            assert m_ours[k] == m_ref[k]
    finally:
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)


@pytest.mark.gpu
def test_precompute_on_gpu_writes_reference_cache(tmp_path):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dsp_final_b200 import synth
    from dsp_final_b200.cache import BatchFeatureCache
    from dsp_final_b200.dsp.mfcc import MfccConfig
    from oracle import oracle as O

    items = _items(20)
    clips = synth.host_clips(20, seed=9, length=30_000)
    cfg = MfccConfig(44100, 1024, 512)
    bc = BatchFeatureCache(tmp_path)
    manifests = bc.precompute(items, cfg, ("mfcc", "log_mel"), clips=clips, batch=8)
    ref = O.features_batch(clips, O.OracleConfig(44100, 1024, 512), want=("mfcc", "log_mel"))
    for i, it in enumerate(items):
        for ft in ("mfcc", "log_mel"):
            got = bc.load_feature(it, ft, cfg)
            assert got.dtype == np.float32
            assert O.relative_error(got, ref[ft][i]) < 1e-4
    m = json.loads(manifests["mfcc"].read_text())
    assert m["num_files"] == 20 and m["files"][3]["shape"] == [57, 13]
    # second call is all cache hits and rewrites only the manifest
    before = {p: p.stat().st_mtime_ns for p in tmp_path.rglob("*.npy")}
    bc.precompute(items, cfg, ("mfcc", "log_mel"), clips=clips, batch=8)
    assert before == {p: p.stat().st_mtime_ns for p in tmp_path.rglob("*.npy")}
