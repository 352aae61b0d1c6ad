"""Batched writer/reader of the reference's feature-cache format.

The reference fills its cache one clip per call (FeatureCache.get_feature, src/features/cache.py:76-87;
scripts/tools/precompute_features.py:96-107 fans that out to processes: 19 minutes for ESC-50).  This
module computes whole batches on the GPU and writes the very same files, so the reference's
run_retrieval.py / train_cnn.py read GPU-made features unchanged:

    <root>/<feature_type>/<sha1(params)[:12]>/fold<k>/<filename>.npy      float32 [n_frames, n_coef], C order
    <root>/<feature_type>/<digest>/manifest.json                          same keys as cache.py:96-112

Format invariants kept bit for bit (cache.py:24-49, 74, 89-112): the digest recipe
(json.dumps(params, sort_keys=True, ensure_ascii=True) with n_fft and f_max defaulted), the filename
keeping its .wav suffix, np.save of a float32 array, atomic tmp + os.replace, manifest keys.
"""
from __future__ import annotations

import hashlib
import json
import os
import uuid
from concurrent.futures import ThreadPoolExecutor
from dataclasses import asdict, is_dataclass
from datetime import datetime, timezone
from pathlib import Path
from typing import Any, Callable, Iterable, Sequence

import numpy as np

FEATURE_TYPES = ("mfcc", "log_mel")


class BatchFeatureCache:
    def __init__(self, root: str | Path = "outputs/features", enabled: bool = True):
        self.root = Path(root)
        self.enabled = enabled

    # ---- naming: identical to FeatureCache._cfg_dict / params_hash / feature_path --------------------
    @staticmethod
    def _cfg_dict(cfg: Any) -> dict:
        params = asdict(cfg) if is_dataclass(cfg) else dict(cfg)
        params["n_fft"] = (cfg.n_fft if not isinstance(cfg, dict) else cfg.get("n_fft")) or params["frame_length"]
        if params.get("f_max") is None:
            params["f_max"] = params["sample_rate"] / 2
        return params

    def params_hash(self, feature_type: str, cfg: Any) -> tuple[str, dict]:
        params = {"feature_type": feature_type, **self._cfg_dict(cfg)}
        payload = json.dumps(params, sort_keys=True, ensure_ascii=True)
        return hashlib.sha1(payload.encode("utf-8")).hexdigest()[:12], params

    def feature_dir(self, feature_type: str, cfg: Any) -> Path:
        return self.root / feature_type / self.params_hash(feature_type, cfg)[0]

    def feature_path(self, item, feature_type: str, cfg: Any) -> Path:
        return self.feature_dir(feature_type, cfg) / f"fold{item.fold}" / f"{item.filename}.npy"

    # ---- I/O ---------------------------------------------------------------------------------------
    @staticmethod
    def _save_one(path: Path, feat: np.ndarray) -> None:
        path.parent.mkdir(parents=True, exist_ok=True)
        tmp = path.with_suffix(path.suffix + f".{uuid.uuid4().hex}.tmp")
        with tmp.open("wb") as f:
            np.save(f, feat)
        os.replace(tmp, path)

    def load_feature(self, item, feature_type: str, cfg: Any) -> np.ndarray | None:
        """Cache hit or None; a corrupt file is removed like cache.py:56-63 does."""
        if not self.enabled:
            return None
        path = self.feature_path(item, feature_type, cfg)
        if path.exists():
            try:
                return np.load(path)
            except (OSError, ValueError, EOFError):
                try:
                    path.unlink()
                except OSError:
                    pass
        return None

    def save_features(self, items: Sequence[Any], feats: np.ndarray, feature_type: str, cfg: Any,
                      workers: int = 8) -> list[dict]:
        """Write feats[i] ([n_frames, n_coef]) for items[i]; returns the manifest records."""
        if feature_type not in FEATURE_TYPES:
            raise ValueError(f"Unsupported feature_type: {feature_type}")
        feats = np.asarray(feats)
        if feats.shape[0] != len(items):
            raise ValueError("one feature array per item expected")
        paths = [self.feature_path(it, feature_type, cfg) for it in items]
        arrays = [np.ascontiguousarray(feats[i], dtype=np.float32) for i in range(len(items))]
        if workers > 1 and len(items) > 1:
            with ThreadPoolExecutor(max_workers=workers) as pool:
                list(pool.map(self._save_one, paths, arrays))
        else:
            for p, a in zip(paths, arrays):
                self._save_one(p, a)
        return [{"filename": it.filename, "fold": it.fold, "path": str(p), "shape": list(a.shape)}
                for it, p, a in zip(items, paths, arrays)]

    def write_manifest(self, feature_type: str, cfg: Any, records: Iterable[dict]) -> Path:
        digest, params = self.params_hash(feature_type, cfg)
        records = list(records)
        manifest = {
            "feature_type": feature_type,
            "hash": digest,
            "params": params,
            "created_at": datetime.now(timezone.utc).replace(tzinfo=None).isoformat() + "Z",
            "num_files": len(records),
            "files": records,
        }
        out_dir = self.feature_dir(feature_type, cfg)
        out_dir.mkdir(parents=True, exist_ok=True)
        path = out_dir / "manifest.json"
        with path.open("w", encoding="utf-8") as f:
            json.dump(manifest, f, indent=2, ensure_ascii=True)
        return path

    # ---- batched precompute (GPU) -----------------------------------------------------------------------
    def precompute(self, items: Sequence[Any], cfg: Any, feature_types: Iterable[str] = FEATURE_TYPES,
                   clips: np.ndarray | None = None, loader: Callable[[Any], np.ndarray] | None = None,
                   batch: int = 256, workers: int = 8, normalize: bool = True, skip_existing: bool = True) -> dict:
        """Fill the cache for `items`: the batched twin of precompute_features.py:62-113.

        Audio comes from `clips` ([N, L] float32 already loaded/normalised, or int16 PCM which is
        converted and peak-normalised on the GPU) or from `loader(item)` (decoding files is out of
        scope here).  Returns {feature_type: manifest path}.
        """
        from .batch import features_batch

        feature_types = tuple(feature_types)
        for ft in feature_types:
            if ft not in FEATURE_TYPES:
                raise ValueError(f"Unsupported feature_type: {ft}")
        items = list(items)
        if clips is None and loader is None:
            raise ValueError("precompute needs clips or a loader")
        records = {ft: [] for ft in feature_types}
        for s in range(0, len(items), batch):
            chunk = items[s:s + batch]
            todo = [i for i, it in enumerate(chunk)
                    if not (skip_existing and all(self.feature_path(it, ft, cfg).exists() for ft in feature_types))]
            if todo:
                if clips is not None:
                    x = np.asarray(clips[s:s + batch])[todo]
                else:
                    x = np.stack([np.asarray(loader(chunk[i])) for i in todo])
                out = features_batch(x, cfg, feature_types, normalize=normalize)
                for ft in feature_types:
                    self.save_features([chunk[i] for i in todo], out[ft], ft, cfg, workers=workers)
            for ft in feature_types:
                for it in chunk:
                    p = self.feature_path(it, ft, cfg)
                    with p.open("rb") as f:
                        ver = np.lib.format.read_magic(f)
                        shape = (np.lib.format.read_array_header_1_0(f) if ver == (1, 0)
                                 else np.lib.format.read_array_header_2_0(f))[0]
                    records[ft].append({"filename": it.filename, "fold": it.fold, "path": str(p), "shape": list(shape)})
        return {ft: self.write_manifest(ft, cfg, records[ft]) for ft in feature_types}
