"""ESC-50 metadata with the reference's shapes (src/datasets/esc50.py:13-51, 137-140).

Only what the retrieval task needs: the item record, the CSV index and the fold split.  The torch
Dataset classes of the reference are replaced by dsp_final_b200.stream (batches made on the GPU).
"""
from __future__ import annotations

import csv
from dataclasses import dataclass
from pathlib import Path
from typing import Iterable, List


@dataclass
class Esc50Item:
    filename: str
    fold: int
    target: int
    category: str
    path: Path


class Esc50Meta:
    """Index of <root>/meta/esc50.csv; audio files live in <root>/audio/."""

    def __init__(self, root: str | Path):
        self.root = Path(root)
        self.meta_path = self.root / "meta" / "esc50.csv"
        self.audio_dir = self.root / "audio"
        with self.meta_path.open("r", encoding="utf-8", newline="") as handle:
            self.items: List[Esc50Item] = [
                Esc50Item(row["filename"], int(row["fold"]), int(row["target"]), row["category"],
                          self.audio_dir / row["filename"])
                for row in csv.DictReader(handle)
            ]

    def by_folds(self, folds: Iterable[int]) -> List[Esc50Item]:
        wanted = frozenset(int(f) for f in folds)
        return [item for item in self.items if item.fold in wanted]


def get_fold_splits(meta: Esc50Meta):
    """(folds 1-4, fold 5): database / queries of the retrieval task, train / test of the classifiers."""
    return meta.by_folds((1, 2, 3, 4)), meta.by_folds((5,))
