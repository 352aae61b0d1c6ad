"""The retrieval task glue (src/tasks/retrieval.py:11-21)."""
from __future__ import annotations

from typing import Iterable

from .datasets import get_fold_splits
from .retrieval import compute_embeddings, evaluate_retrieval


def run_mfcc_retrieval(meta, cfg, k_list: Iterable[int] = (10, 20), feature_cache=None):
    """Fold-5 clips query the folds 1-4 database with MFCC mean/std embeddings; hit@k per k.

    Same signature and result list as the reference.  Embeddings come from one fused GPU batch per
    split (or from the feature cache, filled in batches), ranking from dspx_cosine_topk.
    """
    db_items, query_items = get_fold_splits(meta)
    db = compute_embeddings(db_items, cfg, feature_cache=feature_cache)
    queries = compute_embeddings(query_items, cfg, feature_cache=feature_cache)
    return evaluate_retrieval(db_items, query_items, db, queries, k_list=k_list)
