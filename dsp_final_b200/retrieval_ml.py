"""Scoring half of the reference's src/retrieval/retrieval_ml.py.

retrieval_ml.py:48-74 repeats cosine_similarity / evaluate_retrieval of retrieval.py verbatim for
model embeddings (CNN 128-d, CLAP 512-d, AST 768-d, PANNs 2048-d).  The GPU top-k kernel takes any
dimension (it streams the dimension through shared memory in chunks of 32), so the same two
functions serve here; producing the embeddings (the models themselves) is out of scope.
"""
from .retrieval import RetrievalResult, cosine_similarity, cosine_topk, evaluate_retrieval, hits_at_k  # noqa: F401

__all__ = ["RetrievalResult", "cosine_similarity", "cosine_topk", "evaluate_retrieval", "hits_at_k"]
