"""Hybrid dense + sparse orchestration and score fusion.

Mirrors ``HybridSearch`` (reference retriever/hybrid_search.py:25-403) on top of ``FlatIPSearch`` and
``ImpactSearch`` and restates ``fuse_scores_linear`` / ``fuse_scores_rrf`` (retriever/score_fuse_utils.py:48-90, 3-46).
``fuse_scores_linear`` / ``fuse_scores_rrf`` keep the reference's dict-in / dict-out form (float64 numpy on the host);
``fuse_topk_device`` is the same arithmetic on device over the sorted (scores, ids) arrays the searchers produce
(csrc/fuse_topk.cu), bit-exact with the host form and without the 2*Q*k Python dict operations.
"""
from __future__ import annotations

import logging
from typing import Optional, Sequence

import numpy as np

from .search import FlatIPSearch, add_to_heap
from .sparse_search import ImpactSearch

logger = logging.getLogger(__name__)


def fuse_scores_rrf(results_list: Sequence[dict], k: int = 60) -> dict:
    fused: dict[str, dict[str, float]] = {}
    for system_results in results_list:
        for query_id, passages in system_results.items():
            query_id = str(query_id)
            fq = fused.setdefault(query_id, {})
            passage_ids = list(passages.keys())
            scores = np.array([float(passages[pid]) for pid in passage_ids])
            order = np.argsort(-scores)
            ranks = np.arange(1, len(passage_ids) + 1)
            for pid, r in zip(np.array(passage_ids)[order], 1 / (k + ranks)):
                pid = str(pid)
                fq[pid] = fq.get(pid, 0.0) + float(r)
    return fused


def fuse_scores_linear(results_list: Sequence[dict], weights: Sequence[float] = (0.7, 0.3), eps: float = 1e-8) -> dict:
    assert len(results_list) == len(weights)
    fused: dict[str, dict[str, float]] = {}
    for system_results, weight in zip(results_list, weights):
        for query_id, passages in system_results.items():
            query_id = str(query_id)
            fq = fused.setdefault(query_id, {})
            passage_ids = list(passages.keys())
            scores = np.array([float(passages[pid]) for pid in passage_ids])
            lo, hi = np.min(scores), np.max(scores)
            weighted = (scores - lo) / (hi - lo + eps) * weight
            for pid, s in zip(passage_ids, weighted):
                pid = str(pid)
                fq[pid] = fq.get(pid, 0.0) + float(s)
    return fused


def fuse_topk_device(scores0, ids0, scores1, ids1, method: str = "linear", weights: Sequence[float] = (0.7, 0.3),
                     eps: float = 1e-8, k_rrf: float = 60.0):
    """Device fusion of two per-query top-k lists (arrays straight from ``flatip_topk`` / ``ImpactIndex.search_device``).

    Same arithmetic as ``fuse_scores_linear`` / ``fuse_scores_rrf`` (float64, reference operation order) without the
    2*Q*k Python dict operations.  Returns (ids int64 [Q, k0+k1] with -1 padding, fused float64 [Q, k0+k1], counts int32 [Q]),
    rows sorted by (fused score desc, id asc).
    """
    import torch

    from . import _C
    from ._util import require_cuda, stream_ptr

    s0 = require_cuda(scores0, "scores0").to(torch.float32).contiguous()
    s1 = require_cuda(scores1, "scores1").to(torch.float32).contiguous()
    i0 = require_cuda(ids0, "ids0").to(torch.int64).contiguous()
    i1 = require_cuda(ids1, "ids1").to(torch.int64).contiguous()
    if s0.shape != i0.shape or s1.shape != i1.shape or s0.shape[0] != s1.shape[0] or s0.ndim != 2 or s1.ndim != 2:
        raise ValueError("expected scores/ids of shape [Q, k0] and [Q, k1]")
    if method not in ("linear", "rrf"):
        raise NotImplementedError(f"score_fuse_method {method} is not supported.")
    Q, k0 = s0.shape
    k1 = s1.shape[1]
    dev = s0.device
    out_ids = torch.empty((Q, k0 + k1), dtype=torch.int64, device=dev)
    out_scores = torch.empty((Q, k0 + k1), dtype=torch.float64, device=dev)
    counts = torch.empty(Q, dtype=torch.int32, device=dev)
    lib = _C.load()
    with torch.cuda.device(dev):
        _C.check(lib.lr_fuse_topk(s0.data_ptr(), i0.data_ptr(), k0, s1.data_ptr(), i1.data_ptr(), k1, Q,
                                  0 if method == "linear" else 1, float(weights[0]), float(weights[1]), float(eps),
                                  float(k_rrf), out_ids.data_ptr(), out_scores.data_ptr(), counts.data_ptr(),
                                  stream_ptr(dev)))
    return out_ids, out_scores, counts


class HybridSearch:
    """hybrid_search.py:25-403 with B200 searchers; `faiss_search_map` must be 'flat' (the exact path)."""

    def __init__(self, model, batch_size: int = 128, corpus_chunk_size: Optional[int] = None,
                 use_multiple_gpu: bool = False, faiss_search_map: str = "flat", sparse_search_map: str = "anserini",
                 score_fuse_method: str = "linear", fuse_weights: Sequence[float] = (0.7, 0.3),
                 return_all_results: bool = False, **kwargs):
        if faiss_search_map != "flat":
            raise NotImplementedError(f"Unsupported faiss_search_map {faiss_search_map}: only the exact flat index is built")
        if sparse_search_map != "anserini":
            raise NotImplementedError(f"Unsupported sparse_search_map {sparse_search_map}")
        self.model = model
        self.batch_size = batch_size
        self.corpus_chunk_size = batch_size * 800 if corpus_chunk_size is None else corpus_chunk_size
        self.show_progress_bar = kwargs.get("show_progress_bar", True)
        self.convert_to_tensor = kwargs.get("convert_to_tensor", True)
        self.score_fuse_method = score_fuse_method
        self.fuse_weights = list(fuse_weights)
        self.return_all_results = return_all_results
        self.dense_search = FlatIPSearch(model=model, batch_size=batch_size, corpus_chunk_size=corpus_chunk_size,
                                         use_multiple_gpu=use_multiple_gpu, **kwargs)
        self.sparse_search = ImpactSearch(model=model, batch_size=batch_size, corpus_chunk_size=corpus_chunk_size, **kwargs)

    @classmethod
    def name(cls):
        return "hybrid_search"

    def encode_queries(self, queries, batch_size: int, **kwargs):
        return self.model.encode_queries(queries=queries, batch_size=batch_size, **kwargs)

    def encode_corpus(self, corpus, batch_size: int, **kwargs):
        return self.model.encode_corpus(corpus=corpus, batch_size=batch_size, **kwargs)

    def _clear(self, dense: bool = True, sparse: bool = True):
        if sparse:
            self.sparse_search._clear()
        if dense:
            self.dense_search._clear()

    def index(self, corpus_emb: dict, corpus_ids: Sequence[str]):
        use_dense = corpus_emb.get("dense_reps") is not None
        use_sparse = corpus_emb.get("sparse_reps") is not None
        assert isinstance(corpus_emb, dict) and (use_dense or use_sparse)
        if use_dense:
            self.dense_search.index(corpus_emb["dense_reps"], corpus_ids)
        if use_sparse:
            self.sparse_search.index(corpus_emb["sparse_reps"], corpus_ids)

    def retrieve_with_emb(self, query_emb: dict, query_ids: Sequence[str], top_k: int, dense: bool = True,
                          sparse: bool = True, **kwargs) -> dict:
        use_dense = query_emb.get("dense_reps") is not None
        use_sparse = query_emb.get("sparse_reps") is not None
        use_emb = query_emb.get("emb_reps") is not None
        use_tok = query_emb.get("token_id_reps") is not None
        assert isinstance(query_emb, dict) and (use_dense or use_sparse or use_emb or use_tok)
        assert dense or sparse, "Please indicate retrieval embedding types."
        results = {}
        if dense:
            if use_dense:
                results["den"] = dense_results = self.dense_search.retrieve_with_emb(query_emb["dense_reps"], query_ids, top_k=top_k)
            if use_emb:
                results["emb"] = emb_results = self.dense_search.retrieve_with_emb(query_emb["emb_reps"], query_ids, top_k=top_k)
        if sparse:
            if use_tok:
                results["tok"] = tok_results = self.sparse_search.retrieve_with_emb(query_emb["token_id_reps"], query_ids, top_k=top_k)
            if use_sparse:
                results["spr"] = spr_results = self.sparse_search.retrieve_with_emb(query_emb["sparse_reps"], query_ids, top_k=top_k)
            if dense and use_dense and use_sparse:
                results["den_spr"] = self._fuse_results(dense_results, spr_results, weights=self.fuse_weights)
            if dense and use_emb and use_tok:
                results["emb_tok"] = self._fuse_results(emb_results, tok_results, weights=self.fuse_weights)
                results["default"] = results["emb_tok"]
        return results

    def retrieve_arrays(self, query_emb: dict, top_k: int):
        """emb + tok retrieval and fusion without leaving the device: (fused ids i64 [Q, 2k] as corpus positions, fused
        scores f64 [Q, 2k], counts i32 [Q]).  Both searchers must index the same corpus in the same order."""
        s0, i0 = self.dense_search.retrieve_arrays(query_emb["emb_reps"] if query_emb.get("emb_reps") is not None
                                                   else query_emb["dense_reps"], top_k)
        s1, i1 = self.sparse_search.retrieve_arrays(query_emb["token_id_reps"] if query_emb.get("token_id_reps") is not None
                                                    else query_emb["sparse_reps"], top_k)
        return fuse_topk_device(s0, i0, s1, i1, method=self.score_fuse_method, weights=self.fuse_weights)

    def _add_to_heap(self, sub_results, result_heaps, top_k, ignore_identical_ids):
        return add_to_heap(sub_results, result_heaps, top_k, ignore_identical_ids)

    def _fuse_results(self, dense_results=None, sparse_results=None, weights=(0.7, 0.3)):
        if dense_results is None and sparse_results is not None:
            return sparse_results
        if dense_results is not None and sparse_results is None:
            return dense_results
        if dense_results is not None and sparse_results is not None:
            if self.score_fuse_method == "rrf":
                return fuse_scores_rrf([dense_results, sparse_results])
            if self.score_fuse_method == "linear":
                return fuse_scores_linear([dense_results, sparse_results], weights=weights)
            raise NotImplementedError(f"score_fuse_method {self.score_fuse_method} is not supported.")
        raise ValueError("All scores are None. Please check model settings.")

    def search(self, corpus: dict, queries: dict, top_k: int = 1000, score_function: str = None,
               return_sorted: bool = False, ignore_identical_ids: bool = False, **kwargs):
        """Chunk loop of hybrid_search.py:234-403: dense is indexed/retrieved per chunk and heap-merged, sparse is
        indexed per chunk and retrieved once at the end, then fused."""
        if not isinstance(queries, dict) or not isinstance(corpus, dict):
            raise NotImplementedError("HybridSearch.search takes dict corpora / queries")
        query_ids = list(queries.keys())
        queries_list = [queries[qid] for qid in queries]
        qe = self.model.encode_queries(queries_list, batch_size=self.batch_size,
                                       show_progress_bar=self.show_progress_bar, convert_to_tensor=self.convert_to_tensor)
        use_dense, use_sparse = "dense_reps" in qe, "sparse_reps" in qe
        use_emb, use_tok = "emb_reps" in qe, "token_id_reps" in qe
        assert isinstance(qe, dict) and (use_dense or use_sparse or use_emb or use_tok)
        corpus_ids = sorted(corpus, key=lambda k_: len(corpus[k_].get("text", "")) if isinstance(corpus[k_], dict)
                            else len(corpus[k_]), reverse=True)
        corpus_list = [corpus[cid] for cid in corpus_ids]
        dense_heaps = {qid: [] for qid in query_ids} if use_dense else None
        emb_heaps = {qid: [] for qid in query_ids} if use_emb else None
        for start in range(0, len(corpus_list), self.corpus_chunk_size):
            end = min(start + self.corpus_chunk_size, len(corpus_list))
            sub = self.model.encode_corpus(corpus_list[start:end], batch_size=self.batch_size,
                                           show_progress_bar=self.show_progress_bar,
                                           convert_to_tensor=self.convert_to_tensor)
            assert isinstance(sub, dict)
            self.index(sub, corpus_ids[start:end])
            sub_results = self.retrieve_with_emb(qe, query_ids, top_k=top_k, dense=True, sparse=False)
            if use_dense:
                add_to_heap(sub_results["den"], dense_heaps, top_k, ignore_identical_ids)
            if use_emb:
                add_to_heap(sub_results["emb"], emb_heaps, top_k, ignore_identical_ids)
            self._clear(dense=True, sparse=False)

        def parse(heaps):
            return {qid: {pid: score for score, pid in heaps[qid]} for qid in heaps}

        dense_results = parse(dense_heaps) if use_dense else None
        emb_results = parse(emb_heaps) if use_emb else None
        tok_results = self.sparse_search.retrieve_with_emb(qe["token_id_reps"], query_ids, top_k=top_k) if use_tok else None
        spr_results = self.sparse_search.retrieve_with_emb(qe["sparse_reps"], query_ids, top_k=top_k) if use_sparse else None
        results, default = {}, None
        if use_dense:
            results["den"] = default = dense_results
        if use_sparse:
            results["spr"] = default = spr_results
        if use_emb:
            results["emb"] = default = emb_results
        if use_tok:
            results["tok"] = default = tok_results
        if use_dense and use_sparse:
            results["den_spr"] = default = self._fuse_results(dense_results, spr_results, weights=self.fuse_weights)
        if use_emb and use_tok:
            results["emb_tok"] = default = self._fuse_results(emb_results, tok_results, weights=self.fuse_weights)
        self._clear()
        return results if self.return_all_results else default
