"""Hybrid dense + sparse orchestration and score fusion.

Seam: ``HybridSearch`` (reference retriever/hybrid_search.py:25-403) and ``fuse_scores_linear`` / ``fuse_scores_rrf``
(retriever/score_fuse_utils.py:48-90, 3-46).  The design is array-first: both searchers return sorted ``(scores, ids)``
device arrays whose ids are corpus positions, chunked dense retrieval is folded into a running device top-k with
``lr_topk_merge`` (no Python heap), the two systems are fused by ``lr_fuse_topk`` (csrc/fuse_topk.cu: float64 in the
reference's operation order, bit-exact with the numpy code), and the reference's nested result dicts are built once at
the end.  The dict-in / dict-out fusion functions are adapters that pack dicts into arrays for the same kernel.
"""
from __future__ import annotations

import logging
from typing import Optional, Sequence

import numpy as np
import torch

from . import _C
from ._util import require_cuda, stream_ptr
from .search import FlatIPSearch, identical_positions, rows_to_dict, sorted_corpus
from .sparse_search import ImpactSearch

logger = logging.getLogger(__name__)


def fuse_topk_device(scores0, ids0, scores1, ids1, method: str = "linear", weights: Sequence[float] = (0.7, 0.3),
                     eps: float = 1e-8, k_rrf: float = 60.0):
    """Device fusion of two per-query top-k lists (arrays straight from ``flatip_topk`` / ``ImpactIndex.search_device``).

    Same arithmetic as ``fuse_scores_linear`` / ``fuse_scores_rrf`` (float64, reference operation order) without the
    2*Q*k Python dict operations.  Returns (ids int64 [Q, k0+k1] with -1 padding, fused float64 [Q, k0+k1], counts int32 [Q]),
    rows sorted by (fused score desc, id asc).
    """
    f64 = scores0.dtype == torch.float64 or scores1.dtype == torch.float64  # dict-shaped callers (Python floats)
    sdt = torch.float64 if f64 else torch.float32
    s0 = require_cuda(scores0, "scores0").to(sdt).contiguous()
    s1 = require_cuda(scores1, "scores1").to(sdt).contiguous()
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
        _C.check((lib.lr_fuse_topk_f64 if f64 else lib.lr_fuse_topk)(
            s0.data_ptr(), i0.data_ptr(), k0, s1.data_ptr(), i1.data_ptr(), k1, Q,
            0 if method == "linear" else 1, float(weights[0]), float(weights[1]), float(eps),
            float(k_rrf), out_ids.data_ptr(), out_scores.data_ptr(), counts.data_ptr(), stream_ptr(dev)))
    return out_ids, out_scores, counts


# ---- dict-shaped fusion (the reference's call form) as an adapter over the device kernel
def _pack_systems(results_list: Sequence[dict], device):
    """Two ``dict[qid -> dict[pid -> score]]`` -> per system (scores f64 [Q, k_s], ids i64 [Q, k_s]) sorted by score
    descending with -1 padding, over a joint numbering of query and passage ids.  Scores stay float64 (Python floats)."""
    if len(results_list) != 2:
        raise NotImplementedError("device fusion combines two systems (dense + sparse), as HybridSearch does")
    qids: dict[str, int] = {}
    pids: dict[str, int] = {}
    for res in results_list:
        for qid in res:
            qids.setdefault(str(qid), len(qids))
    packed = []
    for res in results_list:
        width = max((len(p) for p in res.values()), default=0) or 1
        s = np.full((len(qids), width), -np.inf, dtype=np.float64)
        i = np.full((len(qids), width), -1, dtype=np.int64)
        for qid, passages in res.items():
            r = qids[str(qid)]
            if not passages:
                continue
            vals = np.fromiter((float(v) for v in passages.values()), dtype=np.float64, count=len(passages))
            ids = np.fromiter((pids.setdefault(str(p), len(pids)) for p in passages), dtype=np.int64, count=len(passages))
            order = np.argsort(-vals, kind="stable")  # rank = position in the system's sorted list (rrf)
            s[r, :len(order)], i[r, :len(order)] = vals[order], ids[order]
        packed.append((torch.from_numpy(s).to(device), torch.from_numpy(i).to(device)))
    return list(qids), list(pids), packed


def _fuse_dicts(results_list: Sequence[dict], method: str, weights, eps: float, k_rrf: float) -> dict:
    dev = torch.device("cuda", torch.cuda.current_device())
    qnames, pnames, ((s0, i0), (s1, i1)) = _pack_systems(results_list, dev)
    if not qnames:
        return {}
    ids, fused, _ = fuse_topk_device(s0, i0, s1, i1, method=method, weights=weights, eps=eps, k_rrf=k_rrf)
    return rows_to_dict(fused, ids, qnames, pnames)


def fuse_scores_rrf(results_list: Sequence[dict], k: int = 60) -> dict:
    """score_fuse_utils.py:3-46 (two systems), computed by ``lr_fuse_topk``."""
    return _fuse_dicts(results_list, "rrf", (1.0, 1.0), 1e-8, float(k))


def fuse_scores_linear(results_list: Sequence[dict], weights: Sequence[float] = (0.7, 0.3), eps: float = 1e-8) -> dict:
    """score_fuse_utils.py:48-90 (two systems), computed by ``lr_fuse_topk``."""
    assert len(results_list) == len(weights)
    return _fuse_dicts(results_list, "linear", weights, eps, 60.0)


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
        if score_fuse_method not in ("linear", "rrf"):
            raise NotImplementedError(f"score_fuse_method {score_fuse_method} is not supported.")
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

    # ---- array level: every result is (scores [Q, k], corpus positions [Q, k]) on device
    def _retrieve_systems(self, query_emb: dict, top_k: int, dense: bool, sparse: bool) -> dict:
        out = {}
        if dense:
            for key, name in (("dense_reps", "den"), ("emb_reps", "emb")):
                if query_emb.get(key) is not None:
                    out[name] = self.dense_search.retrieve_arrays(query_emb[key], top_k)
        if sparse:
            for key, name in (("token_id_reps", "tok"), ("sparse_reps", "spr")):
                if query_emb.get(key) is not None:
                    out[name] = self.sparse_search.retrieve_arrays(query_emb[key], top_k)
        return out

    def _fuse_arrays(self, a, b):
        ids, fused, _ = fuse_topk_device(a[0], a[1], b[0], b[1], method=self.score_fuse_method, weights=self.fuse_weights)
        return fused, ids

    def retrieve_arrays(self, query_emb: dict, top_k: int):
        """emb + tok (or den + spr) retrieval and fusion without leaving the device: (fused ids i64 [Q, 2k] as corpus
        positions, fused scores f64 [Q, 2k], counts i32 [Q]).  Both searchers must index the same corpus in the same order."""
        s0, i0 = self.dense_search.retrieve_arrays(query_emb["emb_reps"] if query_emb.get("emb_reps") is not None
                                                   else query_emb["dense_reps"], top_k)
        s1, i1 = self.sparse_search.retrieve_arrays(query_emb["token_id_reps"] if query_emb.get("token_id_reps") is not None
                                                    else query_emb["sparse_reps"], top_k)
        return fuse_topk_device(s0, i0, s1, i1, method=self.score_fuse_method, weights=self.fuse_weights)

    def _results_from_arrays(self, arrays: dict, query_ids: Sequence[str], dense_names, sparse_names) -> dict:
        """Fuse on device, then build every result dict once.  Sparse systems omit queries without a hit
        (anserini_search.py:208-214); fused results keep every query, like the union of the two dicts."""
        results = {}
        for name in ("den", "emb"):
            if name in arrays:
                results[name] = rows_to_dict(*arrays[name], query_ids, dense_names)
        for name in ("tok", "spr"):
            if name in arrays:
                results[name] = rows_to_dict(*arrays[name], query_ids, sparse_names, keep_empty=False)
        for d, s, name in (("den", "spr", "den_spr"), ("emb", "tok", "emb_tok")):
            if d in arrays and s in arrays:
                results[name] = rows_to_dict(*self._fuse_arrays(arrays[d], arrays[s]), query_ids, dense_names)
        return results

    def retrieve_with_emb(self, query_emb: dict, query_ids: Sequence[str], top_k: int, dense: bool = True,
                          sparse: bool = True, **kwargs) -> dict:
        """hybrid_search.py:123-180: per-system results and their fusion as ``dict[name -> dict[qid -> dict[pid -> score]]]``."""
        assert isinstance(query_emb, dict) and any(query_emb.get(k) is not None for k in
                                                   ("dense_reps", "sparse_reps", "emb_reps", "token_id_reps"))
        assert dense or sparse, "Please indicate retrieval embedding types."
        arrays = self._retrieve_systems(query_emb, top_k, dense, sparse)
        results = self._results_from_arrays(arrays, query_ids, self.dense_search._doc_name,
                                            self.sparse_search._corpus_ids)
        if "emb_tok" in results:
            results["default"] = results["emb_tok"]
        return results

    def _fuse_results(self, dense_results=None, sparse_results=None, weights=(0.7, 0.3)):
        """hybrid_search.py:207-232 for dict-shaped results (kept for callers that hold dicts)."""
        if dense_results is None and sparse_results is not None:
            return sparse_results
        if dense_results is not None and sparse_results is None:
            return dense_results
        if dense_results is None:
            raise ValueError("All scores are None. Please check model settings.")
        if self.score_fuse_method == "rrf":
            return fuse_scores_rrf([dense_results, sparse_results])
        return fuse_scores_linear([dense_results, sparse_results], weights=weights)

    def search(self, corpus: dict, queries: dict, top_k: int = 1000, score_function: str = None,
               return_sorted: bool = False, ignore_identical_ids: bool = False, **kwargs):
        """Chunk loop of hybrid_search.py:234-403: dense is indexed / retrieved per chunk and merged (on device), sparse is
        indexed per chunk and retrieved once at the end, then the systems are fused (on device)."""
        if not isinstance(queries, dict) or not isinstance(corpus, dict):
            raise NotImplementedError("HybridSearch.search takes dict corpora / queries")
        query_ids = list(queries.keys())
        qe = self.model.encode_queries([queries[qid] for qid in queries], batch_size=self.batch_size,
                                       show_progress_bar=self.show_progress_bar, convert_to_tensor=self.convert_to_tensor)
        kinds = {name: key for key, name in (("dense_reps", "den"), ("sparse_reps", "spr"), ("emb_reps", "emb"),
                                             ("token_id_reps", "tok")) if isinstance(qe, dict) and key in qe}
        assert kinds, "the model returned none of dense_reps / sparse_reps / emb_reps / token_id_reps"
        corpus_ids, corpus_list = sorted_corpus(corpus)
        self_pos = identical_positions(query_ids, corpus_ids) if ignore_identical_ids else None
        dense_kinds = [n for n in ("den", "emb") if n in kinds]
        sparse_kinds = [n for n in ("tok", "spr") if n in kinds]

        def chunks():
            for start in range(0, len(corpus_list), self.corpus_chunk_size):
                sub = self.model.encode_corpus(corpus_list[start:start + self.corpus_chunk_size],
                                               batch_size=self.batch_size, show_progress_bar=self.show_progress_bar,
                                               convert_to_tensor=self.convert_to_tensor)
                assert isinstance(sub, dict)
                if sparse_kinds:
                    assert "sparse_reps" in sub
                    self.sparse_search.index(sub["sparse_reps"], corpus_ids[start:start + self.corpus_chunk_size])
                if dense_kinds:
                    assert "dense_reps" in sub
                yield start, sub.get("dense_reps")

        arrays = {}
        if dense_kinds:
            # one pass over the chunks serves both dense query kinds: queries are stacked, results split afterwards
            q_all = torch.cat([torch.as_tensor(qe[kinds[n]]) for n in dense_kinds], dim=0)
            pos_all = None if self_pos is None else torch.cat([self_pos] * len(dense_kinds))
            s_all, i_all = self.dense_search.search_arrays(chunks(), q_all, top_k, pos_all)
            for j, n in enumerate(dense_kinds):
                arrays[n] = (s_all[j * len(query_ids):(j + 1) * len(query_ids)],
                             i_all[j * len(query_ids):(j + 1) * len(query_ids)])
        else:
            for _ in chunks():
                pass
        for n in sparse_kinds:
            arrays[n] = self.sparse_search.retrieve_arrays(qe[kinds[n]], top_k)
        results = self._results_from_arrays(arrays, query_ids, corpus_ids, corpus_ids)
        self._clear()
        default = None
        for name in ("den", "spr", "emb", "tok", "den_spr", "emb_tok"):  # the reference's precedence of the default result
            if name in results:
                default = results[name]
        return results if self.return_all_results else default
