"""Exact flat inner-product search (K2) — host side, seam S2.

Mirrors
  * ``FaissIndex`` (reference retriever/faiss_index.py:20-73): ``build / search / to_gpu / reset / save``;
    ``search(q, k) -> (scores float32 [Q,k] descending, ids int64 [Q,k])``;
  * ``FlatIPFaissSearch`` / ``DenseRetrievalFaissSearch`` (retriever/faiss_search.py:46-293, 477-510):
    ``index(corpus_emb, corpus_ids)``, ``retrieve_with_emb(query_emb, query_ids, top_k)``, ``_clear()``, ``search``.
The corpus lives in HBM as bf16 rows; scoring + top-k is one fused tcgen05 kernel plus a merge
(csrc/umma_gemm.cuh, csrc/topk_merge.cu).  Differences from Faiss that are part of the contract:
  * equal scores are ordered by ascending id (Faiss: unspecified);
  * when fewer than k documents exist the tail is (score -inf, id -1) and is dropped from result dicts
    (the reference lets numpy wrap id -1 to the last passage, SURVEY §8c trap 8).
"""
from __future__ import annotations

import csv
import heapq
import logging
import os
import time
from typing import Optional, Sequence

import numpy as np
import torch

from . import _C
from ._util import Workspace, require_cuda, stream_ptr

logger = logging.getLogger(__name__)

_WS = Workspace()


def flatip_topk(query: torch.Tensor, corpus: torch.Tensor, k: int, d_used: Optional[int] = None,
                q_scale: Optional[torch.Tensor] = None, c_scale: Optional[torch.Tensor] = None,
                id_offset: int = 0, return_keys: bool = False, workspace: Optional[Workspace] = None):
    """scores, ids (and optionally sorted u64 keys) of the k largest ``query @ corpus.T`` per query.

    query [Q, >=d_used] bf16, corpus [N, >=d_used] bf16, both on the same B200; rows may be strided views
    (e.g. an MRL prefix ``x[:, :m]`` of full-width vectors) as long as the inner stride is 1.
    """
    q = require_cuda(query, "query")
    c = require_cuda(corpus, "corpus")
    if q.dtype != torch.bfloat16 or c.dtype != torch.bfloat16:
        raise ValueError("flatip_topk takes bfloat16 query and corpus (convert once at index / encode time)")
    if q.ndim != 2 or c.ndim != 2:
        raise ValueError("query and corpus must be 2-D")
    if q.stride(1) != 1:
        q = q.contiguous()
    if c.stride(1) != 1:
        c = c.contiguous()
    d = int(d_used) if d_used else min(q.shape[1], c.shape[1])
    if d_used is None and q.shape[1] != c.shape[1]:
        raise ValueError(f"dimension mismatch: query {q.shape[1]} vs corpus {c.shape[1]}")
    Q, N = q.shape[0], c.shape[0]
    dev = q.device
    if c.device != dev:
        raise ValueError("query and corpus must be on the same device")
    lib = _C.load()
    ws = (workspace or _WS).get(lib.lr_flatip_workspace_bytes_for(Q, N, k, d), dev)
    scores = torch.empty((Q, k), dtype=torch.float32, device=dev)
    ids = torch.empty((Q, k), dtype=torch.int64, device=dev)
    keys = torch.empty((Q, k), dtype=torch.int64, device=dev) if return_keys else None
    for name, s, n in (("q_scale", q_scale, Q), ("c_scale", c_scale, N)):
        if s is not None and (s.dtype != torch.float32 or s.numel() != n or not s.is_contiguous() or s.device != dev):
            raise ValueError(f"{name} must be a contiguous float32 tensor with {n} elements on {dev}")
    with torch.cuda.device(dev):
        _C.check(lib.lr_flatip_topk(
            q.data_ptr(), q.stride(0), c.data_ptr(), c.stride(0), Q, N, d,
            None if q_scale is None else q_scale.data_ptr(), None if c_scale is None else c_scale.data_ptr(),
            int(id_offset), int(k), scores.data_ptr(), ids.data_ptr(),
            None if keys is None else keys.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr(dev)))
    return (scores, ids, keys) if return_keys else (scores, ids)


def flatip_scores(query: torch.Tensor, corpus: torch.Tensor, d_used: Optional[int] = None) -> torch.Tensor:
    """Full [Q, N] f32 score matrix through the same TMA + tcgen05 main loop (parity/debug; small shapes)."""
    q = require_cuda(query, "query")
    c = require_cuda(corpus, "corpus")
    if q.dtype != torch.bfloat16 or c.dtype != torch.bfloat16:
        raise ValueError("flatip_scores takes bfloat16 inputs")
    d = int(d_used) if d_used else q.shape[1]
    out = torch.empty((q.shape[0], c.shape[0]), dtype=torch.float32, device=q.device)
    lib = _C.load()
    with torch.cuda.device(q.device):
        _C.check(lib.lr_flatip_scores(q.data_ptr(), q.stride(0), c.data_ptr(), c.stride(0), q.shape[0], c.shape[0], d,
                                      out.data_ptr(), stream_ptr(q.device)))
    return out


def topk_merge(keys: torch.Tensor, k: int, counts: Optional[torch.Tensor] = None, score_kind: int = _C.LR_SCORE_F32,
               id_offset: int = 0, return_keys: bool = False):
    """Exact top-k of L candidate lists per query.  keys: int64/uint64-bit [L, Q, cap] (0 = empty slot)."""
    keys = require_cuda(keys, "keys")
    if keys.ndim != 3 or keys.dtype != torch.int64 or not keys.is_contiguous():
        raise ValueError("keys must be a contiguous int64 [L, Q, cap] tensor of candidate keys")
    L, Q, cap = keys.shape
    dev = keys.device
    if counts is not None:
        counts = counts.to(torch.int32).contiguous()
        if counts.shape != (L, Q) or counts.device != dev:
            raise ValueError("counts must be [L, Q] on the keys' device")
    scores = torch.empty((Q, k), dtype=torch.float32, device=dev)
    ids = torch.empty((Q, k), dtype=torch.int64, device=dev)
    okeys = torch.empty((Q, k), dtype=torch.int64, device=dev) if return_keys else None
    lib = _C.load()
    with torch.cuda.device(dev):
        _C.check(lib.lr_topk_merge(keys.data_ptr(), None if counts is None else counts.data_ptr(), L, Q, Q, cap, int(k),
                                   score_kind, int(id_offset), scores.data_ptr(), ids.data_ptr(),
                                   None if okeys is None else okeys.data_ptr(), stream_ptr(dev)))
    return (scores, ids, okeys) if return_keys else (scores, ids)


def encode_keys(scores: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    scores = require_cuda(scores, "scores").to(torch.float32).contiguous()
    ids = require_cuda(ids, "ids").to(torch.int64).contiguous()
    keys = torch.empty(scores.shape, dtype=torch.int64, device=scores.device)
    lib = _C.load()
    with torch.cuda.device(scores.device):
        _C.check(lib.lr_encode_keys(scores.data_ptr(), ids.data_ptr(), scores.numel(), keys.data_ptr(),
                                    stream_ptr(scores.device)))
    return keys


class FlatIPIndex:
    """HBM-resident flat inner-product index with the ``FaissIndex`` surface (faiss_index.py:20-73)."""

    def __init__(self, dim: Optional[int] = None, passage_ids: Optional[Sequence[int]] = None,
                 device: Optional[torch.device] = None, id_offset: int = 0):
        self.dim = dim
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._chunks: list[torch.Tensor] = []
        self._corpus: Optional[torch.Tensor] = None
        self._passage_ids = None if passage_ids is None else np.asarray(passage_ids, dtype=np.int64)
        self.id_offset = int(id_offset)

    # faiss.IndexFlatIP.add
    def add(self, emb) -> None:
        t = emb if isinstance(emb, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(emb))
        if t.ndim != 2:
            raise ValueError("embeddings must be [n, d]")
        if self.dim is None:
            self.dim = t.shape[1]
        if t.shape[1] != self.dim:
            raise ValueError(f"dimension mismatch: index {self.dim}, got {t.shape[1]}")
        self._chunks.append(t.to(device=self.device, dtype=torch.bfloat16, non_blocking=True))
        self._corpus = None

    @property
    def ntotal(self) -> int:
        return sum(c.shape[0] for c in self._chunks)

    @property
    def corpus(self) -> torch.Tensor:
        if self._corpus is None:
            if not self._chunks:
                raise RuntimeError("index is empty")
            self._corpus = self._chunks[0] if len(self._chunks) == 1 else torch.cat(self._chunks, dim=0)
            self._chunks = [self._corpus]
        return self._corpus

    @classmethod
    def build(cls, passage_ids: Sequence[int], passage_embeddings, index: Optional["FlatIPIndex"] = None,
              buffer_size: int = 50000) -> "FlatIPIndex":
        if index is None:
            index = cls(passage_embeddings.shape[1])
        for start in range(0, len(passage_ids), buffer_size):
            index.add(passage_embeddings[start:start + buffer_size])
        index._passage_ids = np.asarray(passage_ids, dtype=np.int64)
        return index

    def search_device(self, query_embeddings: torch.Tensor, k: int, d_used: Optional[int] = None, **kw):
        q = query_embeddings
        if not isinstance(q, torch.Tensor):
            q = torch.from_numpy(np.ascontiguousarray(q))
        q = q.to(device=self.device, dtype=torch.bfloat16, non_blocking=True)
        return flatip_topk(q, self.corpus, k, d_used=d_used, id_offset=self.id_offset, **kw)

    def search(self, query_embeddings, k: int, **kwargs) -> tuple[np.ndarray, np.ndarray]:
        num_queries = query_embeddings.shape[0]
        start_time = time.time()
        scores, ids = self.search_device(query_embeddings, k)
        scores_arr = scores.cpu().numpy()
        ids_arr = ids.cpu().numpy()
        if self._passage_ids is not None:
            valid = ids_arr >= 0
            mapped = self._passage_ids[np.where(valid, ids_arr - self.id_offset, 0).reshape(-1)].reshape(num_queries, -1)
            ids_arr = np.where(valid, mapped, -1)
        total_time = time.time() - start_time
        logger.info("Num of queries: %d\tSearch time (s): %.3f\tQPS: %.3f", num_queries, total_time,
                    num_queries / max(total_time, 1e-9))
        return scores_arr, ids_arr

    def to_gpu(self):
        return self  # already HBM-resident

    def reset(self) -> None:
        self._chunks = []
        self._corpus = None

    def save(self, fname: str) -> None:
        torch.save({"corpus": self.corpus.cpu(), "passage_ids": self._passage_ids, "id_offset": self.id_offset}, fname)

    @classmethod
    def load(cls, fname: str, device: Optional[torch.device] = None) -> "FlatIPIndex":
        blob = torch.load(fname, map_location="cpu", weights_only=False)
        idx = cls(blob["corpus"].shape[1], blob["passage_ids"], device=device, id_offset=blob.get("id_offset", 0))
        idx.add(blob["corpus"])
        return idx


def save_dict_to_tsv(_dict: dict, output_path: str, keys=()):
    with open(output_path, "w") as f:
        writer = csv.writer(f, delimiter="\t", quoting=csv.QUOTE_MINIMAL)
        if keys:
            writer.writerow(keys)
        for key, value in _dict.items():
            writer.writerow([key, value])


def load_tsv_to_dict(input_path: str, header: bool = True) -> dict:
    mappings = {}
    with open(input_path, encoding="utf-8") as f:
        reader = csv.reader(f, delimiter="\t", quoting=csv.QUOTE_MINIMAL)
        if header:
            next(reader)
        for row in reader:
            mappings[row[0]] = int(row[1])
    return mappings


class FlatIPSearch:
    """``FlatIPFaissSearch`` surface (faiss_search.py:46-293, 477-510) over ``FlatIPIndex``."""

    def __init__(self, model=None, batch_size: int = 128, corpus_chunk_size: Optional[int] = None,
                 use_single_gpu: bool = False, use_multiple_gpu: bool = False, **kwargs):
        self.model = model
        self.batch_size = batch_size
        self.corpus_chunk_size = batch_size * 800 if corpus_chunk_size is None else corpus_chunk_size
        self.show_progress_bar = kwargs.get("show_progress_bar", True)
        self.convert_to_tensor = kwargs.get("convert_to_tensor", True)
        self.mapping_tsv_keys = ["beir-docid", "faiss-docid"]
        self.faiss_index: Optional[FlatIPIndex] = None
        self.use_single_gpu = use_single_gpu
        self.use_multiple_gpu = use_multiple_gpu
        self.dim_size = None
        self.mapping: dict = {}
        self.rev_mapping: dict = {}
        self.device = kwargs.get("device", None)

    @classmethod
    def name(cls):
        return "flat_ip_b200_search"

    def get_index_name(self):
        return "flat_faiss_index"

    def encode(self, sentences, batch_size: int, **kwargs):
        return self.model.encode(sentences=sentences, batch_size=batch_size, **kwargs)

    def encode_queries(self, queries, batch_size: int, **kwargs):
        return self.model.encode_queries(queries=queries, batch_size=batch_size, **kwargs)

    def encode_corpus(self, corpus, batch_size: int, **kwargs):
        return self.model.encode_corpus(corpus=corpus, batch_size=batch_size, **kwargs)

    def _create_mapping_ids(self, corpus_ids):
        if not all(isinstance(doc_id, int) for doc_id in corpus_ids):
            for idx in range(len(corpus_ids)):
                self.mapping[corpus_ids[idx]] = idx
                self.rev_mapping[idx] = corpus_ids[idx]

    def _clear(self):
        if self.faiss_index is not None:
            self.faiss_index.reset()
            del self.faiss_index
        self.faiss_index = None
        self.dim_size = None
        self.mapping = {}
        self.rev_mapping = {}

    def index(self, corpus_emb, corpus_ids: Sequence[str]):
        self._create_mapping_ids(corpus_ids)
        self.dim_size = corpus_emb.shape[1]
        faiss_ids = [self.mapping.get(cid) for cid in corpus_ids] if self.mapping else list(corpus_ids)
        if self.mapping:
            passage_ids = faiss_ids
        else:
            passage_ids = [int(c) for c in corpus_ids]
        base = FlatIPIndex(self.dim_size, device=self.device)
        self.faiss_index = FlatIPIndex.build(passage_ids, corpus_emb, base)

    def retrieve_with_emb(self, query_emb, query_ids: Sequence[str], top_k: int, **kwargs) -> dict:
        if self.faiss_index is None:
            raise RuntimeError("index() must be called before retrieve_with_emb()")
        scores_arr, ids_arr = self.faiss_index.search(query_emb, top_k, **kwargs)
        results: dict[str, dict[str, float]] = {}
        for i in range(len(query_ids)):
            row = {}
            for doc_id, score in zip(ids_arr[i].tolist(), scores_arr[i].tolist()):
                if doc_id < 0:
                    continue  # fewer than k documents: (-inf, -1) padding is dropped
                row[self.rev_mapping[doc_id] if self.rev_mapping else str(doc_id)] = float(score)
            results[query_ids[i]] = row
        return results

    def retrieve_arrays(self, query_emb, top_k: int, **kwargs):
        """Device arrays (scores f32 [Q,k], doc positions i64 [Q,k], -1 padded): the f2 row — results stay on device and
        are turned into the reference's nested dicts only when a caller asks for them (`arrays_to_dict`)."""
        if self.faiss_index is None:
            raise RuntimeError("index() must be called before retrieve_arrays()")
        return self.faiss_index.search_device(query_emb, top_k, **kwargs)

    def arrays_to_dict(self, scores, ids, query_ids: Sequence[str]) -> dict:
        s, i = scores.cpu().numpy(), ids.cpu().numpy()
        pid = self.faiss_index._passage_ids
        out = {}
        for r, qid in enumerate(query_ids):
            row = {}
            for doc, sc in zip(i[r].tolist(), s[r].tolist()):
                if doc < 0:
                    continue
                doc = int(pid[doc - self.faiss_index.id_offset]) if pid is not None else doc
                row[self.rev_mapping[doc] if self.rev_mapping else str(doc)] = float(sc)
            out[qid] = row
        return out

    def save(self, output_dir: str, prefix: str = "my-index", ext: str = "flat"):
        save_dict_to_tsv(self.mapping, os.path.join(output_dir, f"{prefix}.{ext}.tsv"), keys=self.mapping_tsv_keys)
        self.faiss_index.save(os.path.join(output_dir, f"{prefix}.{ext}.lrb200"))

    def load(self, input_dir: str, prefix: str = "my-index", ext: str = "flat"):
        self.mapping = load_tsv_to_dict(os.path.join(input_dir, f"{prefix}.{ext}.tsv"), header=True)
        self.rev_mapping = {v: k for k, v in self.mapping.items()}
        self.faiss_index = FlatIPIndex.load(os.path.join(input_dir, f"{prefix}.{ext}.lrb200"), device=self.device)
        self.dim_size = self.faiss_index.dim

    def search(self, corpus, queries, top_k: int = 1000, score_function: str = None, return_sorted: bool = False,
               ignore_identical_ids: bool = False, **kwargs) -> dict:
        """Chunked encode -> index -> retrieve -> merge, as DenseRetrievalFaissSearch.search (faiss_search.py:176-293)."""
        if not isinstance(queries, dict) or not isinstance(corpus, dict):
            raise NotImplementedError("FlatIPSearch.search takes dict corpora / queries")
        query_ids = list(queries.keys())
        queries_list = [queries[qid] for qid in queries]
        query_embeddings = self.model.encode_queries(queries_list, batch_size=self.batch_size,
                                                     show_progress_bar=self.show_progress_bar,
                                                     convert_to_tensor=self.convert_to_tensor)
        corpus_ids = sorted(corpus, key=lambda k_: len(corpus[k_].get("text", "")) if isinstance(corpus[k_], dict)
                            else len(corpus[k_]), reverse=True)
        corpus_list = [corpus[cid] for cid in corpus_ids]
        heaps: dict[str, list] = {qid: [] for qid in query_ids}
        for start in range(0, len(corpus_list), self.corpus_chunk_size):
            end = min(start + self.corpus_chunk_size, len(corpus_list))
            sub = self.model.encode_corpus(corpus_list[start:end], batch_size=self.batch_size,
                                           show_progress_bar=self.show_progress_bar,
                                           convert_to_tensor=self.convert_to_tensor)
            if isinstance(sub, dict):
                if "dense_reps" not in sub:
                    raise ValueError(f"HybridModel: Return Multi-vector with keys {sub.keys()}, but not `dense_reps`")
                sub = sub["dense_reps"]
            self.index(sub, corpus_ids[start:end])
            sub_results = self.retrieve_with_emb(query_embeddings, query_ids, top_k=top_k)
            self._clear()
            add_to_heap(sub_results, heaps, top_k, ignore_identical_ids)
        return {qid: {pid: score for score, pid in heaps[qid]} for qid in heaps}


def add_to_heap(sub_results: dict, result_heaps: dict, top_k: int, ignore_identical_ids: bool) -> dict:
    """Host merge with the reference's semantics (hybrid_search.py:182-205) for dict-shaped per-chunk results."""
    for qid, pid_to_score in sub_results.items():
        heap = result_heaps.setdefault(qid, [])
        for pid, score in pid_to_score.items():
            if ignore_identical_ids and qid == pid:
                continue
            if len(heap) < top_k:
                heapq.heappush(heap, (score, pid))
            else:
                heapq.heappushpop(heap, (score, pid))
    return result_heaps
