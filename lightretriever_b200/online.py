"""Online serving step (BASELINE configs[4]: batch 1 / 32, p50 / p99 latency) as ONE CUDA-graph replay.

The reference answers an online query with a Python call chain — EmbeddingBag forward (finetune/modeling_hybrid.py:474),
``F.normalize`` (:489-490), ``FaissIndex.search`` (retriever/faiss_index.py:27-40) and, over shards, a host merge
(:65-68).  Here the same chain is five to eight kernels and at most one NCCL all-gather; at ~1.5 ms per request the
host's launch path is a visible part of the latency, so the whole step is captured once per (batch, max_tokens) shape
and replayed: the request's ids / offsets are copied into static device buffers, unused id slots hold ``padding_idx``
(skipped by the encoder exactly like torch's ``padding_idx``) and unused bags are empty.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .encode import B200EmbeddingBag
from .search import flatip_topk, topk_merge
from .sharded import exchange_candidates


class OnlineSearcher:
    def __init__(self, bag: B200EmbeddingBag, corpus: torch.Tensor, k: int, batch: int, max_tokens: int,
                 id_offset: int = 0, shrink_dim: Optional[int] = None, group=None, use_graph: bool = True):
        if bag.padding_idx is None:
            raise ValueError("OnlineSearcher pads requests with padding_idx: the bag needs one")
        if corpus.dtype != torch.bfloat16 or not corpus.is_cuda:
            raise ValueError("corpus must be a bfloat16 CUDA tensor (the resident shard)")
        self.bag, self.corpus, self.k, self.batch, self.max_tokens = bag, corpus, int(k), int(batch), int(max_tokens)
        self.id_offset, self.shrink_dim, self.group = int(id_offset), shrink_dim, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        dev = corpus.device
        self._ids = torch.full((self.max_tokens,), bag.padding_idx, dtype=torch.int64, device=dev)
        self._offs = torch.full((self.batch,), self.max_tokens, dtype=torch.int64, device=dev)
        self._graph = None
        self._out = None
        if use_graph:
            self._capture()

    def _step(self):
        qv = self.bag.encode(self._ids, self._offs, shrink_dim=self.shrink_dim, normalize=True, check_ids=False)
        corpus = self.corpus if self.shrink_dim is None else self.corpus[:, :self.shrink_dim]
        if self.world == 1:
            return flatip_topk(qv, corpus, self.k, id_offset=self.id_offset)
        _, _, keys = flatip_topk(qv, corpus, self.k, id_offset=self.id_offset, return_keys=True)
        return topk_merge(exchange_candidates(keys, self.group), self.k)

    def _capture(self):
        side = torch.cuda.Stream(device=self.corpus.device)
        side.wait_stream(torch.cuda.current_stream(self.corpus.device))
        with torch.cuda.stream(side):  # warm-up: workspace allocation, attribute calls, NCCL channels
            for _ in range(3):
                self._step()
        torch.cuda.current_stream(self.corpus.device).wait_stream(side)
        torch.cuda.synchronize(self.corpus.device)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._out = self._step()

    def search(self, input_ids: torch.Tensor, offsets: torch.Tensor):
        """(scores [n_bags, k] f32, ids [n_bags, k] i64) for a request of up to ``batch`` bags / ``max_tokens`` ids.
        The returned tensors are views of the searcher's static outputs: consume them before the next call."""
        n_ids, n_bags = input_ids.numel(), offsets.numel()
        if n_bags > self.batch or n_ids > self.max_tokens:
            raise ValueError(f"request ({n_bags} bags, {n_ids} ids) exceeds the captured shape "
                             f"({self.batch}, {self.max_tokens})")
        self._ids.fill_(self.bag.padding_idx)
        self._ids[:n_ids].copy_(input_ids, non_blocking=True)
        self._offs.fill_(n_ids)  # unused bags are empty; the last real bag ends where the padding ids start
        self._offs[:n_bags].copy_(offsets, non_blocking=True)
        if self._graph is not None:
            self._graph.replay()
            s, i = self._out
        else:
            s, i = self._step()
        return s[:n_bags], i[:n_bags]
