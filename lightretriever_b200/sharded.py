"""Row-sharded flat inner-product search across the GPUs of one box (one process per GPU).

The reference shards the corpus with Faiss ``GpuMultipleClonerOptions.shard = True`` and merges per-shard results on
the host (retriever/faiss_index.py:60-70), and additionally walks the corpus in chunks merged by a Python heap
(retriever/hybrid_search.py:301-344, 182-205).  Here every rank keeps one contiguous shard resident in HBM, runs the
fused scoring/top-k kernel on it, and only the per-shard top-k candidate keys ([Q, k] u64 = 8 bytes per candidate,
ids already global) cross NVLink: one ``all_gather_into_tensor`` (NCCL) followed by the on-device merge kernel.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous row shard [lo, hi) of rank `rank`; sizes differ by at most one row."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    return (n_total * rank) // world, (n_total * (rank + 1)) // world


def exchange_candidates(local_keys: torch.Tensor, group=None) -> torch.Tensor:
    """all-gather of the per-shard candidate keys: [Q, k] int64 -> [world, Q, k] (same on every rank)."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    local_keys = local_keys.contiguous()
    if world == 1:
        return local_keys.unsqueeze(0)
    out = torch.empty((world,) + tuple(local_keys.shape), dtype=local_keys.dtype, device=local_keys.device)
    if local_keys.device.type == "cuda":
        dist.all_gather_into_tensor(out, local_keys, group=group)
    else:  # gloo (CPU tests of the host logic)
        parts = [torch.empty_like(local_keys) for _ in range(world)]
        dist.all_gather(parts, local_keys, group=group)
        out = torch.stack(parts, dim=0)
    return out


class ShardedFlatIPIndex:
    """One shard per rank; ``search_device`` returns the exact global top-k on every rank."""

    def __init__(self, dim: int, n_total: int, device: Optional[torch.device] = None, group=None):
        from .search import FlatIPIndex  # requires the CUDA library

        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = int(n_total)
        self.lo, self.hi = shard_range(self.n_total, self.rank, self.world)
        if self.n_total >= (1 << 32) - 2:
            raise ValueError("global ids must stay below 2^32")
        self.local = FlatIPIndex(dim, device=device, id_offset=self.lo)

    def add_local(self, emb) -> None:
        """Append rows of this rank's shard (global row ids lo .. hi-1, in order)."""
        self.local.add(emb)
        if self.local.ntotal > self.hi - self.lo:
            raise ValueError("more rows than this rank's shard holds")

    def search_device(self, query: torch.Tensor, k: int, d_used: Optional[int] = None):
        from .search import flatip_topk, topk_merge

        q = query.to(device=self.local.device, dtype=torch.bfloat16)
        if self.local.ntotal != self.hi - self.lo:
            raise RuntimeError(f"shard incomplete: {self.local.ntotal} of {self.hi - self.lo} rows")
        if self.world == 1:
            return flatip_topk(q, self.local.corpus, k, d_used=d_used, id_offset=self.lo)
        _, _, keys = flatip_topk(q, self.local.corpus, k, d_used=d_used, id_offset=self.lo, return_keys=True)
        gathered = exchange_candidates(keys, self.group)  # [world, Q, k]
        return topk_merge(gathered, k)


class ShardedImpactIndex:
    """Document-sharded sparse impact index: the sparse twin of ``ShardedFlatIPIndex`` ("same for CSR postings
    — shard by doc, per-shard inverted index").  Rank r indexes documents ``[lo, hi)`` as its own token-major inverted index
    (local doc ids, ``id_offset = lo``); a search runs ``lr_sparse_score_topk`` on the shard and only the per-shard top-k
    keys ``[Q, k]`` u64 (integer score << 32 | ~global id) cross NVLink, then ``lr_topk_merge`` with integer scores.  The
    reference has no multi-GPU sparse path (Anserini is one JVM, retriever/anserini_search.py:143-216); the exchange is the
    one the dense path uses."""

    def __init__(self, vocab_size: int, n_total: int, device: Optional[torch.device] = None, group=None):
        from .sparse_search import ImpactIndex  # requires the CUDA library

        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = int(n_total)
        self.lo, self.hi = shard_range(self.n_total, self.rank, self.world)
        if self.n_total >= (1 << 32) - 2:
            raise ValueError("global ids must stay below 2^32")
        self.local = ImpactIndex(vocab_size, device=device, id_offset=self.lo)

    def add_local_csr(self, indptr, tok, imp) -> None:
        """Append documents of this rank's shard (global ids lo .. hi-1, in order) as doc-major CSR."""
        self.local.add_csr(indptr, tok, imp)
        if self.local.N > self.hi - self.lo:
            raise ValueError("more documents than this rank's shard holds")

    def search_device(self, q_indptr, q_tok, q_cnt, k: int):
        from . import _C
        from .search import topk_merge

        if self.local.N != self.hi - self.lo:
            raise RuntimeError(f"shard incomplete: {self.local.N} of {self.hi - self.lo} documents")
        if self.world == 1:
            return self.local.search_device(q_indptr, q_tok, q_cnt, k)
        _, _, keys = self.local.search_device(q_indptr, q_tok, q_cnt, k, return_keys=True)
        gathered = exchange_candidates(keys, self.group)  # [world, Q, k]
        return topk_merge(gathered, k, score_kind=_C.LR_SCORE_U32)
