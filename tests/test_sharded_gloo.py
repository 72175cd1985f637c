"""world_size-2 gloo test of the N>1 host path: row shards -> per-shard top-k -> all-gather of candidate keys -> merge.
The per-shard scoring and the merge are played by the oracle here (no GPU); what is under test is the sharding and the
exchange step of lightretriever_b200.sharded, which are the same code the NCCL path runs."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lightretriever_b200.sharded import exchange_candidates, shard_range
        from oracle import oracle
        rng = np.random.default_rng(0)  # same data on every rank
        N, d, Q, k = 1003, 16, 5, 20
        corpus = rng.standard_normal((N, d)).astype(np.float32)
        corpus[500:520] = corpus[3]  # ties that straddle the shard boundary
        q = rng.standard_normal((Q, d)).astype(np.float32)
        lo, hi = shard_range(N, rank, world)
        s, i = oracle.flatip_topk(q, corpus[lo:hi], k, id_offset=lo)
        keys = torch.from_numpy(oracle.encode_keys(s, i).view(np.int64))
        gathered = exchange_candidates(keys)  # [world, Q, k]
        assert gathered.shape == (world, Q, k)
        g = gathered.numpy().view(np.uint64)
        parts = [oracle.decode_keys(g[r]) for r in range(world)]
        ms, mi = oracle.merge_topk([p[0] for p in parts], [p[1] for p in parts], k)
        fs, fi = oracle.flatip_topk(q, corpus, k)
        np.testing.assert_array_equal(mi, fi)
        np.testing.assert_array_equal(ms, fs)
        # sparse twin: document-sharded impact search, integer scores, same exchange
        V, nd = 60, 403
        docs = [{int(t): int(rng.integers(1, 300)) for t in rng.choice(V, size=int(rng.integers(0, 9)), replace=False)}
                for _ in range(nd)]
        queries = [{int(t): int(rng.integers(1, 3)) for t in rng.choice(V, size=6, replace=False)} for _ in range(4)]
        lo, hi = shard_range(nd, rank, world)
        ss, si = oracle.impact_topk(queries, docs[lo:hi], 15, id_offset=lo)
        skeys = np.where(si >= 0, (ss.astype(np.uint64) << np.uint64(32)) |
                         (np.uint64(0xFFFFFFFF) - np.where(si >= 0, si, 0).astype(np.uint64)), np.uint64(0))
        g2 = exchange_candidates(torch.from_numpy(skeys.view(np.int64))).numpy().view(np.uint64)
        assert g2.shape == (world, 4, 15)
        sc = [np.where(g2[r] != 0, (g2[r] >> np.uint64(32)).astype(np.float32), -np.inf).astype(np.float32)
              for r in range(world)]
        ids = [np.where(g2[r] != 0, (np.uint64(0xFFFFFFFF) - (g2[r] & np.uint64(0xFFFFFFFF))).astype(np.int64), -1)
               for r in range(world)]
        ms2, mi2 = oracle.merge_topk(sc, ids, 15)
        fs2, fi2 = oracle.impact_topk(queries, docs, 15)
        np.testing.assert_array_equal(mi2, fi2)
        np.testing.assert_array_equal(ms2, fs2)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_exchange_merge(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
