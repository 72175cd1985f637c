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
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_exchange_merge(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
