"""Worker of tests/test_multigpu_nccl.py: one process per GPU under torch.distributed.run (NCCL).  Row-sharded dense
search, document-sharded sparse search and the searcher API with use_multiple_gpu=True, each against the UNSHARDED oracle
on the same seeded inputs.  Every rank checks the global result (it must be identical on all ranks)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, local, world = (int(os.environ[k]) for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import lightretriever_b200 as lr
    from lightretriever_b200.sparse_search import parse_queries
    from oracle import oracle

    # ---- dense: ShardedFlatIPIndex.search_device vs the unsharded fp32 oracle (ties straddle the shard boundary)
    gen = torch.Generator().manual_seed(3)  # same data on every rank
    N, d, Q, k = 50_001, 256, 300, 100
    corpus = F.normalize(torch.randn(N, d, generator=gen), dim=-1).bfloat16()
    mid = N // world
    corpus[mid - 4:mid + 4] = corpus[11]
    q = F.normalize(torch.randn(Q, d, generator=gen), dim=-1).bfloat16()
    sh = lr.ShardedFlatIPIndex(d, N, device=dev)
    sh.add_local(corpus[sh.lo:sh.hi])
    s, i = sh.search_device(q, k)
    ref = (q.float() @ corpus.float().T).numpy()
    oracle.check_topk_parity(s.cpu().numpy(), i.cpu().numpy(), ref, k, rtol=1e-2)
    es, ei = oracle.flatip_topk(q.float(), corpus.float(), k)
    assert (ei == i.cpu().numpy()).mean() > 0.999
    # identical on every rank
    mine = torch.stack([s.double(), i.double()])
    other = mine.clone()
    dist.broadcast(other, src=0)
    assert torch.equal(mine, other), "ranks disagree on the merged result"

    # ---- the same with a warm-start pass in the plan: every rank scores 1/world of the prefix, the prefix top-k are
    #      exchanged and the merged k-th best seeds all ranks (lr_flatip_topk_begin / _finish)
    os.environ["LR_FLATIP_PREFIX_DOCS"] = "4096"
    lr._C.reload_env()
    s2, i2 = sh.search_device(q, k)
    oracle.check_topk_parity(s2.cpu().numpy(), i2.cpu().numpy(), ref, k, rtol=1e-2)
    assert torch.equal(i2, i) and torch.equal(s2, s), "shared warm start changed the result"
    del os.environ["LR_FLATIP_PREFIX_DOCS"]
    lr._C.reload_env()

    # ---- the searcher API: FlatIPSearch(use_multiple_gpu=True) == single-GPU searcher, chunked search() included
    cids = [f"doc-{j}" for j in range(3000)]
    qids = [f"q{j}" for j in range(20)]
    multi = lr.FlatIPSearch(model=None, use_multiple_gpu=True)
    single = lr.FlatIPSearch(model=None)
    multi.index(corpus[:3000].float(), cids)
    single.index(corpus[:3000].float(), cids)
    assert type(multi.faiss_index).__name__ == "ShardedFlatIPIndex" and multi.faiss_index.world == world
    assert multi.retrieve_with_emb(q[:20].float(), qids, 30) == single.retrieve_with_emb(q[:20].float(), qids, 30)
    chunks = [(lo, corpus[lo:lo + 1100].float()) for lo in range(0, 3000, 1100)]
    sm, im = multi.search_arrays(iter(chunks), q[:20].float(), 30)
    ss, is_ = single.search_arrays(iter(chunks), q[:20].float(), 30)
    assert torch.equal(im, is_) and torch.equal(sm, ss)

    # ---- sparse: ShardedImpactIndex.search_device vs the unsharded int64 oracle (bit-exact)
    rng = np.random.default_rng(5)
    V, nd, ks = 300, 20_003, 50
    docs = [{int(t): int(rng.integers(1, 400)) for t in rng.choice(V, size=int(rng.integers(0, 20)), replace=False)}
            for _ in range(nd)]
    queries = [" ".join(str(int(t)) for t in rng.integers(0, V, size=int(rng.integers(1, 33)))) for _ in range(40)]
    from lightretriever_b200.sparse_search import json_to_csr
    shi = lr.ShardedImpactIndex(V, nd, device=dev)
    shi.add_local_csr(*json_to_csr([{str(a): b for a, b in dd.items()} for dd in docs[shi.lo:shi.hi]]))
    gs, gi = shi.search_device(*parse_queries(queries, V), ks)
    es, ei = oracle.impact_topk([oracle.query_counts([int(t) for t in s_.split()]) for s_ in queries], docs, ks)
    assert (gi.cpu().numpy() == ei).all() and (gs.cpu().numpy() == es).all()

    dist.barrier()
    if rank == 0:
        print(f"nccl worker OK world={world}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
