"""Online-serving latency (BASELINE configs[4]): Qwen2.5-7B-shaped (d=3584, V=152064) EmbeddingBag encode + exact top-100
for batch 1 / 32, corpus row-sharded over the GPUs of the box (8.8M docs over 8 GPUs = 1.1M rows per GPU).
Per iteration: K1 encode -> K2 (warm-start + main pass + merge) -> [NCCL all-gather of keys + merge when world > 1].
Latency = CUDA events around one iteration on the current stream, max over ranks; p50 / p99 over --iters iterations.

    python tools/bench_latency.py [--docs-per-gpu 1100000]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_latency.py
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import lightretriever_b200 as lr  # noqa: E402
from lightretriever_b200.sharded import exchange_candidates  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs-per-gpu", type=int, default=1_100_000)
    ap.add_argument("--dim", type=int, default=3584)
    ap.add_argument("--vocab", type=int, default=152064)
    ap.add_argument("--topk", type=int, default=100)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--batches", type=int, nargs="+", default=[1, 32])
    ap.add_argument("--graph", type=int, default=1, help="1: replay the step as one CUDA graph; 0: eager launches")
    a = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator(device=dev).manual_seed(rank)
    table = (torch.randn(a.vocab, a.dim, device=dev, generator=torch.Generator(device=dev).manual_seed(0)) * 0.02).bfloat16()
    bag = lr.B200EmbeddingBag.from_pretrained(table, padding_idx=a.vocab - 1)
    n = a.docs_per_gpu
    corpus = torch.empty((n, a.dim), dtype=torch.bfloat16, device=dev)
    for c0 in range(0, n, 131072):
        m = min(131072, n - c0)
        corpus[c0:c0 + m] = torch.nn.functional.normalize(torch.randn(m, a.dim, device=dev, generator=g), dim=-1).bfloat16()
    lo = rank * n
    for B in a.batches:
        gq = torch.Generator().manual_seed(B)
        lens = torch.randint(1, 33, (B,), generator=gq)
        ids = torch.randint(0, a.vocab - 1, (int(lens.sum()),), generator=gq).to(dev)
        offs = torch.cumsum(torch.cat([torch.zeros(1, dtype=torch.long), lens[:-1]]), 0).to(dev)

        def step():
            qv = bag.encode(ids, offs, normalize=True, check_ids=False)
            if world == 1:
                return lr.flatip_topk(qv, corpus, a.topk, id_offset=lo)
            _, _, keys = lr.flatip_topk(qv, corpus, a.topk, id_offset=lo, return_keys=True)
            return lr.topk_merge(exchange_candidates(keys), a.topk)

        mode = "eager"
        if a.graph:  # one CUDA-graph replay per request (lightretriever_b200.online.OnlineSearcher)
            srv = lr.OnlineSearcher(bag, corpus, a.topk, batch=B, max_tokens=32 * B, id_offset=lo)
            eager_step, step = step, (lambda: srv.search(ids, offs))
            mode = "cuda_graph"
        for _ in range(10):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ts = []
        for _ in range(a.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = torch.tensor(ts, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t = t.sort().values
        if rank == 0:
            floor_ms = n * a.dim * 2 / (6547.8e9) * 1e3
            print(json.dumps({"config": "C5 online", "n_gpus": world, "docs_total": n * world, "docs_per_gpu": n, "dim": a.dim,
                              "batch": B, "k": a.topk, "mode": mode, "p50_ms": float(t[len(t) // 2]), "p99_ms": float(t[int(len(t) * 0.99)]),
                              "min_ms": float(t[0]), "hbm_floor_ms": floor_ms,
                              "frac_of_hbm_roofline_p50": floor_ms / float(t[len(t) // 2]), "qps_p50": B / float(t[len(t) // 2]) * 1e3}),
                  flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
