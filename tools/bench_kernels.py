"""Per-kernel measurements beside the headline bench: K1 (EmbeddingBag encode, HBM-bound), K3 (sparse head, tensor-bound),
K4 (sparse impact scoring, HBM-bound), K2 at the online shapes of BASELINE configs[4] (batch 1 / 32, HBM-bound) and the MRL
widths of configs[2].  CUDA events on the launching stream, >=3 warm-ups, inputs larger than L2 or L2 flushed between
iterations.  Prints one JSON object per line.  Usage: python tools/bench_kernels.py [k1 k3 k4 k2small mrl]"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import lightretriever_b200 as lr  # noqa: E402


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def timed(fn, iters=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()  # > L2 bytes written between iterations
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def k1():
    P = peaks()
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for (V, d, name) in [(128256, 2048, "C1 Llama-3.2-1B"), (128256, 4096, "C2 Llama-3.1-8B"), (152064, 3584, "C5 Qwen2.5-7B")]:
        table = (torch.randn(V, d, device=dev) * 0.02).bfloat16()
        bag = lr.B200EmbeddingBag.from_pretrained(table, padding_idx=V - 7)
        for Q in (1000, 10000, 100000):
            g = torch.Generator().manual_seed(Q)
            lens = torch.randint(1, 33, (Q,), generator=g)
            ids = torch.randint(0, V - 8, (int(lens.sum()),), generator=g).to(dev)
            offs = torch.cumsum(torch.cat([torch.zeros(1, dtype=torch.long), lens[:-1]]), 0).to(dev)
            for m in (None, 256):
                mm = m or d
                out = torch.empty((Q, mm), dtype=torch.bfloat16, device=dev)
                med, best = timed(lambda: bag.encode(ids, offs, shrink_dim=m, normalize=True, out=out, check_ids=False),
                                  flush=flush)
                nbytes = ids.numel() * mm * 2 + Q * mm * 2 + ids.numel() * 8 + Q * 8  # SURVEY §8d bytes/query x Q
                print(json.dumps({"kernel": "K1 embbag_encode", "table": name, "V": V, "d": d, "m": mm, "queries": Q,
                                  "tokens": ids.numel(), "ms": med, "ms_best": best,
                                  "roofline": {"bound": "hbm", "achieved": nbytes / med / 1e6, "peak": P["hbm_gbs"],
                                               "unit": "GB/s", "frac": nbytes / med / 1e6 / P["hbm_gbs"]},
                                  "qps": Q / med * 1e3}), flush=True)
        del table, bag


def k3():
    P = peaks()
    dev = "cuda"
    V, d, S = 128256, 4096, 512
    W = (torch.randn(V, d, device=dev) * 0.02).bfloat16()
    for B in (16, 64, 256):
        h = torch.randn(B, S, d, device=dev).bfloat16()
        lens = torch.randint(16, S + 1, (B,), device=dev)
        mask = (torch.arange(S, device=dev)[None] < lens[:, None])
        mask[:, 0] = False
        total = int(mask.sum())
        for packed in (True, False):
            med, best = timed(lambda: lr.max_linear_mapping(h, W, None, mask, relu=True, log1p=True, weight_is_vd=True,
                                                            packed=packed, valid_tokens=total), iters=5)
            flops = 2.0 * float(lens.sum()) * d * V  # 2 * sum_b S_b * d * V (document lengths, not the padded S)
            print(json.dumps({"kernel": "K3 sparse_head_max (umma_gemm_kernel<EPI_MAXTOK>)", "layout": "packed tokens" if packed else "padded [B,S]",
                              "B": B, "S": S, "d": d, "V": V, "tokens": float(lens.sum()), "ms": med, "ms_best": best, "docs_per_s": B / med * 1e3,
                              "roofline": {"bound": "tensor", "achieved": flops / med / 1e9, "peak": P["bf16_tflops_sustained"],
                                           "unit": "TFLOP/s", "frac": flops / med / 1e9 / P["bf16_tflops_sustained"],
                                           "algorithmic": "2*sum_b S_b*d*V; includes lr_pack_tokens for the packed layout"}}), flush=True)
        reps = lr.max_linear_mapping(h, W, None, mask, relu=True, log1p=True, weight_is_vd=True)
        med2, _ = timed(lambda: lr.sparsify_quantize(reps, top_k=256, min_tokens_to_keep=8), iters=5)
        print(json.dumps({"kernel": "K3 sparsify_quantize (select + scan + write)", "B": B, "V": V, "top_k": 256, "ms": med2}), flush=True)


def k4():
    P = peaks()
    dev = "cuda"
    V, nnz = 128256, 256
    for N in (1_100_000,):
        g = torch.Generator(device=dev).manual_seed(0)
        tok = torch.randint(0, V, (N, nnz), generator=g, device=dev, dtype=torch.int32)
        imp = torch.randint(1, 401, (N * nnz,), generator=g, device=dev, dtype=torch.int32)
        indptr = torch.arange(0, N * nnz + 1, nnz, dtype=torch.int64)
        idx = lr.ImpactIndex(V)
        idx.add_csr(indptr, tok.reshape(-1), imp)
        del tok, imp
        t0 = time.time()
        post = idx.build()
        torch.cuda.synchronize()
        build_s = time.time() - t0
        df = (post[0][1:] - post[0][:-1])
        for Q in (32, 1000, 10000):
            gq = torch.Generator().manual_seed(Q)
            lens = torch.randint(1, 33, (Q,), generator=gq)
            qt = torch.randint(0, V, (int(lens.sum()),), generator=gq, dtype=torch.int32)
            qc = torch.randint(1, 3, (int(lens.sum()),), generator=gq, dtype=torch.int32)
            qi = torch.cat([torch.zeros(1, dtype=torch.int32), torch.cumsum(lens, 0).to(torch.int32)])
            qi_d, qt_d, qc_d = qi.to(dev), qt.to(dev), qc.to(dev)
            for k in (100, 1000):
                med, best = timed(lambda: idx.search_device(qi_d, qt_d, qc_d, k), iters=5)
                nbytes = float(df[qt_d.long()].sum().item()) * 6  # sum_t df(t) * 6 B
                print(json.dumps({"kernel": "K4 sparse_score_topk", "N": N, "V": V, "nnz_per_doc": nnz, "queries": Q, "k": k,
                                  "ms": med, "ms_best": best, "qps": Q / med * 1e3, "index_build_s": build_s,
                                  "roofline": {"bound": "hbm", "achieved": nbytes / med / 1e6, "peak": P["hbm_gbs"], "unit": "GB/s",
                                               "frac": nbytes / med / 1e6 / P["hbm_gbs"],
                                               "note": "algorithmic bytes = postings only (accumulators live in shared memory)"}}),
                      flush=True)


def k2small():
    """BASELINE configs[4]: Qwen2.5-7B-shaped (d=3584), per-GPU shard of an 8.8M corpus over 8 GPUs, batch 1 / 32, top-100."""
    P = peaks()
    dev = "cuda"
    N, d, k = 1_100_000, 3584, 100
    c = torch.nn.functional.normalize(torch.randn(N, d, device=dev), dim=-1).bfloat16()
    for Q in (1, 32, 128):
        q = torch.nn.functional.normalize(torch.randn(Q, d, device=dev), dim=-1).bfloat16()
        med, best = timed(lambda: lr.flatip_topk(q, c, k), iters=20)
        nbytes = N * d * 2
        print(json.dumps({"kernel": "K2 flatip_topk small batch", "N": N, "d": d, "queries": Q, "k": k, "ms_p50": med,
                          "ms_best": best, "roofline": {"bound": "hbm", "achieved": nbytes / med / 1e6, "peak": P["hbm_gbs"],
                                                        "unit": "GB/s", "frac": nbytes / med / 1e6 / P["hbm_gbs"]}}), flush=True)


def mrl():
    """BASELINE configs[2]: MRL widths 128..1024 over full-width stored rows (prefix scoring + reciprocal prefix norms)."""
    P = peaks()
    dev = "cuda"
    N, d, Q, k = 1_100_000, 4096, 10000, 100
    c = torch.nn.functional.normalize(torch.randn(N, d, device=dev), dim=-1).bfloat16()
    q = torch.nn.functional.normalize(torch.randn(Q, d, device=dev), dim=-1).bfloat16()
    for m in (128, 256, 512, 1024, 4096):
        cs = (1.0 / c[:, :m].float().norm(dim=1)).contiguous()
        qs = (1.0 / q[:, :m].float().norm(dim=1)).contiguous()
        med, best = timed(lambda: lr.flatip_topk(q, c, k, d_used=m, q_scale=qs, c_scale=cs), iters=5)
        flops = 2.0 * Q * N * m
        print(json.dumps({"kernel": "K2 flatip_topk MRL prefix", "N": N, "d_stored": d, "m": m, "queries": Q, "k": k, "ms": med,
                          "qps": Q / med * 1e3, "roofline": {"bound": "tensor", "achieved": flops / med / 1e9,
                                                             "peak": P["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                                             "frac": flops / med / 1e9 / P["bf16_tflops_sustained"]}}), flush=True)


if __name__ == "__main__":
    todo = sys.argv[1:] or ["k1", "k2small", "mrl", "k3", "k4"]
    for name in todo:
        globals()[name]()
        torch.cuda.empty_cache()
