#!/usr/bin/env python
"""bench.py — headline benchmark of the serving-side retrieval hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1] at the metric's k=100): Llama-3.1-8B-shaped EmbeddingBag table (V=128256, d=4096,
bf16, random init), batches of 10 000 synthetic queries (1..32 tokens), exact inner-product top-100 over an
8.8M x 4096 bf16 synthetic corpus (72.1 GB), row-sharded over the N GPUs (strong scaling: the corpus is fixed).
One step = one query batch: K1 EmbeddingBag encode (+L2 norm) -> K2 fused tcgen05 scoring/top-k -> merge
(-> NCCL all-gather of the per-shard top-k keys -> merge, when N > 1).

Prints ONE JSON line (rank 0).  `value` = QPS with the step's inputs resident in HBM; `e2e` = QPS through the public
API with HOST buffers (pinned ids/offsets in, scores/ids out, copies inside the timed region).
`--impl reference` times the reference's CPU path (torch.nn.EmbeddingBag fp32 + torch.matmul + torch.topk — faiss is
not installed) on the host cores, on a bounded sample of the same workload, extrapolated linearly in N.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "QPS (EmbBag encode + exact top-100, 8.8M x 4096 bf16)"
VOCAB, DIM, N_DOCS, Q_BATCH, TOPK, MAX_TOK = 128256, 4096, 8_800_000, 10_000, 100, 32
PAD_ID = 128002
CHUNK_ROWS = 131072


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    # development overrides (the judged run uses the defaults)
    ap.add_argument("--docs", type=int, default=N_DOCS)
    ap.add_argument("--queries", type=int, default=Q_BATCH)
    ap.add_argument("--dim", type=int, default=DIM)
    ap.add_argument("--topk", type=int, default=TOPK)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    return ap.parse_args()


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def load_traffic(args, world):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this exact config."""
    try:
        with open(os.path.join(ROOT, "profiles", "k2_traffic.json")) as f:
            t = json.load(f)
        c = t["config"]
        if (c["docs"], c["queries"], c["dim"], c["k"], c["n_gpus"]) == (args.docs, args.queries, args.dim, args.topk, world):
            return t["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def make_queries(n_queries: int, seed: int):
    import torch
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(1, MAX_TOK + 1, (n_queries,), generator=g)
    ids = torch.randint(0, VOCAB, (int(lens.sum()),), generator=g)
    ids[ids == PAD_ID] = 0
    offsets = torch.cumsum(torch.cat([torch.zeros(1, dtype=torch.long), lens[:-1]]), 0)
    return ids, offsets


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.3] or [r for _, r in self.rows]
        for r in rows:
            parts = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
                for nme, val in zip(names, parts[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ reference arm / CPU baseline
def cpu_reference_sample(args, seconds: float, steps: int = 1, warmup: int = 0):
    """The reference's CPU path on a bounded sample: fp32 EmbeddingBag + normalize, fp32 matmul + topk over a row
    sub-sample of the corpus (same d / k / query shape).  Returns QPS extrapolated linearly to the full corpus."""
    import torch
    from oracle import oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    d, k = args.dim, args.topk
    g = torch.Generator().manual_seed(0)
    table = torch.randn(VOCAB, d, generator=g) * 0.02  # fp32 table, as the reference library path keeps it
    qn = min(args.queries, 1000)
    ids, offsets = make_queries(qn, seed=1)
    # calibrate the matmul rate, then size the row sample for ~`seconds` of work
    probe = torch.randn(4096, d)
    qprobe = torch.randn(qn, d)
    t = time.time()
    (qprobe @ probe.T).topk(min(k, 4096), dim=1)
    rate = 4096 / max(time.time() - t, 1e-4)  # corpus rows per second at this query batch
    n_s = int(min(args.docs, max(20_000, min(400_000, rate * seconds))))
    corpus = torch.nn.functional.normalize(torch.randn(n_s, d, generator=g), dim=-1)
    times = []
    for it in range(warmup + steps):
        t0 = time.time()
        qv = oracle.embbag_encode(ids, offsets, table, PAD_ID, None, True)
        oracle.flatip_topk_fast(qv, corpus, k)
        times.append(time.time() - t0)
    t_step = statistics.mean(times[warmup:])
    qps_sample = qn / t_step
    qps_full = qps_sample * n_s / args.docs
    return {"value": qps_full, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": (f"{qn} queries x {n_s} of {args.docs} docs, d={d}, k={k}, fp32 torch EmbeddingBag+matmul+topk "
                       f"(faiss absent); measured {qps_sample:.1f} q/s on the sample, scaled by {n_s}/{args.docs}"),
            "ms_per_step_sample": t_step * 1e3, "threads": torch.get_num_threads()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference_sample(args, seconds=max(5.0, args.cpu_seconds), steps=max(args.steps, 1), warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step_sample"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, max(int(args.gpus), 1)),  # the B200 arm's config for the same N
        "cpu_baseline": {k_: cb[k_] for k_ in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": "C2@k=%d: EmbeddingBag(V=128256,d=%d,bf16) encode + exact IP top-%d, %d-query batches vs %d-doc "
                        "bf16 corpus" % (args.topk, args.dim, args.topk, args.queries, args.docs),
            "queries_per_step": args.queries, "docs": args.docs, "dim": args.dim, "k": args.topk,
            "max_query_tokens": MAX_TOK, "sharding": f"corpus row-sharded over {world} GPU(s)",
            "l2": "inputs larger than L2 (corpus shard streams from HBM every step)"}


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import lightretriever_b200 as lr
    from lightretriever_b200.sharded import exchange_candidates, shard_range
    lib = lr._C.load()
    d, k, Q = args.dim, args.topk, args.queries

    # ---- synthetic state: table, corpus shard (generated on device, chunk-seeded so every N sees the same corpus)
    gt = torch.Generator(device=dev).manual_seed(0)
    table = (torch.randn(VOCAB, d, generator=gt, device=dev) * 0.02).bfloat16()
    bag = lr.B200EmbeddingBag.from_pretrained(table, padding_idx=PAD_ID)
    lo, hi = shard_range(args.docs, rank, world)
    corpus = torch.empty((hi - lo, d), dtype=torch.bfloat16, device=dev)
    for c0 in range((lo // CHUNK_ROWS) * CHUNK_ROWS, hi, CHUNK_ROWS):
        gc = torch.Generator(device=dev).manual_seed(1000 + c0 // CHUNK_ROWS)
        blk = torch.nn.functional.normalize(torch.randn(CHUNK_ROWS, d, generator=gc, device=dev), dim=-1).bfloat16()
        a, b = max(c0, lo), min(c0 + CHUNK_ROWS, hi)
        corpus[a - lo:b - lo] = blk[a - c0:b - c0]
        del blk
    n_batches = 1 if os.environ.get('LR_BENCH_SAME_BATCH') else 4
    host_batches = []
    for i in range(n_batches):
        ids, offs = make_queries(Q, seed=100 + i)
        host_batches.append((ids.pin_memory(), offs.pin_memory()))
    dev_batches = [(a.to(dev), b.to(dev)) for a, b in host_batches]
    out_s_host = torch.empty((Q, k), dtype=torch.float32).pin_memory()
    out_i_host = torch.empty((Q, k), dtype=torch.int64).pin_memory()

    def search_step(ids, offs):
        qv = bag.encode(ids, offs, normalize=True, check_ids=False)               # K1
        if world == 1:
            return lr.flatip_topk(qv, corpus, k, id_offset=lo)                      # K2 + merge
        _, _, keys = lr.flatip_topk(qv, corpus, k, id_offset=lo, return_keys=True)  # K2 + merge (per shard)
        return lr.topk_merge(exchange_candidates(keys), k)                          # all-gather + merge

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # profile events around the main kernel of every timed step
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in kev:
        a.record()
        b.record()  # materialise the cudaEvent_t handles
    torch.cuda.synchronize()

    # ---- warm-up
    for w in range(max(args.warmup, 3)):
        search_step(*dev_batches[w % n_batches])
    barrier()

    # ---- timed region 1: inputs resident in HBM
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    t_wall0 = time.time()
    sev[0].record()
    for s in range(args.steps):
        lib.lr_set_profile_events(ctypes.c_void_p(kev[s][0].cuda_event), ctypes.c_void_p(kev[s][1].cuda_event))
        res = search_step(*dev_batches[s % n_batches])
        sev[s + 1].record()
    lib.lr_set_profile_events(None, None)
    plan = (ctypes.c_int64 * 8)()
    lib.lr_flatip_last_plan(plan)
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    step_ms = [sev[s].elapsed_time(sev[s + 1]) for s in range(args.steps)]
    total_ms = sev[0].elapsed_time(sev[args.steps])
    kern_ms = [a.elapsed_time(b) for a, b in kev]

    # ---- timed region 2: end to end through the public API with host buffers
    for w in range(2):
        ids, offs = host_batches[w % n_batches]
        s_, i_ = search_step(ids.to(dev, non_blocking=True), offs.to(dev, non_blocking=True))
        out_s_host.copy_(s_, non_blocking=True)
        out_i_host.copy_(i_, non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        ids, offs = host_batches[s % n_batches]
        s_, i_ = search_step(ids.to(dev, non_blocking=True), offs.to(dev, non_blocking=True))
        out_s_host.copy_(s_, non_blocking=True)
        out_i_host.copy_(i_, non_blocking=True)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    h2d = int(statistics.mean(a.numel() * 8 + b.numel() * 8 for a, b in host_batches))
    d2h = Q * k * (4 + 8)

    # ---- max over ranks
    t = torch.tensor([total_ms, e2e_ms, statistics.mean(kern_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kern_ms_mean = t.tolist()

    # ---- parity spot check outside the timed region (device fp32 matmul on a few queries, local shard)
    ids, offs = dev_batches[(args.steps - 1) % n_batches]
    qv = bag.encode(ids, offs, normalize=True, check_ids=False)
    nchk = 8
    ls, li = lr.flatip_topk(qv[:nchk], corpus, k, id_offset=lo)
    best = None
    for c0 in range(0, corpus.shape[0], 1 << 18):
        sc = qv[:nchk].float() @ corpus[c0:c0 + (1 << 18)].float().T
        ts, ti = sc.topk(min(k, sc.shape[1]), dim=1)
        ti = ti + c0 + lo
        if best is None:
            best = (ts, ti)
        else:
            cs, ci = torch.cat([best[0], ts], 1), torch.cat([best[1], ti], 1)
            ms_, mi_ = cs.topk(k, dim=1)
            best = (ms_, torch.gather(ci, 1, mi_))
    ids_same = float((best[1] == li).float().mean())
    score_err = float((best[0] - ls).abs().max())

    if rank == 0:
        peaks = load_peaks()
        n_local = hi - lo
        flops = 2.0 * Q * n_local * d
        ach = flops / (kern_ms_mean * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        line = {
            "metric": METRIC, "value": Q * args.steps / (total_ms * 1e-3), "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, world),
            "p50_ms_per_batch": statistics.median(step_ms),
            "e2e": {"value": Q * args.steps / (e2e_ms * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": args.steps * ((3 if world == 1 else 4) + (3 if plan[7] > 0 else 0)),
            "roofline": {"kernel": "umma_gemm_kernel<EPI_TOPK> (tcgen05 bf16 GEMM + fused top-k epilogue)",
                         "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                         "peak_source": f"{peaks['source']} MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)",
                         "flop_per_launch": flops, "kernel_ms": kern_ms_mean, "traffic": load_traffic(args, world),
                         "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read+write, profiles/k2_traffic.json)",
                         "kernel_share_of_step": kern_ms_mean / (total_ms / args.steps)},
            "plan": dict(zip(["m_tiles", "n_tiles", "splits", "band", "cap", "grid", "units", "prefix_tiles"], list(plan))),
            "parity_spot_check": {"queries": nchk, "ids_identical_to_torch_fp32_topk": ids_same, "max_abs_score_err": score_err},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            cb = cpu_reference_sample(args, seconds=args.cpu_seconds)
            line["cpu_baseline"] = {k_: cb[k_] for k_ in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
