"""Per-kernel timeline of ONE bench step on rank 0 (every rank runs the step), from torch.profiler (CUPTI): which kernels run,
for how long, and the idle gaps between them — the N>1 counterpart of the single-GPU ncu launch list (ncu must not wrap
a multi-rank command).  Durations are CUPTI activity records, not ncu's serialised cold-cache replays.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/trace_step.py --config c2 --out profiles/x.json
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

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--out", default="gpurun_out/trace_step.json")
    ap.add_argument("--docs", type=int, default=None)
    a = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    args = argparse.Namespace(config=a.config, docs=a.docs, queries=None, dim=None, topk=None, parity_rows=8)
    wl = bench.make_workload(args)
    wl.setup(dev, rank, world)
    for w in range(4):
        wl.step(*wl.dev_batches[w % wl.n_batches])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for s in range(2):
            wl.step(*wl.dev_batches[s % wl.n_batches])
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if rank == 0:
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda e: e.time_range.start)
        half = len(evs) // 2
        step = evs[half:]  # the second of the two profiled steps
        t0 = step[0].time_range.start
        rows, busy_end, gaps = [], t0, 0.0
        for e in step:
            st, en = e.time_range.start, e.time_range.end
            if st > busy_end:
                gaps += st - busy_end
            busy_end = max(busy_end, en)
            rows.append({"kernel": e.name[:110], "start_us": round(st - t0, 1), "dur_us": round(en - st, 1)})
        total = busy_end - t0
        by = {}
        for r in rows:
            key = ("nccl" if "nccl" in r["kernel"].lower() else r["kernel"].split("<")[0].split("(")[0])
            by[key] = by.get(key, 0.0) + r["dur_us"]
        out = {"config": a.config, "n_gpus": world, "step_us": round(total, 1), "idle_gaps_us": round(gaps, 1),
               "by_kernel_us": {k: round(v, 1) for k, v in sorted(by.items(), key=lambda kv: -kv[1])}, "timeline": rows,
               "source": "torch.profiler (CUPTI) on rank 0, second of two profiled steps after 4 warm-ups"}
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps({k: out[k] for k in ("config", "n_gpus", "step_us", "idle_gaps_us", "by_kernel_us")}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
