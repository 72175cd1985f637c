"""K3 (sparse head) same-box A/B: team schedule (pairs / multicast, window, band) vs the round-robin schedule."""
import json, os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lightretriever_b200 as lr

dev = "cuda"
V, d, S = 128256, 4096, 512
W = (torch.randn(V, d, device=dev) * 0.02).bfloat16()
KNOBS = ("LR_SPARSE_HEAD_BAND", "LR_SPARSE_HEAD_DOCS_PER_UNIT", "LR_SPARSE_HEAD_CLUSTER", "LR_SPARSE_HEAD_SCHED",
         "LR_SPARSE_HEAD_TEAM_BAND", "LR_SPARSE_HEAD_TEAM_WINDOW")


def run(B, h, mask, env, ref=None):
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(env)
    f = lambda: lr.max_linear_mapping(h, W, None, mask, relu=True, log1p=True, weight_is_vd=True)
    for _ in range(3):
        out = f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = statistics.median(ts)
    same = None if ref is None else bool(torch.equal(out, ref))
    print(json.dumps({"B": B, **env, "ms": round(ms, 3), "tflops": round(2.0 * B * S * d * V / ms / 1e9, 1), "equal_to_rr": same}), flush=True)
    return out


for B in (16, 64, 256):
    h = torch.randn(B, S, d, device=dev).bfloat16()
    lens = torch.randint(16, S + 1, (B,), device=dev)
    mask = (torch.arange(S, device=dev)[None] < lens[:, None])
    ref = run(B, h, mask, {"LR_SPARSE_HEAD_SCHED": "0"})
    run(B, h, mask, {}, ref)
    run(B, h, mask, {"LR_SPARSE_HEAD_CLUSTER": "2"}, ref)
    run(B, h, mask, {"LR_SPARSE_HEAD_TEAM_WINDOW": "2"}, ref)
    run(B, h, mask, {"LR_SPARSE_HEAD_TEAM_BAND": "12"}, ref)
    run(B, h, mask, {"LR_SPARSE_HEAD_DOCS_PER_UNIT": "2"}, ref)
    run(B, h, mask, {"LR_SPARSE_HEAD_SCHED": "0", "LR_SPARSE_HEAD_CLUSTER": "3"}, ref)
