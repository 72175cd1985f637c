"""K3 (sparse head) schedule sweep: band of vocabulary tiles, documents per unit, cluster mode, L2 policies (same box)."""
import itertools, json, os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lightretriever_b200 as lr

dev = "cuda"
V, d, S = 128256, 4096, 512
W = (torch.randn(V, d, device=dev) * 0.02).bfloat16()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
h = torch.randn(B, S, d, device=dev).bfloat16()
lens = torch.randint(16, S + 1, (B,), device=dev)
mask = (torch.arange(S, device=dev)[None] < lens[:, None])
flops = 2.0 * B * S * d * V


def run(env):
    for k in ("LR_SPARSE_HEAD_BAND", "LR_SPARSE_HEAD_DOCS_PER_UNIT", "LR_SPARSE_HEAD_CLUSTER", "LR_SPARSE_HEAD_POLICY_A", "LR_SPARSE_HEAD_POLICY_B"):
        os.environ.pop(k, None)
    os.environ.update(env)
    f = lambda: lr.max_linear_mapping(h, W, None, mask, relu=True, log1p=True, weight_is_vd=True)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = statistics.median(ts)
    print(json.dumps({"B": B, **env, "ms": round(ms, 3), "tflops": round(flops / ms / 1e9, 1)}), flush=True)


run({})
for band in (6, 12, 24, 48):
    for dpu in (1, 2, 4):
        run({"LR_SPARSE_HEAD_BAND": str(band), "LR_SPARSE_HEAD_DOCS_PER_UNIT": str(dpu)})
run({"LR_SPARSE_HEAD_CLUSTER": "3"})
run({"LR_SPARSE_HEAD_CLUSTER": "1"})
run({"LR_SPARSE_HEAD_POLICY_A": "2"})
run({"LR_SPARSE_HEAD_POLICY_B": "2"})
run({})
