"""K2 at the online shape (BASELINE configs[4]: 1.1M x 3584 shard, batch 32, top-100) for an ncu capture of the HBM-bound
regime: ncu --set full -k regex:umma_gemm -s 7 -c 1 python tools/k2small_ncu.py  (launches alternate prefix / main pass)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lightretriever_b200 as lr
N, d, k, Q = 1_100_000, 3584, 100, int(os.environ.get("K2_Q", 32))
c = torch.nn.functional.normalize(torch.randn(N, d, device="cuda"), dim=-1).bfloat16()
q = torch.nn.functional.normalize(torch.randn(Q, d, device="cuda"), dim=-1).bfloat16()
for _ in range(5):
    lr.flatip_topk(q, c, k)
torch.cuda.synchronize()
