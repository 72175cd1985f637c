"""One MRL configuration (compact [N, m] rows, 10k queries, 1.1M docs) for an ncu source-level capture of the epilogue-bound
regime: ncu --set full --import-source on -k regex:umma_gemm -s 5 -c 1 python tools/mrl_ncu.py  (launch 5 = main pass of call 3)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lightretriever_b200 as lr
m = int(os.environ.get("MRL_M", 128))
N, Q, k = 1_100_000, 10000, 100
c = torch.nn.functional.normalize(torch.randn(N, m, device="cuda"), dim=-1).bfloat16()
q = torch.nn.functional.normalize(torch.randn(Q, m, device="cuda"), dim=-1).bfloat16()
for _ in range(4):
    lr.flatip_topk(q, c, k)
torch.cuda.synchronize()
