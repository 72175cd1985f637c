"""K4 probe: one configuration (10k queries, 1.1M docs x 256 nnz, k=100), a few launches — for ncu source-level profiles."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lightretriever_b200 as lr
dev = "cuda"
V, nnz, N, Q, k = 128256, 256, 1_100_000, int(os.environ.get("K4_Q", 10000)), 100
g = torch.Generator(device=dev).manual_seed(0)
tok = torch.randint(0, V, (N, nnz), generator=g, device=dev, dtype=torch.int32)
imp = torch.randint(1, 401, (N * nnz,), generator=g, device=dev, dtype=torch.int32)
idx = lr.ImpactIndex(V)
idx.add_csr(torch.arange(0, N * nnz + 1, nnz, dtype=torch.int64), tok.reshape(-1), imp)
del tok, imp
idx.build()
gq = torch.Generator().manual_seed(Q)
lens = torch.randint(1, 33, (Q,), generator=gq)
qt = torch.randint(0, V, (int(lens.sum()),), generator=gq, dtype=torch.int32).to(dev)
qc = torch.randint(1, 3, (int(lens.sum()),), generator=gq, dtype=torch.int32).to(dev)
qi = torch.cat([torch.zeros(1, dtype=torch.int32), torch.cumsum(lens, 0).to(torch.int32)]).to(dev)
for _ in range(3):
    idx.search_device(qi, qt, qc, k)
torch.cuda.synchronize()
