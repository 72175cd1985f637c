"""First-contact GPU diagnostics: every kernel checked against the oracle in its own subprocess (a device trap in
one section cannot poison the others), printing error *patterns* (which rows / columns / k-steps are wrong) rather
than a bare pass/fail.  Usage: python tools/gpu_check.py [section ...]   (run under gpurun)."""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SECTIONS = ["embbag", "gemm_small", "gemm_shapes", "topk", "topk_edge", "merge", "sparse_score", "sparse_head", "big"]


def sec_embbag():
    import torch
    import lightretriever_b200 as lr
    from oracle import oracle
    torch.manual_seed(0)
    V, d = 5000, 2048
    table = (torch.randn(V, d) * 0.02).bfloat16()
    lens = torch.randint(0, 33, (257,))
    lens[3] = 0
    ids = torch.randint(0, V, (int(lens.sum()),))
    pad = 7
    ids[::11] = pad
    offsets = torch.cumsum(torch.cat([torch.zeros(1, dtype=torch.long), lens[:-1]]), 0)
    for m, norm, odt in [(None, False, torch.float32), (256, True, torch.float32), (None, True, torch.bfloat16)]:
        ref = oracle.embbag_encode(ids, offsets, table.float(), pad, m, norm)
        bag = lr.B200EmbeddingBag.from_pretrained(table.cuda(), padding_idx=pad)
        got = bag.encode(ids.cuda(), offsets.cuda(), shrink_dim=m, normalize=norm, out_dtype=odt).float().cpu()
        err = (got - ref).abs().max().item()
        print(f"embbag m={m} norm={norm} out={odt}: max abs err {err:.3e} (ref max {ref.abs().max():.3e})")


def _gemm_case(Q, N, d, seed=0):
    import torch
    import lightretriever_b200 as lr
    torch.manual_seed(seed)
    q = torch.randn(Q, d).bfloat16().cuda()
    c = torch.randn(N, d).bfloat16().cuda()
    got = lr.flatip_scores(q, c)
    torch.cuda.synchronize()
    ref = q.float() @ c.float().T
    diff = (got - ref).abs()
    tol = 1e-3 * (d ** 0.5) + 1e-2
    bad = diff > tol
    print(f"gemm Q={Q} N={N} d={d}: max err {diff.max().item():.4e} tol {tol:.3e} bad {int(bad.sum())}/{bad.numel()}")
    if bad.any():
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print("  bad rows (first 16):", rows[:16].tolist(), "count", rows.numel())
        print("  bad cols (first 16):", cols[:16].tolist(), "count", cols.numel())
        print("  got[0,:8]", got[0, :8].tolist())
        print("  ref[0,:8]", ref[0, :8].tolist())
        # is it a K-slice problem? compare with partial sums over 16-wide K slices
        if d <= 256:
            for ks in range(0, d, 16):
                part = q[:, :ks + 16].float() @ c[:, :ks + 16].float().T
                print(f"  vs prefix K<{ks + 16}: max err {(got - part).abs().max().item():.3e}")
    return not bool(bad.any())


def sec_gemm_small():
    _gemm_case(128, 256, 64)
    _gemm_case(128, 256, 128)
    _gemm_case(100, 300, 256)


def sec_gemm_shapes():
    _gemm_case(300, 1000, 4096)
    _gemm_case(129, 257, 72)
    _gemm_case(1, 5, 8)
    _gemm_case(1000, 70000, 512)


def sec_topk():
    import numpy as np
    import torch
    import lightretriever_b200 as lr
    from oracle import oracle
    for (Q, N, d, k) in [(64, 5000, 128, 10), (300, 20000, 256, 100), (130, 3000, 2048, 1000), (5, 100000, 64, 100)]:
        torch.manual_seed(1)
        q = torch.nn.functional.normalize(torch.randn(Q, d), dim=-1).bfloat16()
        c = torch.nn.functional.normalize(torch.randn(N, d), dim=-1).bfloat16()
        s, i = lr.flatip_topk(q.cuda(), c.cuda(), k)
        torch.cuda.synchronize()
        ref = (q.float() @ c.float().T).numpy()
        try:
            oracle.check_topk_parity(s.cpu().numpy(), i.cpu().numpy(), ref, k)
            ex_s, ex_i = oracle.flatip_topk(q.float(), c.float(), k)
            same = (ex_i == i.cpu().numpy()).mean()
            print(f"topk Q={Q} N={N} d={d} k={k}: PARITY OK, ids identical to oracle order: {same * 100:.2f}%")
        except AssertionError as e:
            print(f"topk Q={Q} N={N} d={d} k={k}: FAIL {str(e)[:400]}")


def sec_topk_edge():
    import numpy as np
    import torch
    import lightretriever_b200 as lr
    from oracle import oracle
    torch.manual_seed(2)
    # N < k, k = 1, duplicates (exact ties), zero query, id_offset, MRL prefix on strided views, scales
    q = torch.randn(7, 64).bfloat16()
    q[2] = 0
    c = torch.randn(50, 64).bfloat16()
    c[10:20] = c[5]  # exact ties
    for k in (1, 10, 100):
        s, i = lr.flatip_topk(q.cuda(), c.cuda(), k, id_offset=1000)
        es, ei = oracle.flatip_topk(q.float(), c.float(), k, id_offset=1000)
        ok_i = np.array_equal(ei, i.cpu().numpy())
        ok_s = np.allclose(es, s.cpu().numpy(), rtol=1e-3, atol=1e-4, equal_nan=True)
        print(f"edge k={k}: ids exact {ok_i} scores {ok_s}")
        if not ok_i:
            print("  got", i.cpu().numpy()[2][:12], "\n  exp", ei[2][:12])
    big = torch.randn(40, 512).bfloat16()
    cb = torch.randn(3000, 512).bfloat16()
    s, i = lr.flatip_topk(big.cuda()[:, :128], cb.cuda()[:, :128], 20)
    es, ei = oracle.flatip_topk(big[:, :128].float(), cb[:, :128].float(), 20)
    print("MRL strided prefix ids exact:", np.array_equal(ei, i.cpu().numpy()))
    qs = torch.rand(40) + 0.5
    cs = torch.rand(3000) + 0.5
    s, i = lr.flatip_topk(big.cuda(), cb.cuda(), 20, q_scale=qs.cuda(), c_scale=cs.cuda())
    ref = (big.float() @ cb.float().T) * qs[:, None] * cs[None, :]
    try:
        oracle.check_topk_parity(s.cpu().numpy(), i.cpu().numpy(), ref.numpy(), 20, rtol=1e-3)
        print("scaled topk parity OK")
    except AssertionError as e:
        print("scaled topk FAIL", str(e)[:300])


def sec_merge():
    import numpy as np
    import torch
    import lightretriever_b200 as lr
    from oracle import oracle
    rng = np.random.default_rng(0)
    L, Q, cap, k = 8, 33, 100, 100
    scores = rng.standard_normal((L, Q, cap)).astype(np.float32)
    scores[:, :, ::7] = 0.25  # ties
    ids = np.stack([rng.permutation(100000)[:L * cap].reshape(L, cap) for _ in range(Q)], 1).astype(np.int64)
    keys = lr.encode_keys(torch.from_numpy(scores).cuda(), torch.from_numpy(ids).cuda())
    exp_keys = oracle.encode_keys(scores, ids)
    print("encode_keys exact:", np.array_equal(keys.cpu().numpy().view(np.uint64), exp_keys))
    s, i = lr.topk_merge(keys, k)
    es, ei = oracle.merge_topk([scores[l] for l in range(L)], [ids[l] for l in range(L)], k)
    print("merge ids exact:", np.array_equal(ei, i.cpu().numpy()), "scores exact:", np.array_equal(es, s.cpu().numpy()))
    counts = torch.from_numpy(rng.integers(0, cap + 1, (L, Q)).astype(np.int32)).cuda()
    s, i = lr.topk_merge(keys, 37, counts=counts)
    cn = counts.cpu().numpy()
    es, ei = oracle.merge_topk([np.where(np.arange(cap)[None] < cn[l][:, None], scores[l], -np.inf) for l in range(L)],
                               [np.where(np.arange(cap)[None] < cn[l][:, None], ids[l], -1) for l in range(L)], 37)
    print("merge with counts ids exact:", np.array_equal(ei, i.cpu().numpy()))


def _rand_sparse_docs(rng, n, V, nnz, max_imp=400):
    docs = []
    for _ in range(n):
        m = int(rng.integers(0, nnz + 1))
        toks = rng.choice(V, size=m, replace=False)
        docs.append({str(int(t)): int(rng.integers(1, max_imp + 1)) for t in toks})
    return docs


def sec_sparse_score():
    import numpy as np
    import torch
    import lightretriever_b200 as lr
    from oracle import oracle
    rng = np.random.default_rng(3)
    for (N, V, nnz, Q, k) in [(3000, 500, 40, 20, 10), (40000, 2000, 64, 16, 100), (20000, 300, 30, 8, 1000)]:
        docs = _rand_sparse_docs(rng, N, V, nnz)
        queries = []
        for _ in range(Q):
            toks = rng.integers(0, V + 5, size=int(rng.integers(1, 33)))
            queries.append(" ".join(str(int(t)) for t in toks))
        searcher = lr.ImpactSearch(vocab_size=V)
        half = N // 2
        searcher.index(docs[:half], [f"d{j}" for j in range(half)])
        searcher.index(docs[half:], [f"d{j}" for j in range(half, N)])
        res = searcher.retrieve_with_emb(queries, [f"q{j}" for j in range(Q)], k)
        qd = [oracle.query_counts([int(t) for t in s.split()]) for s in queries]
        es, ei = oracle.impact_topk(qd, [{int(a): b for a, b in d.items()} for d in docs], k)
        bad = 0
        for r in range(Q):
            exp = {f"d{j}": float(s) for s, j in zip(es[r], ei[r]) if j >= 0}
            got = res.get(f"q{r}", {})
            if exp != got:
                bad += 1
                if bad == 1:
                    print("  first mismatch q", r, "exp n", len(exp), "got n", len(got),
                          "missing", list(set(exp) - set(got))[:5], "extra", list(set(got) - set(exp))[:5])
        print(f"sparse_score N={N} V={V} k={k}: {Q - bad}/{Q} queries bit-exact")


def sec_sparse_head():
    import numpy as np
    import torch
    import lightretriever_b200 as lr
    from oracle import oracle
    torch.manual_seed(4)
    for (B, S, d, V) in [(3, 64, 128, 1000), (5, 100, 256, 3001), (2, 512, 512, 5000)]:
        h = torch.randn(B, S, d).bfloat16()
        W = (torch.randn(V, d) * 0.05).bfloat16()
        bias = torch.randn(V) * 0.1
        lens = torch.randint(3, S + 1, (B,))
        am = (torch.arange(S)[None] < lens[:, None]).long()
        if B > 2:
            am[1] = 0
            am[1, :2] = 1  # no valid token after removing first/last
        mask = oracle.sparse_attention_mask(torch.zeros(B, S, dtype=torch.long), am, sep_token_id=-1)
        ref = oracle.max_linear_map(h.float(), W.float().T, bias, mask)
        got = lr.max_linear_mapping(h.cuda(), W.cuda(), bias.cuda(), mask.cuda(), weight_is_vd=True).cpu()
        err = (got - ref).abs()
        fin = ref > -1e30
        print(f"sparse_head B={B} S={S} d={d} V={V}: max err {err[fin].max().item():.3e}; empty-doc rows equal: "
              f"{bool((got[~fin] < -1e30).all())}")
        ref2 = oracle.get_sparse_emb(ref, True, True, top_k=64, min_tokens_to_keep=8)
        exp_json = oracle.quantize_reps(ref2, 100)
        indptr, tok, imp = lr.sparse_head(h.cuda(), W.cuda(), bias.cuda(), mask.cuda(), True, True, 64, 8, 100.0)
        got_json = lr.csr_to_json(indptr, tok, imp)
        nbad = 0
        for b in range(B):
            e, g = exp_json[b], got_json[b]
            keys = set(e) | set(g)
            for t in keys:
                if abs(e.get(t, 0) - g.get(t, 0)) > max(1, 0.01 * max(e.get(t, 0), g.get(t, 0))):
                    # membership may differ only inside the tie band at the top-k threshold
                    nbad += 1
        print(f"  sparsify: docs {B}, nnz got {[len(g) for g in got_json]} exp {[len(e) for e in exp_json]}, "
              f"entries outside the band: {nbad}")
        # quantiser alone is exact integer work: feed the oracle's own fp32 reps
        ip2, tk2, im2 = lr.sparsify_quantize(ref2.cuda(), top_k=0)
        print("  quantiser bit-exact:", lr.csr_to_json(ip2, tk2, im2) == exp_json)
        ip3, tk3, im3 = lr.sparsify_quantize(torch.log1p(torch.relu(ref)).cuda(), top_k=64, min_tokens_to_keep=8)
        print("  top-k + quantiser bit-exact:", lr.csr_to_json(ip3, tk3, im3) == exp_json)


def sec_big():
    import numpy as np
    import torch
    import lightretriever_b200 as lr
    torch.manual_seed(5)
    Q, N, d, k = 2048, 400_000, 4096, 100
    q = torch.nn.functional.normalize(torch.randn(Q, d, device="cuda"), dim=-1).bfloat16()
    c = torch.nn.functional.normalize(torch.randn(N, d, device="cuda"), dim=-1).bfloat16()
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        s, i = lr.flatip_topk(q, c, k)
        torch.cuda.synchronize()
        dt = time.time() - t0
        print(f"big Q={Q} N={N} d={d} k={k}: {dt * 1e3:.1f} ms  {2 * Q * N * d / dt / 1e12:.1f} TFLOP/s")
    ref = q[:64].float() @ c.float().T
    rs, ri = torch.topk(ref, k, dim=1)
    same = (ri == i[:64]).float().mean().item()
    print(f"  ids identical to torch.topk on 64 queries: {same * 100:.2f}%  max score err "
          f"{(rs - s[:64]).abs().max().item():.3e}")
    import ctypes
    plan = (ctypes.c_int64 * 8)()
    lr._C.load().lr_flatip_last_plan(plan)
    print("  plan m_tiles,n_tiles,splits,band,cap,grid,units,rounds =", list(plan))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--run":
        globals()["sec_" + sys.argv[2]]()
        return
    todo = sys.argv[1:] or SECTIONS
    for name in todo:
        print(f"===== {name}", flush=True)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", name], timeout=420,
                               capture_output=True, text=True, cwd=ROOT)
            print(r.stdout[-6000:])
            if r.returncode != 0:
                print(f"[exit {r.returncode}] stderr tail:\n{r.stderr[-3000:]}")
        except subprocess.TimeoutExpired as e:
            print("[timeout]", (e.stdout or b"")[-2000:] if isinstance(e.stdout, bytes) else e.stdout)
        print(f"----- {name} done in {time.time() - t0:.1f}s", flush=True)


if __name__ == "__main__":
    main()
