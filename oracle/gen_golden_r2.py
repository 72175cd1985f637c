"""gen_golden_r2.py — round-2 additions to tests/golden/, again produced by RUNNING THE REFERENCE (build container only):

  * top_p.npz          lightretriever.finetune.sparse_pooling.top_p_sampling (imported) on relu/log1p-shaped rows;
  * notebook_cells.npz the reference's own score definitions, exec'd from its notebooks:
        scripts/asymmetric_sparse_infer.ipynb  cell with `def compute_similarity` (:207-228)   -> K4 integer scores
        scripts/asymmetric_dense_infer.ipynb   cell `scores = query_embeddings @ corpus_embedding.T` (:231) -> K2 scores

    python oracle/gen_golden_r2.py         # needs /root/reference (read-only); writes tests/golden/
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
sys.dont_write_bytecode = True


def notebook_cell(path: str, needle: str) -> str:
    nb = json.load(open(path))
    for c in nb["cells"]:
        src = "".join(c["source"])
        if c["cell_type"] == "code" and needle in src:
            return src
    raise SystemExit(f"{path}: no code cell contains {needle!r}")


def main():
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not mounted; golden vectors can only be generated in the build container")
    sys.path.insert(0, os.path.join(REF, "src"))
    from lightretriever.finetune.sparse_pooling import top_p_sampling

    g = torch.Generator().manual_seed(4321)
    # rows shaped like get_sparse_emb's input to top-p: log1p(relu(logits)) — many exact zeros, a positive tail
    V = 3000
    logits = torch.randn(6, V, generator=g) * 2.0 - 1.5
    logits[1] = torch.randn(V, generator=g) * 4.0 - 9.0          # very sparse row
    logits[2, :] = -1.0                                            # all zeros after relu
    logits[3, :5] = torch.tensor([9.0, 8.0, 7.0, 6.0, 5.0])       # peaked row
    reps = torch.log1p(torch.relu(logits))
    reps[4, 100:140] = reps[4, 100]                                # ties among non-zero values
    reps[3, :5] = torch.tensor([30.0, 20.0, 10.0, 9.0, 8.0])       # one entry holds almost all the mass: the min_keep cap binds
    reps[5] = reps[5] * 4.0                                        # a wider spread of probabilities
    out = {}
    for tp, mk in [(0.9, 8), (0.5, 8), (0.3, 8), (0.1, 8), (0.05, 8), (0.05, 1), (0.01, 8), (0.01, 40), (1.0, 8), (0.0, 8)]:
        out[f"p{tp}_k{mk}"] = top_p_sampling(reps.clone(), tp, min_tokens_to_keep=mk).numpy()
    np.savez_compressed(os.path.join(OUT, "top_p.npz"), reps=reps.numpy(), **out)

    # ---- notebook cells
    rng = np.random.default_rng(99)
    Vt = 400
    query_embeddings = [{int(t): int(rng.integers(1, 4)) for t in rng.choice(Vt, size=int(rng.integers(1, 12)), replace=False)}
                        for _ in range(7)]
    corpus_embeddings = [{int(t): int(rng.integers(1, 300)) for t in rng.choice(Vt, size=int(rng.integers(0, 40)), replace=False)}
                         for _ in range(60)]
    ns = {"query_embeddings": query_embeddings, "corpus_embeddings": corpus_embeddings, "print": lambda *a, **k: None}
    exec(notebook_cell(os.path.join(REF, "scripts", "asymmetric_sparse_infer.ipynb"), "def compute_similarity"), ns)
    sparse_scores = np.asarray(ns["scores"], dtype=np.int64)

    qd = torch.nn.functional.normalize(torch.randn(5, 64, generator=g), dim=-1)
    cd = torch.nn.functional.normalize(torch.randn(200, 64, generator=g), dim=-1)
    ns = {"query_embeddings": qd, "corpus_embedding": cd, "print": lambda *a, **k: None}
    exec(notebook_cell(os.path.join(REF, "scripts", "asymmetric_dense_infer.ipynb"), "query_embeddings @ corpus_embedding.T"), ns)
    dense_scores = ns["scores"].numpy()
    np.savez_compressed(os.path.join(OUT, "notebook_cells.npz"), sparse_queries=json.dumps([{str(k): v for k, v in q.items()} for q in query_embeddings]),
             sparse_docs=json.dumps([{str(k): v for k, v in d.items()} for d in corpus_embeddings]), sparse_scores=sparse_scores,
             dense_q=qd.numpy(), dense_c=cd.numpy(), dense_scores=dense_scores)

    meta_path = os.path.join(OUT, "META.json")
    meta = json.load(open(meta_path))
    meta["round2"] = {"generator": "oracle/gen_golden_r2.py", "torch": torch.__version__,
                      "imported": ["lightretriever.finetune.sparse_pooling.top_p_sampling"],
                      "executed_notebook_cells": ["scripts/asymmetric_sparse_infer.ipynb: def compute_similarity + scores",
                                                  "scripts/asymmetric_dense_infer.ipynb: scores = query_embeddings @ corpus_embedding.T"]}
    json.dump(meta, open(meta_path, "w"), indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
