"""A step-by-step MODEL of the K4 row kernel's selection logic (csrc/sparse_rows.cuh), run on the CPU against the oracle.

The CUDA kernel is parity-tested on the GPU (tests/test_gpu_parity.py).  What those tests cannot vary is TIMING: whether
the unit of the previous split of a query had finished when the next one started (so its list was inherited), and how
stale the query's shared score floor was when a step read it.  The kernel claims its result does not depend on either.
This model restates the logic that claim rests on — candidates taken on the RUNNING sum, in-place replacement of a listed
document, exact cuts at a full list, units chained per query with the inherited list struck from the merge, a floor that
is only a lower bound of the final k-th score — and draws those timing decisions at random:

  * rows are retired term by term inside a step (a document's sum grows posting by posting);
  * a lane passes the cheap test when new_sum >= min_sc, the exact test when key(new_sum, doc) > thr;
  * a passing document whose previous sum was >= min_sc may be listed: replace its entry, else append;
  * a full list (cap entries) is cut to its k largest keys, thr = the k-th key, the floor is published;
  * unit (s, q) starts from the list / thr of unit (s - 1, q) with probability 1/2 ("it had finished"), and then that
    unit's list is dropped from the merge; the floor a step sees is any earlier published value (or 0), at random;
  * the result is the top-k of the union of the surviving lists (score desc, id asc).

It must equal oracle.impact_topk on every draw.
"""
import numpy as np

from oracle import oracle


def key_of(score: int, doc: int) -> int:
    return (int(score) << 32) | (0xFFFFFFFF - int(doc))


def cut(lst: list[int], k: int) -> int:
    lst.sort(reverse=True)
    del lst[k:]
    return lst[-1]  # kstar: every kept key is >= it


def model_search(query: dict[int, int], docs: list[dict[int, int]], k: int, cap: int, step_docs: int, splits: int,
                 rng: np.random.Generator):
    n_docs = len(docs)
    postings: dict[int, list[tuple[int, int]]] = {}
    for d, vec in enumerate(docs):
        for t, imp in vec.items():
            postings.setdefault(t, []).append((d, imp))  # ascending document ids per token
    n_steps = -(-n_docs // step_docs)
    splits = max(1, min(splits, n_steps))
    floor_history = [0]          # every value the query's floor has held: a step may see any earlier one
    units = []                   # per split: (list, thr) as written to the workspace; None when inherited by the next
    prev = None
    for s in range(splits):
        s0, s1 = (s * n_steps) // splits, ((s + 1) * n_steps) // splits
        lst: list[int] = []
        thr = 0xFFFFFFFF          # every score-0 key is <= this
        if prev is not None and rng.random() < 0.5:   # the previous unit had finished: inherit, strike it from the merge
            lst, thr = list(prev[0]), prev[1]
            units[-1] = None
        min_sc = max(1, thr >> 32)
        for step in range(s0, s1):
            d0, d1 = step * step_docs, min(n_docs, (step + 1) * step_docs)
            floor_seen = floor_history[int(rng.integers(0, len(floor_history)))]  # stale or fresh, never from the future
            min_sc = max(min_sc, floor_seen)                                     # the floor only moves between steps
            acc: dict[int, int] = {}
            for t, w in query.items():                                            # rows of one term, then the next term
                for d, imp in postings.get(t, []):
                    if not (d0 <= d < d1) or w <= 0:
                        continue
                    old = acc.get(d, 0)
                    new = old + w * imp
                    acc[d] = new
                    if new < min_sc:
                        continue
                    kk = key_of(new, d)
                    if not kk > thr:
                        continue
                    if old >= min_sc:                                             # may be listed: replace in place
                        low = 0xFFFFFFFF - d
                        hit = [i for i, e in enumerate(lst) if (e & 0xFFFFFFFF) == low]
                        if hit:
                            lst[hit[0]] = kk
                            continue
                    lst.append(kk)
                    if len(lst) == cap:
                        thr = cut(lst, k)
                        floor_history.append(max(floor_history[-1], thr >> 32))
                        min_sc = max(min_sc, thr >> 32)
        units.append((lst, thr))
        prev = units[-1]
    merged = sorted((e for u in units if u is not None for e in u[0]), reverse=True)
    assert len({e & 0xFFFFFFFF for e in merged}) == len(merged)                   # no document is listed twice
    merged = merged[:k]
    return [e >> 32 for e in merged], [0xFFFFFFFF - (e & 0xFFFFFFFF) for e in merged]


def test_row_kernel_selection_logic_is_exact_under_any_timing():
    rng = np.random.default_rng(2024)
    for trial in range(40):
        n_docs = int(rng.integers(50, 700))
        vocab = int(rng.integers(8, 60))
        k = int(rng.choice([1, 5, 20, 100]))
        cap = k + int(rng.integers(2, 40))
        step_docs = int(rng.choice([16, 64, 256]))
        splits = int(rng.integers(1, 9))
        docs = []
        for _ in range(n_docs):
            toks = rng.choice(vocab, size=int(rng.integers(0, min(vocab, 12))), replace=False)
            # few distinct impacts: ties at the threshold score are the hard case for (score desc, id asc)
            docs.append({int(t): int(rng.integers(1, 6 if trial % 2 else 300)) for t in toks})
        queries = [{int(t): int(rng.integers(1, 4)) for t in rng.choice(vocab, size=int(rng.integers(1, min(vocab, 10))), replace=False)}
                   for _ in range(4)]
        es, ei = oracle.impact_topk(queries, docs, k)
        for qi, q in enumerate(queries):
            for _ in range(3):  # several timing draws per query
                gs, gi = model_search(q, docs, k, cap, step_docs, splits, rng)
                n_hit = int((ei[qi] >= 0).sum())
                assert gi == ei[qi][:n_hit].tolist(), (trial, qi, gi[:5], ei[qi][:5])
                assert gs == es[qi][:n_hit].astype(np.int64).tolist()
