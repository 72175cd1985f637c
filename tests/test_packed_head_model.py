"""A MODEL of the document bookkeeping of the packed K3 kernel (csrc/umma_gemm.cuh EPI_MAXTOK with cu_seqlens,
csrc/sparse_head.cu split_doc0_kernel / sparse_head_fix_kernel), run on the CPU over thousands of random layouts.

The GPU tests compare the CUDA kernel with the oracle on a handful of shapes.  The part of that kernel that is easy to get
wrong is not the GEMM but the INDEX LOGIC: tokens are cut into balanced runs of whole 256-token tiles, a document may
begin in one split and end in another (or cover several), empty documents can sit anywhere, and exactly one writer must
produce every document's value.  This file restates that logic statement by statement — same variables, same loop
structure (32-column chunks, `lim` / `ends_here`, the trailing `while`, head / tail pieces, the ownership rule of the
fix-up kernel) — with a scalar "score" per token instead of a vocabulary row, and checks, for random lengths and tile
sizes, that every document is written exactly once and holds the max over its own tokens.
"""
import numpy as np

LOWEST = -3.0e38


def split_range(split, splits, n_tiles, T, BN):
    c0 = (split * n_tiles) // splits * BN
    c1 = ((split + 1) * n_tiles) // splits * BN
    return c0, min(c1, T)


def split_doc0(cu, B, c0, split):
    lo, hi = 0, B - 1                      # smallest d with cu[d + 1] > c0
    while lo < hi:
        mid = (lo + hi) >> 1
        if cu[mid + 1] > c0:
            hi = mid
        else:
            lo = mid + 1
    return 0 if split == 0 else lo


def run_model(lens, scores, BN, tiles_per_split):
    B = len(lens)
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    T = int(cu[-1])
    assert T >= 1
    n_tiles = -(-T // BN)
    tps = max(tiles_per_split, 1)
    splits = max(1, -(-n_tiles // tps))
    out = [None] * B                       # value written for each document
    writes = [0] * B
    edge = np.full((splits, 2), LOWEST)
    doc0 = [split_doc0(cu, B, split_range(s, splits, n_tiles, T, BN)[0], s) for s in range(splits)]

    def write(seg, val):
        out[seg] = val
        writes[seg] += 1

    # ---- the epilogue of every unit (one "row" of the vocabulary)
    for split in range(splits):
        c0, c1 = split_range(split, splits, n_tiles, T, BN)
        seg = doc0[split]
        head_open = cu[seg] < c0
        seg_end = int(cu[seg + 1])
        next_end = int(cu[min(seg + 2, B)])
        run_max, head_piece = LOWEST, LOWEST

        def emit():
            nonlocal head_open, head_piece, run_max, seg
            if head_open:
                head_piece = run_max
                head_open = False
            else:
                write(seg, run_max)
            run_max = LOWEST
            seg += 1

        cb = c0
        while cb < c1:                                            # tiles
            n_valid = min(c1 - cb, BN)
            for c in range(BN // 32):                             # 32-column chunks
                if c * 32 >= n_valid:
                    break
                base = cb + c * 32
                col_lim = n_valid - c * 32
                ncols = min(col_lim, 32)
                m = [(base + j) < c1 for j in range(32)]          # mask == null: every packed token is valid
                while True:
                    lim = seg_end - base
                    ends_here = lim < ncols
                    part = [m[j] and (not ends_here or j < lim) for j in range(32)]
                    for j in range(32):
                        if part[j]:
                            run_max = max(run_max, float(scores[base + j]))
                    if not ends_here:
                        break
                    m = [m[j] and not part[j] for j in range(32)]
                    emit()
                    seg_end = next_end
                    next_end = int(cu[min(seg + 2, B)])
            cb += BN
        while seg < B and seg_end <= c1:
            emit()
            seg_end = next_end
            next_end = int(cu[min(seg + 2, B)])
        tail_piece = LOWEST
        if seg < B:
            if head_open:
                head_piece = run_max
            else:
                tail_piece = run_max
        edge[split, 0], edge[split, 1] = head_piece, tail_piece

    # ---- sparse_head_fix_kernel, one "CTA" per boundary
    for b in range(1, splits):
        c0, _ = split_range(b, splits, n_tiles, T, BN)
        p0, _ = split_range(b - 1, splits, n_tiles, T, BN)
        d = doc0[b]
        d_begin, d_end = int(cu[d]), int(cu[d + 1])
        if d_begin >= c0:
            continue
        if b > 1 and d_begin < p0:
            continue
        acc = edge[b - 1, 1]
        for s in range(b, splits):
            acc = max(acc, edge[s, 0])
            _, s1 = split_range(s, splits, n_tiles, T, BN)
            if d_end <= s1:
                break
        write(d, acc)
    return cu, out, writes


def test_packed_head_bookkeeping_writes_every_document_once_with_its_own_max():
    rng = np.random.default_rng(7)
    n_cases = 0
    for BN in (32, 64, 256):
        for trial in range(400 if BN < 256 else 120):
            B = int(rng.integers(1, 24))
            kind = trial % 4
            hi = [3 * BN, BN // 2 + 1, 40, 6 * BN][kind]
            lens = rng.integers(0, hi + 1, size=B)
            lens[rng.random(B) < 0.25] = 0                          # empty documents anywhere
            if kind == 1 and B > 2:                                 # documents that end exactly on tile boundaries
                lens[0] = BN
                lens[1] = 2 * BN
            if lens.sum() == 0:
                lens[int(rng.integers(0, B))] = int(rng.integers(1, hi + 1))
            T = int(lens.sum())
            scores = rng.standard_normal(T)
            tps = int(rng.choice([1, 1, 2, 3, 1000]))
            cu, out, writes = run_model(lens, scores, BN, tps)
            for d in range(B):
                assert writes[d] == 1, (BN, trial, d, writes[d], lens.tolist(), tps)
                want = float(scores[cu[d]:cu[d + 1]].max()) if lens[d] else LOWEST
                assert out[d] == want, (BN, trial, d, out[d], want, lens.tolist(), tps)
            n_cases += 1
    assert n_cases > 900
