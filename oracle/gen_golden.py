"""gen_golden.py — generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (build container only).

    python oracle/gen_golden.py            # needs /root/reference (read-only); writes tests/golden/

The GPU box has no /root/reference, so the vectors are committed as small fixtures.  Everything here imports the
reference's own modules from /root/reference/src (sys.path, no copy) and records inputs + the reference's outputs.
Missing third-party wheels are stubbed only so that the module *imports*; no stubbed function is ever called:
  * sparse_emb_util (Rust, absent)  -> stub module so finetune/sparse_converter_mixin.py imports; we call the pure
    torch method convert_sparse_reps_to_json_pt (:103-160) only.
  * transformers 5.5 moved PreTrainedTokenizerBase -> alias so finetune/nonctx_emb_utils.py imports (best effort).
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
sys.dont_write_bytecode = True


def main():
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not mounted; golden vectors can only be generated in the build container")
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(os.path.dirname(OUT)))  # tests/: the fake backbone shared with the tests
    os.makedirs(OUT, exist_ok=True)
    meta = {"generator": "oracle/gen_golden.py", "torch": torch.__version__, "imported": []}
    g = torch.Generator().manual_seed(1234)

    # ---------------------------------------------------------------- K1: torch.nn.EmbeddingBag (the reference's dependency)
    V, d = 97, 32
    table = torch.randn(V, d, generator=g)
    pad = 5
    lens = [3, 0, 1, 7, 2, 0, 4]
    ids = torch.randint(0, V, (sum(lens),), generator=g)
    ids[[1, 4, 9]] = pad
    ids[3:4] = pad  # bag 2 becomes all-padding
    offsets = torch.tensor(np.cumsum([0] + lens[:-1]))
    bag = torch.nn.EmbeddingBag.from_pretrained(table, padding_idx=pad)  # nonctx_emb_utils.py:310-312
    full = bag.forward(input=ids, offsets=offsets)                      # modeling_hybrid.py:474
    m16 = torch.nn.functional.normalize(full[..., :16], p=2, dim=-1)     # modeling_hybrid.py:487-490
    np.savez(os.path.join(OUT, "embbag.npz"), table=table.numpy(), ids=ids.numpy(), offsets=offsets.numpy(), pad=pad,
             out_full=full.numpy(), out_m16_norm=m16.numpy(),
             out_full_norm=torch.nn.functional.normalize(full, p=2, dim=-1).numpy())

    # ---------------------------------------------------------------- K1b: dense_pooling.pooling('lasttoken')
    from lightretriever.finetune.dense_pooling import pooling
    meta["imported"].append("lightretriever.finetune.dense_pooling.pooling")
    hid = torch.randn(4, 9, 16, generator=g)
    am_right = torch.tensor([[1] * 9, [1] * 4 + [0] * 5, [1] * 1 + [0] * 8, [1] * 7 + [0] * 2])
    am_left = torch.tensor([[1] * 9, [0] * 5 + [1] * 4, [0] * 8 + [1] * 1, [0] * 2 + [1] * 7])
    np.savez(os.path.join(OUT, "lasttoken.npz"), hidden=hid.numpy(), am_right=am_right.numpy(), am_left=am_left.numpy(),
             out_right=pooling(hid, attention_mask=am_right, pooling_strategy="lasttoken").numpy(),
             out_left=pooling(hid, attention_mask=am_left, pooling_strategy="lasttoken").numpy())

    # ---------------------------------------------------------------- K3: mask, max_linear_mapping, top_k_sampling
    from lightretriever.finetune.sparse_pooling import get_sparse_attention_mask, top_k_sampling
    from lightretriever.utils.max_linear_map import max_linear_mapping
    meta["imported"] += ["lightretriever.finetune.sparse_pooling.get_sparse_attention_mask",
                         "lightretriever.finetune.sparse_pooling.top_k_sampling",
                         "lightretriever.utils.max_linear_map.max_linear_mapping"]
    B, S, dh, Vv = 4, 12, 16, 50
    h = torch.randn(B, S, dh, generator=g)
    W = torch.randn(dh, Vv, generator=g)
    bias = torch.randn(Vv, generator=g)
    am = torch.tensor([[1] * 12, [1] * 6 + [0] * 6, [1] * 2 + [0] * 10, [1] * 9 + [0] * 3])
    input_ids = torch.randint(10, 40, (B, S), generator=g)
    input_ids[:, 3] = 7  # separator
    mask = get_sparse_attention_mask(input_ids, am, sep_token_id=7, remove_prompt=False)
    mask_rp = get_sparse_attention_mask(input_ids, am, sep_token_id=7, remove_prompt=True)
    with torch.no_grad():
        out_f32 = max_linear_mapping(h, W, bias, mask)
        out_f32_nobias = max_linear_mapping(h, W, None, mask_rp)
        out_bf16 = max_linear_mapping(h.bfloat16(), W.bfloat16(), bias.bfloat16(), mask)
    reps = torch.log1p(torch.relu(out_f32.clone()))
    topk5 = top_k_sampling(reps, 5, min_tokens_to_keep=8)
    topk20 = top_k_sampling(reps, 20, min_tokens_to_keep=8)
    tied = torch.tensor([[0.5, 0.25, 0.5, 0.0, 0.25, 0.5, 0.125, 0.0, 0.25, 1.0, 0.0, 0.25]])
    np.savez(os.path.join(OUT, "sparse_head.npz"), h=h.numpy(), W=W.numpy(), bias=bias.numpy(), am=am.numpy(),
             input_ids=input_ids.numpy(), mask=mask.numpy(), mask_rp=mask_rp.numpy(), out_f32=out_f32.numpy(),
             out_f32_nobias=out_f32_nobias.numpy(), out_bf16=out_bf16.float().numpy(), reps=reps.numpy(),
             topk5=topk5.numpy(), topk20=topk20.numpy(), tied=tied.numpy(),
             tied_top3=top_k_sampling(tied, 3, min_tokens_to_keep=1).numpy(),
             tied_top2_min8=top_k_sampling(tied, 2, min_tokens_to_keep=8).numpy(),
             tied_top0=top_k_sampling(tied, 0, min_tokens_to_keep=8).numpy())

    # ---------------------------------------------------------------- quantiser (torch twin of the Rust converter)
    sys.modules.setdefault("sparse_emb_util", types.SimpleNamespace(Converter=lambda *a, **k: None,
                                                                    ICUWordPreTokenizer=None))
    from lightretriever.finetune.sparse_converter_mixin import SparseConverterMixin
    meta["imported"].append("lightretriever.finetune.sparse_converter_mixin.SparseConverterMixin.convert_sparse_reps_to_json_pt")
    conv = SparseConverterMixin.__new__(SparseConverterMixin)
    conv.vocab_dict = None
    qin = torch.cat([topk20, torch.zeros(1, Vv),
                     torch.tensor([[0.005, 0.015, 0.025, 0.035, -1.0, 2.345, 0.0049999, 6.55] + [0.0] * (Vv - 8)])])
    qjson = conv.convert_sparse_reps_to_json_pt(qin, quantization_factor=100)
    np.savez(os.path.join(OUT, "quantize.npz"), reps=qin.numpy(), json=json.dumps(qjson))

    # ---------------------------------------------------------------- fusion + heap merge
    from lightretriever.retriever.score_fuse_utils import fuse_scores_linear, fuse_scores_rrf
    from lightretriever.retriever.hybrid_search import HybridSearch
    meta["imported"] += ["lightretriever.retriever.score_fuse_utils", "lightretriever.retriever.hybrid_search.HybridSearch._add_to_heap"]
    rng = np.random.default_rng(7)
    dense = {f"q{i}": {f"d{j}": float(rng.standard_normal()) for j in rng.choice(30, 8, replace=False)} for i in range(3)}
    sparse = {f"q{i}": {f"d{j}": float(rng.integers(1, 500)) for j in rng.choice(30, 8, replace=False)} for i in range(3)}
    hs = HybridSearch.__new__(HybridSearch)
    heaps = {}
    chunks = [{f"q{i}": {f"d{c * 10 + j}": float(np.round(rng.standard_normal(), 3)) for j in range(10)} for i in range(2)}
              for c in range(3)]
    for ch in chunks:
        hs._add_to_heap(ch, heaps, top_k=7, ignore_identical_ids=False)
    with open(os.path.join(OUT, "fusion.json"), "w") as f:
        json.dump({"dense": dense, "sparse": sparse,
                   "linear": fuse_scores_linear([dense, sparse], weights=[0.7, 0.3]),
                   "rrf": fuse_scores_rrf([dense, sparse]),
                   "chunks": chunks, "heap_top7": {q: sorted(v) for q, v in heaps.items()}}, f, indent=1, sort_keys=True)

    # ---------------------------------------------------------------- flatten ids (tokenize_nonctx_qry_emb_bag)
    try:
        import transformers
        import transformers.tokenization_utils as tu
        if not hasattr(tu, "PreTrainedTokenizerBase"):
            tu.PreTrainedTokenizerBase = transformers.PreTrainedTokenizerBase
        from lightretriever.finetune.nonctx_emb_utils import tokenize_nonctx_qry_emb_bag

        class FakeTok:
            def __call__(self, queries, max_length, truncation, add_special_tokens, return_attention_mask):
                assert truncation and not add_special_tokens and not return_attention_mask
                return {"input_ids": [[(ord(c) % 50) for c in q][:max_length] for q in queries]}

        qs = ["what is a llm", "b", "retrieval on blackwell gpus", "x y"]
        enc = tokenize_nonctx_qry_emb_bag(qs, FakeTok(), max_len=12)
        np.savez(os.path.join(OUT, "flatten.npz"), queries=np.array(qs), input_ids=enc["input_ids"].numpy(),
                 offsets=enc["offsets"].numpy())
        meta["imported"].append("lightretriever.finetune.nonctx_emb_utils.tokenize_nonctx_qry_emb_bag")
    except Exception as e:  # recorded, not fatal: the oracle restates those two lines
        meta["flatten_import_error"] = repr(e)[:300]

    # ---------------------------------------------------------------- table construction (construct_embedding_bag)
    try:
        from lightretriever.finetune.nonctx_emb_utils import construct_embedding_bag
        from tests_support_fake_backbone import FakeBackbone, FakeTokenizer  # tests/tests_support_fake_backbone.py
        for name, add_bos, prompt in (("embbag_table_bos_prompt", True, "query: "), ("embbag_table_plain", False, None)):
            tok, mdl = FakeTokenizer(n_vocab=53, add_bos=add_bos), FakeBackbone(n_vocab=53, hidden=8)
            ref_bag = construct_embedding_bag(mdl, tok, prompt=prompt, batch_size=20)
            np.savez(os.path.join(OUT, name + ".npz"), table=ref_bag.weight.detach().numpy(),
                     padding_idx=ref_bag.padding_idx, inputs_seen=np.concatenate(mdl.seen, 0))
        meta["imported"].append("lightretriever.finetune.nonctx_emb_utils.construct_embedding_bag")
    except Exception as e:
        meta["construct_import_error"] = repr(e)[:300]

    with open(os.path.join(OUT, "META.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
