"""The oracle is pinned against outputs of the reference itself (tests/golden/, made by oracle/gen_golden.py)."""
import json
import os

import numpy as np
import torch

from oracle import oracle


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_embbag_matches_torch_embeddingbag(golden_dir):
    g = _load(golden_dir, "embbag.npz")
    pad = int(g["pad"])
    full = oracle.embbag_encode(g["ids"], g["offsets"], g["table"], pad)
    np.testing.assert_array_equal(full.numpy(), g["out_full"])
    # empty bag (index 1, 5) and all-padding bag (index 2) -> zero vectors
    assert not full[1].any() and not full[5].any() and not full[2].any()
    np.testing.assert_array_equal(oracle.embbag_encode(g["ids"], g["offsets"], g["table"], pad, 16, True).numpy(),
                                  g["out_m16_norm"])
    np.testing.assert_array_equal(oracle.embbag_encode(g["ids"], g["offsets"], g["table"], pad, None, True).numpy(),
                                  g["out_full_norm"])


def test_flatten_matches_reference_tokenizer_wrapper(golden_dir):
    g = _load(golden_dir, "flatten.npz")
    lists = [[(ord(c) % 50) for c in str(q)][:12] for q in g["queries"]]
    ids, offsets = oracle.flatten_token_ids(lists)
    np.testing.assert_array_equal(ids, g["input_ids"])
    np.testing.assert_array_equal(offsets, g["offsets"])


def test_lasttoken_matches_reference_pooling(golden_dir):
    g = _load(golden_dir, "lasttoken.npz")
    np.testing.assert_array_equal(oracle.lasttoken_head(g["hidden"], g["am_right"]).numpy(), g["out_right"])
    np.testing.assert_array_equal(oracle.lasttoken_head(g["hidden"], g["am_left"]).numpy(), g["out_left"])


def test_sparse_mask_and_max_linear_map_match_reference(golden_dir):
    g = _load(golden_dir, "sparse_head.npz")
    m = oracle.sparse_attention_mask(g["input_ids"], g["am"], sep_token_id=7, remove_prompt=False)
    np.testing.assert_array_equal(m.numpy(), g["mask"])
    m_rp = oracle.sparse_attention_mask(g["input_ids"], g["am"], sep_token_id=7, remove_prompt=True)
    np.testing.assert_array_equal(m_rp.numpy(), g["mask_rp"])
    fmin32 = torch.finfo(torch.float32).min
    out = oracle.max_linear_map(g["h"], g["W"], g["bias"], g["mask"], fill=fmin32)
    np.testing.assert_allclose(out.numpy(), g["out_f32"], rtol=1e-6, atol=1e-6)
    out_nb = oracle.max_linear_map(g["h"], g["W"], None, g["mask_rp"], fill=fmin32)
    np.testing.assert_allclose(out_nb.numpy(), g["out_f32_nobias"], rtol=1e-6, atol=1e-6)
    # row 2 has no valid token after dropping first/last: stays at finfo.min (-> 0 after relu)
    assert (g["out_f32"][2] == fmin32).all()
    # the reference's own bf16 run agrees with the fp32 oracle inside its documented bf16 band (max_linear_map.py:192-196)
    fin = g["out_f32"] > -1e30
    np.testing.assert_allclose(g["out_bf16"][fin], g["out_f32"][fin], rtol=5e-2, atol=5e-2)


def test_topk_sampling_matches_reference(golden_dir):
    g = _load(golden_dir, "sparse_head.npz")
    reps = torch.from_numpy(g["reps"])
    np.testing.assert_array_equal(oracle.top_k_sampling(reps, 5, min_tokens_to_keep=8).numpy(), g["topk5"])
    np.testing.assert_array_equal(oracle.top_k_sampling(reps, 20, min_tokens_to_keep=8).numpy(), g["topk20"])
    tied = torch.from_numpy(g["tied"])
    np.testing.assert_array_equal(oracle.top_k_sampling(tied, 3, min_tokens_to_keep=1).numpy(), g["tied_top3"])
    np.testing.assert_array_equal(oracle.top_k_sampling(tied, 2, min_tokens_to_keep=8).numpy(), g["tied_top2_min8"])
    np.testing.assert_array_equal(oracle.top_k_sampling(tied, 0, min_tokens_to_keep=8).numpy(), g["tied_top0"])
    np.testing.assert_array_equal(oracle.get_sparse_emb(g["out_f32"], True, True, 20, 8).numpy(), g["topk20"])


def test_top_p_sampling_matches_reference(golden_dir):
    """oracle.top_p_sampling vs the imported sparse_pooling.top_p_sampling (oracle/gen_golden_r2.py), bit for bit."""
    g = _load(golden_dir, "top_p.npz")
    reps = torch.from_numpy(g["reps"])
    cases = [k for k in g.files if k != "reps"]
    assert len(cases) >= 8
    for k in cases:
        tp, mk = float(k.split("_")[0][1:]), int(k.split("_k")[1])
        np.testing.assert_array_equal(oracle.top_p_sampling(reps.clone(), tp, min_tokens_to_keep=mk)[0].numpy(), g[k])


def test_score_definitions_match_the_reference_notebook_cells(golden_dir):
    """The K4 and K2 score definitions are the reference's own notebook cells, exec'd by oracle/gen_golden_r2.py:
    compute_similarity (asymmetric_sparse_infer.ipynb:207-228) and `query_embeddings @ corpus_embedding.T`
    (asymmetric_dense_infer.ipynb:231)."""
    g = _load(golden_dir, "notebook_cells.npz")
    queries = [{int(k): v for k, v in q.items()} for q in json.loads(str(g["sparse_queries"]))]
    docs = [{int(k): v for k, v in d.items()} for d in json.loads(str(g["sparse_docs"]))]
    np.testing.assert_array_equal(oracle.impact_scores(queries, docs), g["sparse_scores"])
    s, i = oracle.impact_topk(queries, docs, 10)
    for r in range(len(queries)):
        hit = i[r] >= 0
        np.testing.assert_array_equal(s[r][hit], g["sparse_scores"][r][i[r][hit]].astype(np.float32))
    dense = oracle.flatip_scores(torch.from_numpy(g["dense_q"]), torch.from_numpy(g["dense_c"])).numpy()
    np.testing.assert_array_equal(dense, g["dense_scores"])


def test_quantiser_matches_reference_torch_twin(golden_dir):
    g = _load(golden_dir, "quantize.npz")
    assert oracle.quantize_reps(g["reps"], 100) == json.loads(str(g["json"]))


def test_fusion_and_heap_match_reference(golden_dir):
    with open(os.path.join(golden_dir, "fusion.json")) as f:
        g = json.load(f)
    lin = oracle.fuse_linear([g["dense"], g["sparse"]], weights=[0.7, 0.3])
    rrf = oracle.fuse_rrf([g["dense"], g["sparse"]])
    for q in g["linear"]:
        assert lin[q].keys() == g["linear"][q].keys()
        for p in lin[q]:
            assert abs(lin[q][p] - g["linear"][q][p]) < 1e-12
            assert abs(rrf[q][p] - g["rrf"][q][p]) < 1e-12
    heaps = {}
    for ch in g["chunks"]:
        oracle.add_to_heap(ch, heaps, 7)
    assert {q: sorted([s, p] for s, p in v) for q, v in heaps.items()} == g["heap_top7"]


def test_merge_of_chunk_topk_equals_full_topk():
    rng = np.random.default_rng(0)
    q = rng.standard_normal((9, 24)).astype(np.float32)
    c = rng.standard_normal((500, 24)).astype(np.float32)
    c[100:110] = c[7]
    full_s, full_i = oracle.flatip_topk(q, c, 40)
    parts = [oracle.flatip_topk(q, c[lo:lo + 130], 40, id_offset=lo) for lo in range(0, 500, 130)]
    ms, mi = oracle.merge_topk([p[0] for p in parts], [p[1] for p in parts], 40)
    np.testing.assert_array_equal(mi, full_i)
    np.testing.assert_array_equal(ms, full_s)
    # key round trip preserves (score desc, id asc)
    keys = oracle.encode_keys(full_s, full_i)
    assert (np.diff(keys.astype(np.float64), axis=1) <= 0).all()
    ds, di = oracle.decode_keys(keys)
    np.testing.assert_array_equal(ds, full_s)
    np.testing.assert_array_equal(di, full_i)
    fs, fi = oracle.flatip_topk_fast(q, c, 40)
    np.testing.assert_allclose(fs.numpy(), full_s, rtol=1e-6)


def test_impact_formula_matches_notebook_definition():
    # compute_similarity of scripts/asymmetric_sparse_infer.ipynb:207-228: sum over shared tokens of q[t]*d[t]
    docs = [{1: 300, 2: 50}, {2: 10}, {}, {5: 700, 1: 1}]
    qs = [{1: 2, 2: 1}, {9: 1}, {5: 1, 1: 1}]
    S = oracle.impact_scores(qs, docs)
    assert S.tolist() == [[650, 10, 0, 2], [0, 0, 0, 0], [300, 0, 0, 701]]
    s, i = oracle.impact_topk(qs, docs, 3)
    assert i.tolist() == [[0, 1, 3], [-1, -1, -1], [3, 0, -1]]
    assert oracle.query_counts([4, 4, 9], "sum") == {4: 2, 9: 1} and oracle.query_counts([4, 4, 9], "bow") == {4: 1, 9: 1}


def test_embedding_bag_table_construction_matches_reference(golden_dir):
    """Row f4: ``construct_embedding_bag`` (finetune/nonctx_emb_utils.py:239-313) — the golden tables were produced by the
    reference's own function driving the fake backbone of tests_support_fake_backbone.py; ours must feed the backbone the
    same input ids in the same batches and return the same table and padding_idx (host logic: runs without a GPU)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from tests_support_fake_backbone import FakeBackbone, FakeTokenizer
    import lightretriever_b200 as lr
    for name, add_bos, prompt in (("embbag_table_bos_prompt", True, "query: "), ("embbag_table_plain", False, None)):
        g = _load(golden_dir, name + ".npz")
        tok, mdl = FakeTokenizer(n_vocab=53, add_bos=add_bos), FakeBackbone(n_vocab=53, hidden=8)
        bag = lr.construct_embedding_bag(mdl, tok, prompt=prompt, batch_size=20)
        np.testing.assert_array_equal(np.concatenate(mdl.seen, 0), g["inputs_seen"])
        np.testing.assert_array_equal(bag.weight.numpy(), g["table"])
        assert bag.padding_idx == int(g["padding_idx"]) and bag.weight.dtype == torch.float32
        # the oracle's plain-loop restatement of the input layout
        prompt_ids = ([tok.bos_token_id] if add_bos else []) + (tok.encode(prompt, add_special_tokens=False) if prompt else [])
        np.testing.assert_array_equal(oracle.emb_bag_table_inputs(prompt_ids, tok.eos_token_id, 0, 53), g["inputs_seen"])
        np.testing.assert_array_equal(lr.emb_bag_inputs(prompt_ids, tok.eos_token_id, 20, 40).numpy(), g["inputs_seen"][20:40])
        b16 = lr.construct_embedding_bag(FakeBackbone(n_vocab=53, hidden=8), tok, prompt=prompt, batch_size=53,
                                         table_dtype=torch.bfloat16)
        np.testing.assert_array_equal(b16.weight.float().numpy(), torch.from_numpy(g["table"]).bfloat16().float().numpy())
