"""Query encoder (K1) and document dense head (K1b) — host side.

Mirrors the reference objects at seam S1:
  * ``model.emb_bag`` — ``torch.nn.EmbeddingBag.from_pretrained(table, padding_idx=pad)`` (mode='mean'), built by
    ``construct_embedding_bag`` (reference finetune/nonctx_emb_utils.py:239-313), held by ``EmbeddingBagMixin``
    (finetune/emb_bag_mixin.py:14-39) and called as ``emb_bag.forward(input=ids, offsets=offsets)``
    (finetune/modeling_hybrid.py:474), followed by ``[..., :dense_shrink_dim]`` and ``F.normalize`` (:487-490).
  * ``pooling(last_hidden, attention_mask=..., pooling_strategy='lasttoken')`` (finetune/dense_pooling.py:48-55).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import _C
from ._util import dtype_tag, require_cuda, stream_ptr


class B200EmbeddingBag:
    """Drop-in for the frozen mean-mode ``nn.EmbeddingBag`` the reference serves queries with."""

    def __init__(self, weight: torch.Tensor, padding_idx: Optional[int] = None):
        if weight.ndim != 2:
            raise ValueError(f"weight must be [V, d], got {tuple(weight.shape)}")
        if weight.dtype not in (torch.float32, torch.bfloat16):
            raise ValueError(f"weight dtype {weight.dtype} unsupported (float32 or bfloat16)")
        self.weight = weight.contiguous()
        self.num_embeddings, self.embedding_dim = self.weight.shape
        if padding_idx is not None:
            if padding_idx < 0:
                padding_idx += self.num_embeddings  # torch normalises negative padding_idx the same way
            if not 0 <= padding_idx < self.num_embeddings:
                raise ValueError("padding_idx must be within num_embeddings")
        self.padding_idx = padding_idx
        self.mode = "mean"
        self._err = None

    # -- construction / placement, as the reference uses them
    @classmethod
    def from_pretrained(cls, embeddings: torch.Tensor, freeze: bool = True, padding_idx: Optional[int] = None,
                        **_unused) -> "B200EmbeddingBag":
        return cls(embeddings, padding_idx=padding_idx)

    def to(self, device=None, dtype: Optional[torch.dtype] = None) -> "B200EmbeddingBag":
        w = self.weight
        if device is not None:
            w = w.to(device)
        if dtype is not None:
            w = w.to(dtype)
        self.weight = w.contiguous()
        self._err = None
        return self

    def cuda(self, device=None) -> "B200EmbeddingBag":
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))

    def bfloat16(self) -> "B200EmbeddingBag":
        return self.to(dtype=torch.bfloat16)

    # -- the hot call
    def encode(self, input: torch.Tensor, offsets: torch.Tensor, shrink_dim: Optional[int] = None,
               normalize: bool = False, out_dtype: Optional[torch.dtype] = None,
               out: Optional[torch.Tensor] = None, check_ids: bool = True) -> torch.Tensor:
        """mean-EmbeddingBag + ``[..., :shrink_dim]`` + optional ``F.normalize`` in one launch."""
        w = require_cuda(self.weight, "EmbeddingBag weight")
        dev = w.device
        if input.ndim != 1 or offsets.ndim != 1:
            raise ValueError("input and offsets must be 1-D (flattened ids + bag offsets)")
        ids = require_cuda(input, "input").to(torch.int64).contiguous()
        offs = require_cuda(offsets, "offsets").to(torch.int64).contiguous()
        n_bags = offs.numel()
        m = int(shrink_dim) if shrink_dim else self.embedding_dim
        out_dtype = out_dtype or w.dtype
        if out is None:
            out = torch.empty((n_bags, m), dtype=out_dtype, device=dev)
        elif out.shape != (n_bags, m) or not out.is_contiguous() or out.device != dev:
            raise ValueError("out has the wrong shape/device or is not contiguous")
        if self._err is None or self._err.device != dev:
            self._err = torch.zeros(1, dtype=torch.int32, device=dev)
        lib = _C.load()
        with torch.cuda.device(dev):
            _C.check(lib.lr_embbag_encode(
                ids.data_ptr(), offs.data_ptr(), ids.numel(), n_bags, w.data_ptr(), dtype_tag(w.dtype),
                self.num_embeddings, self.embedding_dim, -1 if self.padding_idx is None else self.padding_idx,
                m, int(bool(normalize)), out.data_ptr(), dtype_tag(out.dtype), self._err.data_ptr(),
                stream_ptr(dev)))
        if check_ids and int(self._err.item()) != 0:
            self._err.zero_()
            raise IndexError("EmbeddingBag: token id out of range [0, num_embeddings)")
        return out

    def forward(self, input: torch.Tensor, offsets: Optional[torch.Tensor] = None,
                per_sample_weights=None) -> torch.Tensor:
        if per_sample_weights is not None:
            raise NotImplementedError("per_sample_weights is not used by the reference's mean-mode bag")
        if input.ndim == 2:  # torch also accepts [B, L] fixed-length bags
            B, L = input.shape
            offsets = torch.arange(0, B * L, L, device=input.device)
            input = input.reshape(-1)
        if offsets is None:
            raise ValueError("offsets is required for 1-D input")
        return self.encode(input, offsets)

    __call__ = forward


def emb_bag_inputs(prompt_token_ids: Sequence[int], eos_token_id: int, start: int, end: int,
                   device=None) -> torch.Tensor:
    """Input ids of one table-construction batch: ``[bos] + prompt + [vocab id] + [eos]`` for vocab ids ``start..end-1``
    (reference finetune/nonctx_emb_utils.py:270-296; ``prompt_token_ids`` already holds the bos when the tokenizer adds one)."""
    n, p = end - start, len(prompt_token_ids)
    ids = torch.zeros((n, p + 2), dtype=torch.long, device=device)
    if p:
        ids[:, :p] = torch.tensor([list(prompt_token_ids)], dtype=torch.long, device=device)
    ids[:, -2] = torch.arange(start, end, dtype=torch.long, device=device)
    ids[:, -1] = eos_token_id
    return ids


def construct_embedding_bag(model, tokenizer, prompt: Optional[str] = None, batch_size: int = 2000,
                            table_dtype: torch.dtype = torch.float32) -> B200EmbeddingBag:
    """Builds the query encoder's table exactly as the reference does (finetune/nonctx_emb_utils.py:239-313): row v is the
    backbone's last-position hidden state of ``[bos] + prompt + [v] + [eos]`` for every v in ``[0, len(tokenizer))``,
    accumulated in fp32; ``padding_idx = tokenizer.pad_token_id``.  The LLM backbone is the caller's untouched torch module
    (north star: "PyTorch ... for the untouched LLM backbone"); what changes is the object returned — a ``B200EmbeddingBag``
    (``table_dtype=torch.bfloat16`` gives the table the reference's notebook serves, scripts/asymmetric_dense_infer.ipynb:50)."""
    model.eval()
    bos, eos, pad = tokenizer.bos_token_id, tokenizer.eos_token_id, tokenizer.pad_token_id
    n_vocab = len(tokenizer)
    add_bos = bos in tokenizer.encode("", add_special_tokens=True)  # fast tokenizers add it without exposing a switch
    prompt_ids = list(tokenizer.encode(prompt, add_special_tokens=False)) if prompt is not None else []
    if add_bos:
        prompt_ids.insert(0, bos)
    dev = model.device
    table = torch.zeros((n_vocab, model.config.hidden_size), dtype=torch.float32, device=dev)
    for start in range(0, n_vocab, batch_size):
        end = min(start + batch_size, n_vocab)
        ids = emb_bag_inputs(prompt_ids, eos, start, end, device=dev)
        with torch.no_grad(), torch.autocast(device_type=dev.type):
            out = model(input_ids=ids, return_dict=True, use_cache=False, output_hidden_states=False)
        table[start:end] = out.last_hidden_state[:, -1]
    return B200EmbeddingBag.from_pretrained(table.to(table_dtype), padding_idx=pad)


def tokenize_nonctx_qry_emb_bag(queries: Sequence[str], tokenizer, max_len: int = 512) -> dict:
    """Host restatement of reference finetune/nonctx_emb_utils.py:197-219 (flattened ids + bag offsets)."""
    encodings_ids = tokenizer(list(queries), max_length=max_len, truncation=True, add_special_tokens=False,
                              return_attention_mask=False)["input_ids"]
    return flatten_token_ids(encodings_ids)


def flatten_token_ids(token_id_lists: Sequence[Sequence[int]]) -> dict:
    """``offsets = cumsum([0] + lens[:-1])``, ``input_ids = concat(ids)`` (nonctx_emb_utils.py:217-218)."""
    lens = [len(x) for x in token_id_lists]
    offsets = torch.from_numpy(np.cumsum([0] + lens[:-1]).astype(np.int64))
    flat = np.concatenate([np.asarray(x, dtype=np.int64) for x in token_id_lists]) if sum(lens) else np.zeros(0, np.int64)
    return {"input_ids": torch.from_numpy(flat).long(), "offsets": offsets}


def lasttoken_head(last_hidden: torch.Tensor, attention_mask: torch.Tensor, shrink_dim: Optional[int] = None,
                   normalize: bool = False, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """``pooling(..., 'lasttoken')`` + shrink + normalize (dense_pooling.py:48-55, modeling_hybrid.py:266-278)."""
    h = require_cuda(last_hidden, "last_hidden").contiguous()
    if h.ndim != 3:
        raise ValueError("last_hidden must be [B, S, d]")
    B, S, d = h.shape
    mask = require_cuda(attention_mask, "attention_mask").to(torch.int64).contiguous()
    if mask.shape != (B, S):
        raise ValueError("attention_mask must be [B, S]")
    m = int(shrink_dim) if shrink_dim else d
    out = torch.empty((B, m), dtype=out_dtype or h.dtype, device=h.device)
    scratch = torch.empty(B + 1, dtype=torch.int32, device=h.device)
    lib = _C.load()
    with torch.cuda.device(h.device):
        _C.check(lib.lr_lasttoken_head(h.data_ptr(), dtype_tag(h.dtype), mask.data_ptr(), B, S, d, m,
                                       int(bool(normalize)), out.data_ptr(), dtype_tag(out.dtype),
                                       scratch.data_ptr(), stream_ptr(h.device)))
    return out
