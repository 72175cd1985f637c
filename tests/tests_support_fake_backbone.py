"""A deterministic stand-in for the LLM backbone and its tokenizer, shared by oracle/gen_golden.py (which feeds it to the
REFERENCE's construct_embedding_bag to record golden tables) and tests/test_oracle_golden.py (which feeds it to ours).
No matmul inside: under ``torch.autocast('cpu')`` the output stays fp32 and bit-reproducible."""
import types

import torch


class FakeTokenizer:
    bos_token_id, eos_token_id, pad_token_id = 1, 2, 0

    def __init__(self, n_vocab: int, add_bos: bool):
        self.n_vocab, self.add_bos = n_vocab, add_bos

    def __len__(self):
        return self.n_vocab

    def encode(self, text, add_special_tokens=True):
        ids = [3 + (ord(c) % (self.n_vocab - 3)) for c in text]
        return ([self.bos_token_id] if (add_special_tokens and self.add_bos) else []) + ids


class FakeBackbone:
    """last_hidden_state[b, t] = sum_{s<=t} (s+1) * E[ids[b, s]]  (position-dependent, depends on every token)."""

    def __init__(self, n_vocab: int, hidden: int):
        g = torch.Generator().manual_seed(7)
        self.E = torch.randn(n_vocab, hidden, generator=g)
        self.config = types.SimpleNamespace(hidden_size=hidden)
        self.device = torch.device("cpu")
        self.seen = []

    def eval(self):
        return self

    def __call__(self, input_ids, return_dict=True, use_cache=False, output_hidden_states=False):
        assert return_dict and not use_cache and not output_hidden_states
        self.seen.append(input_ids.clone().numpy())
        pos = torch.arange(1, input_ids.shape[1] + 1, dtype=torch.float32)[None, :, None]
        return types.SimpleNamespace(last_hidden_state=torch.cumsum(self.E[input_ids] * pos, dim=1))
