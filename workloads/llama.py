"""Llama-3 style decoder of BASELINE config 3 (reference llm/llama/model.py:10-269): RMSNorm blocks, interleaved-pair
RoPE, per-layer KV cache, SwiGLU feed-forward, greedy ``generate``.

Results follow the reference exactly — including ``generate``'s position bookkeeping: decode step i feeds the token of
position L+i-1 with start_pos = L+i (model.py:258-267), so one zero KV slot sits between prompt and generated tokens; and
``lm_head`` having a bias (model.py:190).  This file is a plain model definition with the reference's structure and parameter
names: on a cuda device in eval mode ``Module.__call__`` serves it — exactly like the reference's own unchanged model.py —
through the inference plan of pydynet_b200/nn/_plans.py (fused blocks, CUDA-graph / persistent-kernel decode); nothing
model-specific lives here.  It is the stand-in for the reference's file where that is not available.
"""
import math
import os

import numpy as np

import pydynet_b200 as pdn
import pydynet_b200.nn as nn
import pydynet_b200.nn.functional as F
from pydynet_b200.core.tensor import Tensor
from pydynet_b200.nn import _fused


def compute_cos_sin_cache(head_dim: int, max_seq_len: int, base: int = 10000, dtype=None):
    inv_freq = 1.0 / (base**(np.arange(0, head_dim, 2)[:(head_dim // 2)] / head_dim))
    freqs = np.outer(np.arange(max_seq_len), inv_freq).astype(dtype)
    return Tensor(np.cos(freqs)), Tensor(np.sin(freqs))


def synthetic_llama_params(V, D, H, FF, n_layers, seed=0, std=0.05, dtype=np.float32):
    """Random-init weights of the BASELINE config-3 architecture keyed by the reference's ``_parameters`` names
    (llm/llama/model.py:216): matrices N(0, std), norm weights 1, lm_head bias N(0, std) — SURVEY.md §8(d) C3. Draw order is
    part of the contract (tests compare against the oracle's own generator bit for bit)."""
    rng = np.random.default_rng(seed)
    draw = lambda *shape: (rng.standard_normal(shape) * std).astype(dtype)
    out = {"tok_embedding.weight": draw(V, D), "norm.weight": np.ones(D, dtype), "lm_head.weight": draw(D, V), "lm_head.bias": draw(V)}
    for i in range(n_layers):
        pre = f"layers.{i}."
        for nm in "QKVO":
            out[pre + f"attention.{nm}.weight"] = draw(D, D)
        out[pre + "ffn.up.weight"] = draw(D, FF)
        out[pre + "ffn.gate.weight"] = draw(D, FF)
        out[pre + "ffn.down.weight"] = draw(FF, D)
        out[pre + "input_norm.weight"] = np.ones(D, dtype)
        out[pre + "post_attn_norm.weight"] = np.ones(D, dtype)
    return out


def apply_rotary_emb(xq, xk, freqs_cos, freqs_sin):
    """Rotates consecutive (even, odd) feature pairs of every head by the position angle."""
    cos, sin = pdn.unsqueeze(freqs_cos, axis=-2), pdn.unsqueeze(freqs_sin, axis=-2)

    def rot(x):
        pairs = x.reshape(*(x.shape[:-1] + (-1, 2)))
        re, im = pairs[..., 0], pairs[..., 1]
        out = pdn.concat([pdn.unsqueeze(re * cos - im * sin, -1), pdn.unsqueeze(re * sin + im * cos, -1)], axis=-1)
        return out.reshape(*(out.shape[:-2] + (-1, )))

    return rot(xq), rot(xk)


class FeedForward(nn.Module):

    def __init__(self, dim, up_dim, dtype=None):
        super().__init__()
        self.dim, self.up_dim = dim, up_dim
        self.up = nn.Linear(dim, up_dim, bias=False, dtype=dtype)
        self.gate = nn.Linear(dim, up_dim, bias=False, dtype=dtype)
        self.down = nn.Linear(up_dim, dim, bias=False, dtype=dtype)

    def forward(self, x):
        if _fused.usable(x, op="swiglu"):
            return self.down(_fused.swiglu(self.gate(x), self.up(x)))
        return self.down(F.silu(self.gate(x)) * self.up(x))


class Attention(nn.Module):

    def __init__(self, dim: int, n_heads: int, max_seq_len: int, max_batch_size: int = None, dtype=None):
        super().__init__()
        assert dim % n_heads == 0
        self.dim, self.n_heads, self.head_dim = dim, n_heads, dim // n_heads
        for name in "QKVO":
            setattr(self, name, nn.Linear(dim, dim, bias=False, dtype=dtype))
        self.max_seq_len = max_seq_len
        self.max_batch_size = max_batch_size if max_batch_size is not None else 1
        shape = (self.max_batch_size, max_seq_len, n_heads, self.head_dim)
        self.cache_k = nn.Parameter(pdn.special.zeros(shape, dtype=dtype), requires_grad=False)
        self.cache_v = nn.Parameter(pdn.special.zeros(shape, dtype=dtype), requires_grad=False)

    def __call__(self, x, start_pos: int, mask, freqs_cos, freqs_sin):
        B, L, _ = x.shape
        H, D = self.n_heads, self.head_dim
        xq, xk, xv = (proj(x).reshape(B, L, H, D) for proj in (self.Q, self.K, self.V))
        scale = 1.0 / math.sqrt(D)
        xq, xk = apply_rotary_emb(xq, xk, freqs_cos, freqs_sin)
        if not self._train:
            self.cache_k[:B, start_pos:start_pos + L] = xk
            self.cache_v[:B, start_pos:start_pos + L] = xv
            xk = self.cache_k[:B, :start_pos + L]
            xv = self.cache_v[:B, :start_pos + L]
        if _fused.usable(xq, xk, xv, op="attention") and _fused.attention_fits(xq, xk):
            out = _fused.attention(xq, xk, xv, mask, scale)
        else:
            scores = xq.transpose(0, 2, 1, 3) @ xk.transpose(0, 2, 3, 1) / math.sqrt(D)
            if mask is not None:
                scores = scores + mask
            out = F.softmax(scores, axis=-1) @ xv.transpose(0, 2, 1, 3)
            out = out.transpose(0, 2, 1, 3).reshape(B, L, -1)
        return self.O(out)


class TransformerBlock(nn.Module):

    def __init__(self, dim, n_heads, ffn_dim, max_seq_len, max_batch_size=None, dtype=None):
        super().__init__()
        self.attention = Attention(dim, n_heads, max_seq_len, max_batch_size, dtype)
        self.ffn = FeedForward(dim, ffn_dim, dtype)
        self.input_norm = nn.RMSNorm(dim, dtype=dtype)
        self.post_attn_norm = nn.RMSNorm(dim, dtype=dtype)

    def forward(self, x, start_pos, mask, freqs_cos, freqs_sin):
        z = x + self.attention(self.input_norm(x), start_pos, mask, freqs_cos, freqs_sin)
        return z + self.ffn(self.post_attn_norm(z))


class Llama(nn.Module):

    def __init__(self, vocab_size, embed_dim, n_heads, ffn_dim: int, max_seq_len: int, max_batch_size: int = None,
                 n_layers: int = 6, dtype=None):
        super().__init__()
        self.vocab_size, self.embed_dim, self.n_heads, self.ffn_dim = vocab_size, embed_dim, n_heads, ffn_dim
        self.max_seq_len, self.max_batch_size, self.n_layers = max_seq_len, max_batch_size, n_layers
        self.tok_embedding = nn.Embedding(vocab_size, embed_dim, dtype=dtype)
        cos, sin = compute_cos_sin_cache(embed_dim // n_heads, max_seq_len, dtype=dtype)
        self.freqs_cos = nn.Parameter(cos, False)
        self.freqs_sin = nn.Parameter(sin, False)
        self.layers = nn.ModuleList(
            [TransformerBlock(embed_dim, n_heads, ffn_dim, max_seq_len, max_batch_size, dtype) for _ in range(n_layers)])
        self.norm = nn.RMSNorm(embed_dim, dtype=dtype)
        self.lm_head = nn.Linear(embed_dim, vocab_size, dtype=dtype)

    def _forward_hidden(self, input_ids, start_pos):
        L = input_ids.shape[-1]
        h = self.tok_embedding(input_ids)
        cos, sin = self.freqs_cos[start_pos:start_pos + L], self.freqs_sin[start_pos:start_pos + L]
        mask = None
        if L > 1:  # causal mask over [cached positions | new positions], built on the host like the reference
            mask = np.concatenate([np.zeros((L, start_pos)), np.triu(np.full((L, L), float("-inf")), k=1)], axis=1)
            mask = pdn.Tensor(mask, device=h.device, dtype=h.dtype)
        for layer in self.layers:
            h = layer(h, start_pos, mask, cos, sin)
        return self.norm(h)

    def forward_logits(self, input_ids, start_pos: int = 0):
        return self.lm_head(self._forward_hidden(input_ids, start_pos))

    def set_trainable_parameters(self, trainable_prefixes=("lm_head", )):
        n_train = 0
        for name, p in self._parameters.items():
            p.requires_grad = any(name.startswith(pre) for pre in trainable_prefixes)
            n_train += p.requires_grad
        return n_train, len(self._parameters) - n_train

    def finetune_step(self, input_ids, target_ids, optimizer, criterion=None, start_pos: int = 0):
        criterion = criterion if criterion is not None else nn.CrossEntropyLoss()
        self.train(True)
        optimizer.zero_grad()
        logits = self.forward_logits(input_ids, start_pos)
        B, L, V = logits.shape
        targets = pdn.Tensor(np.asarray(target_ids).reshape(-1), dtype=np.int64, device=logits.device)
        loss = criterion(logits.reshape(B * L, V), targets)
        loss.backward()
        optimizer.step()
        return loss.item()

    def forward(self, input_ids, start_pos):
        h = self._forward_hidden(input_ids, start_pos)
        return self.lm_head(h if h.shape[1] == 1 else h[:, [-1], :])  # logits of the last position, [B, 1, V]

    def generate(self, input_ids, max_new_tokens: int):
        """Greedy decoding; yields one (B, 1) id tensor per step until the total length reaches max_new_tokens (reference
        llm/llama/model.py:258-269, including its position bookkeeping: decode step i feeds the token produced at step i-1 with
        start_pos = L + i)."""
        _, L = input_ids.shape
        next_id = None
        for i, curr_pos in enumerate(range(L, max_new_tokens)):
            logits = self(input_ids, 0) if i == 0 else self(next_id, curr_pos)
            next_id = logits[:, -1, :].argmax(-1, True)
            yield next_id
