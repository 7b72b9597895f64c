"""Llama-3 style decoder of BASELINE config 3 (reference llm/llama/model.py:10-269): RMSNorm blocks, interleaved-pair
RoPE, per-layer KV cache, SwiGLU feed-forward, greedy ``generate``.

Results follow the reference exactly — including ``generate``'s position bookkeeping: decode step i feeds the token of
position L+i-1 with start_pos = L+i (model.py:258-267), so one zero KV slot sits between prompt and generated tokens; and
``lm_head`` having a bias (model.py:190).  On a cuda device the per-token work runs as fused kernels (RoPE + cache append,
attention over the strided cache view, SwiGLU, RMSNorm, skinny GEMMs) instead of ~490 eager nodes per token.
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

    def hidden_fast(self, x):
        """Inference: gate|up as one GEMM on concatenated weight planes, then SwiGLU; the down projection is applied by the
        block together with the residual add."""
        gu = _fused.linear_cat(self, "_pdn_gate_up", x, (self.gate.weight, self.up.weight))
        return _fused.swiglu_rows_planes(gu, self.up_dim)


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
        if not self._train and not pdn.autograd.is_grad_enable() and _fused.usable(xq, xk, xv, op="llama_cached_attention"):
            self._rope_tables = self._rope_src()
            out = _fused.llama_cached_attention(self, xq, xk, xv, start_pos, mask, scale)
            return self.O(out)
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

    def _fast_ok(self, x, rows) -> bool:
        """Inference fast path applies: eval mode, no tape, fp32 cuda tensors, at least 32 token rows (below that the
        skinny FFMA GEMM of the generic path is the better kernel)."""
        return (os.environ.get("PDN_LLAMA_FAST", "1") != "0" and not self._train and not pdn.autograd.is_grad_enable()
                and rows >= 32 and _fused.usable(x, self.attention.Q.weight, op="llama_cached_attention"))

    def forward(self, x, start_pos, mask, freqs_cos, freqs_sin):
        if self._fast_ok(x, x.shape[0] * x.shape[1]):
            return self._forward_inference(x, start_pos, mask)
        z = x + self.attention(self.input_norm(x), start_pos, mask, freqs_cos, freqs_sin)
        return z + self.ffn(self.post_attn_norm(z))

    def _forward_inference(self, x, start_pos, mask):
        """Same block (reference llm/llama/model.py:142-150) with 9 launches instead of 22: RMSNorm emitting GEMM operand
        planes, fused QKV GEMM, RoPE + cache append, cached attention emitting operand planes, O-projection accumulating
        onto the residual stream in its epilogue, RMSNorm -> planes, fused gate|up GEMM, SwiGLU -> planes, down-projection
        accumulating onto the residual stream. ``x``'s buffer becomes the block output."""
        att = self.attention
        B, L, dim = x.shape
        H, D = att.n_heads, att.head_dim
        norm_in, norm_post = self.input_norm, self.post_attn_norm
        qkv = _fused.linear_cat(att, "_pdn_qkv", _fused.rmsnorm_planes(x, norm_in.weight, norm_in.eps), (att.Q.weight, att.K.weight, att.V.weight))
        with x.device:
            q3 = qkv.data.reshape(B, L, 3, H, D)
        att._rope_tables = att._rope_src()
        parts = [pdn.Tensor(q3[:, :, j], dtype=np.float32, copy=None, device=x.device) for j in range(3)]
        o = _fused.llama_cached_attention(att, parts[0], parts[1], parts[2], start_pos, mask, 1.0 / math.sqrt(D), ld=3 * dim, as_planes=True)
        z = _fused.linear_residual_(o, att.O.weight, x)
        return _fused.linear_residual_(self.ffn.hidden_fast(_fused.rmsnorm_planes(z, norm_post.weight, norm_post.eps)), self.ffn.down.weight, z)


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
        for layer in self.layers:  # the fused inference path reads the full RoPE tables by position
            layer.attention._rope_src = lambda: (self.freqs_cos.data, self.freqs_sin.data)
        self.norm = nn.RMSNorm(embed_dim, dtype=dtype)
        self.lm_head = nn.Linear(embed_dim, vocab_size, dtype=dtype)

    def _forward_hidden(self, input_ids, start_pos, final_norm=True):
        L = input_ids.shape[-1]
        h = self.tok_embedding(input_ids)
        if isinstance(start_pos, _fused.DevicePos):  # graph-recorded decode step: the fused kernels index the RoPE tables
            cos = sin = None
        else:
            cos, sin = self.freqs_cos[start_pos:start_pos + L], self.freqs_sin[start_pos:start_pos + L]
        mask = None
        if L > 1:  # causal mask over [cached positions | new positions], built on the host like the reference
            mask = np.concatenate([np.zeros((L, start_pos)), np.triu(np.full((L, L), float("-inf")), k=1)], axis=1)
            mask = pdn.Tensor(mask, device=h.device, dtype=h.dtype)
        for layer in self.layers:
            h = layer(h, start_pos, mask, cos, sin)
        return self.norm(h) if final_norm else h

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

    def _next_ids(self, input_ids, start_pos):
        """Greedy next token ids [B, 1] = argmax of the last position's logits (reference model.py:266-268). Inference fast
        path: final RMSNorm emitted as GEMM operand planes and the vocabulary argmax taken in the lm_head GEMM epilogue, so
        the [B, 32000] logits never reach HBM."""
        B = input_ids.shape[0]
        if self.layers[0]._fast_ok(self.norm.weight, input_ids.size) and os.environ.get("PDN_LM_HEAD_ARGMAX", "1") != "0":
            h = self._forward_hidden(input_ids, start_pos, final_norm=False)
            last = h if h.shape[1] == 1 else h[:, -1, :]
            pl = _fused.rmsnorm_planes(last, self.norm.weight, self.norm.eps)
            return _fused.lm_head_argmax(pl, self.lm_head.weight, self.lm_head.bias)
        return self(input_ids, start_pos)[:, -1, :].argmax(-1, True)

    def _graph_decode_ok(self, ids) -> bool:
        return (os.environ.get("PDN_DECODE_GRAPH", "1") != "0" and ids.device.is_cuda and not self._train
                and not pdn.autograd.is_grad_enable() and self.freqs_cos.dtype == np.float32
                and _fused.usable(self.freqs_cos, op="llama_cached_attention"))

    def generate(self, input_ids, max_new_tokens: int):
        """Greedy decoding; yields one (B, 1) id tensor per step until the total length reaches max_new_tokens
        (reference llm/llama/model.py:258-269, including its position bookkeeping: decode step i feeds the token produced
        at step i-1 with start_pos = L + i).

        On a cuda device the decode step is recorded ONCE as a CUDA graph (position and current ids live in device
        memory) and replayed for every further token: one graph launch per token instead of ~130 kernel launches driven
        from Python."""
        _, L = input_ids.shape
        next_id, graph, ids_buf, pos = None, None, None, None
        try:
            for i, curr_pos in enumerate(range(L, max_new_tokens)):
                if i == 0:  # prefill
                    next_id = self._next_ids(input_ids, 0)
                elif i == 1 or not self._graph_decode_ok(next_id):  # eager decode step (also warms caches for the capture)
                    next_id = self._next_ids(next_id, curr_pos)
                else:
                    assert curr_pos + 1 <= self.max_seq_len, "generation runs past the KV cache"
                    if graph is None:
                        with next_id.device:
                            ids_buf = pdn.Tensor(next_id.data, dtype=np.int64, device=next_id.device, copy=True)
                            pos = _fused.DevicePos(pdn.Tensor(np.array([curr_pos], dtype=np.int64), device=next_id.device))
                            graph = pdn.cuda.Graph()
                            graph.begin()
                            try:
                                nid = self._next_ids(ids_buf, pos)
                                ids_buf[...] = nid  # feeds the next replay
                                pos.tensor += 1
                            finally:
                                graph.end()
                            del nid
                    with next_id.device:
                        graph.launch()
                        next_id = pdn.Tensor(ids_buf.data, dtype=np.int64, device=ids_buf.device, copy=True)
                yield next_id
        finally:
            if graph is not None:
                with ids_buf.device:
                    graph.destroy()
