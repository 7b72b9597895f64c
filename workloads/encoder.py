"""Transformer encoder classifier of BASELINE config 4 (reference examples/pydynet/transformer.py:53-192).

Faithful to the reference's behaviour including its oddities: SelfAttention projects ``values`` for Q, K and V
(transformer.py:81-91); "LayerNorm" is the reference's batch-statistic norm (nn/modules/norm.py); dropout is accepted and
unused; padding mask entries are turned into -inf in place and added to the scores."""
import numpy as np

import pydynet_b200 as pdn
import pydynet_b200.nn as nn
import pydynet_b200.nn.functional as F
from pydynet_b200.nn import _fused


class SelfAttention(nn.Module):

    def __init__(self, embed_size, heads):
        super().__init__()
        self.embed_size, self.heads = embed_size, heads
        self.head_dim = embed_size // heads
        assert self.head_dim * heads == embed_size, "Embedding size needs to be divisible by heads"
        for name in "QKVO":
            setattr(self, name, nn.Linear(embed_size, embed_size, bias=False, dtype=np.float32))

    def forward(self, values, keys, query, mask):
        N, L = query.shape[0], values.shape[1]
        H, D = self.heads, self.head_dim
        xq, xk, xv = (proj(values).reshape(N, L, H, D) for proj in (self.Q, self.K, self.V))
        if mask is not None:
            mask[mask.eq(1)] = np.float32('-inf')
        if _fused.usable(xq, xk, xv, op="attention") and _fused.attention_fits(xq, xk):
            out = _fused.attention(xq, xk, xv, mask, 1.0 / D**.5)  # (N, L, H*D)
        else:
            scores = xq.transpose(0, 2, 1, 3) @ xk.transpose(0, 2, 3, 1) / D**.5
            if mask is not None:
                scores = scores + mask
            out = F.softmax(scores, axis=-1) @ xv.transpose(0, 2, 1, 3)
            out = out.transpose(0, 2, 1, 3).reshape(N, L, -1)
        return self.O(out)


class TransformerBlock(nn.Module):

    def __init__(self, embed_size, heads, dropout, forward_expansion):
        super().__init__()
        self.attention = SelfAttention(embed_size, heads)
        self.norm1 = nn.LayerNorm(embed_size, dtype=np.float32)
        self.norm2 = nn.LayerNorm(embed_size, dtype=np.float32)
        self.feed_forward = nn.Sequential(
            nn.Linear(embed_size, forward_expansion * embed_size, dtype=np.float32),
            nn.ReLU(),
            nn.Linear(forward_expansion * embed_size, embed_size, dtype=np.float32),
        )

    def forward(self, value, key, query, mask):
        x = self.norm1(self.attention(value, key, query, mask) + query)
        return self.norm2(self.feed_forward(x) + x)


def sinusoidal_positional_encoding(max_len: int, d_model: int):
    pos = np.arange(max_len)[:, None]
    div = np.exp(np.arange(0, d_model, 2) * (-np.log(10000.0) / d_model))
    pe = np.zeros((max_len, d_model))
    pe[:, 0::2], pe[:, 1::2] = np.sin(pos * div), np.cos(pos * div)
    return pdn.Tensor(pe.astype(np.float32))


@pdn.no_grad()
def construct_mask(x, padding_idx=0):
    """[batch, 1, 1, seq] float mask, 1 where the token is padding."""
    return pdn.unsqueeze(x.eq(padding_idx), (1, 2)).astype(np.float32)


class Transformer(nn.Module):

    def __init__(self, embed_size, num_layers, heads, forward_expansion, dropout, vocab_size, max_length):
        super().__init__()
        self.embed_size = embed_size
        self.word_embedding = nn.Embedding(vocab_size, embed_size, padding_idx=0, dtype=np.float32)
        self.position_embedding = nn.Parameter(sinusoidal_positional_encoding(max_length, embed_size), False)
        self.layers = nn.ModuleList([TransformerBlock(embed_size, heads, dropout, forward_expansion) for _ in range(num_layers)])
        self.fc_out = nn.Linear(embed_size, 1, dtype=np.float32)

    def forward(self, x, mask):
        out = self.word_embedding(x) + self.position_embedding
        for layer in self.layers:
            out = layer(out, out, out, mask)
        return self.fc_out(out[:, 0, :])


def logistic_loss(out, y):
    """mean(log(1 + exp(-y * out))) — reference transformer.py:244-245 with unit weights."""
    return pdn.log(1 + pdn.exp(-y * pdn.squeeze(out))).mean()


def train_step(net, optimizer, X, y, mask=None):
    loss = logistic_loss(net(X, mask), y)
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return loss
