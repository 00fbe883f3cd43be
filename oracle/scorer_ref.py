"""Oracle restatement of allRank's scorer: FC input block -> pre-norm Transformer encoder ->
linear head (TEST INFRASTRUCTURE, see oracle/__init__.py).

Reference: /root/reference/allrank/models/model.py      FCModel :12-44, LTRModel :47-92,
                                                         OutputLayer :95-128, make_model :131-151
           /root/reference/allrank/models/transformer.py Encoder :28-56, LayerNorm :59-81,
                                                         SublayerConnection :84-106, EncoderLayer :109-134,
                                                         attention :137-156, MultiHeadedAttention :159-203,
                                                         PositionwiseFeedForward :206-227

The module tree below reproduces the reference's state_dict keys exactly
(`input_layer.layers.{i}`, `encoder.layers.{l}.self_attn.linears.{0..3}`,
`encoder.layers.{l}.feed_forward.w_{1,2}`, `encoder.layers.{l}.sublayer.{0,1}.norm.{a_2,b_2}`,
`encoder.norm.{a_2,b_2}`, `output_layer.w_1`), so weights move both ways with load_state_dict.
It is eager PyTorch and is what bench.py times as the CPU arm.

Parity is UNPINNED by the reference's own tests (the Transformer has none); it is pinned by
tests/golden/scorer_*.npz generated from the reference's make_model (oracle/make_golden.py).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class RowNorm(nn.Module):
    """transformer.py:59-81 -- NOT nn.LayerNorm: unbiased std, eps added to the std."""

    def __init__(self, width, eps=1e-6):
        super().__init__()
        self.a_2 = nn.Parameter(torch.ones(width))
        self.b_2 = nn.Parameter(torch.zeros(width))
        self.eps = eps

    def forward(self, x):
        mu = x.mean(-1, keepdim=True)
        sd = x.std(-1, keepdim=True)
        return self.a_2 * (x - mu) / (sd + self.eps) + self.b_2


class SelfAttention(nn.Module):
    """transformer.py:159-203 + attention() :137-156.  Key-only masking (quirk Q1)."""

    def __init__(self, heads, width, dropout):
        super().__init__()
        assert width % heads == 0
        self.h, self.d_k = heads, width // heads
        self.linears = nn.ModuleList([nn.Linear(width, width) for _ in range(4)])
        self.dropout = nn.Dropout(dropout)

    def forward(self, x, key_is_pad):
        b, s, _ = x.shape
        q, k, v = (lin(x).view(b, s, self.h, self.d_k).transpose(1, 2) for lin in self.linears[:3])
        logits = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(self.d_k)
        logits = logits.masked_fill(key_is_pad[:, None, None, :], float("-inf"))
        p = self.dropout(F.softmax(logits, dim=-1))
        ctx = torch.matmul(p, v).transpose(1, 2).contiguous().view(b, s, self.h * self.d_k)
        return self.linears[3](ctx)


class FeedForward(nn.Module):
    """transformer.py:206-227"""

    def __init__(self, width, hidden, dropout):
        super().__init__()
        self.w_1 = nn.Linear(width, hidden)
        self.w_2 = nn.Linear(hidden, width)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x):
        return self.w_2(self.dropout(F.relu(self.w_1(x))))


class Residual(nn.Module):
    """transformer.py:84-106: x + dropout(f(norm(x)))"""

    def __init__(self, width, dropout):
        super().__init__()
        self.norm = RowNorm(width)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x, f):
        return x + self.dropout(f(self.norm(x)))


class Block(nn.Module):
    """transformer.py:109-134"""

    def __init__(self, width, heads, hidden, dropout):
        super().__init__()
        self.self_attn = SelfAttention(heads, width, dropout)
        self.feed_forward = FeedForward(width, hidden, dropout)
        self.sublayer = nn.ModuleList([Residual(width, dropout), Residual(width, dropout)])
        self.size = width

    def forward(self, x, key_is_pad):
        x = self.sublayer[0](x, lambda y: self.self_attn(y, key_is_pad))
        return self.sublayer[1](x, self.feed_forward)


class Position(nn.Module):
    """allrank/models/positional.py:15-77: x = sqrt(d) x + table[idx]; padded items and indices beyond the table use
    the last (padding) row.  strategy "fixed": sinusoidal buffer `pe`; "learned": nn.Embedding `pe`."""

    def __init__(self, width, max_len, strategy):
        super().__init__()
        self.strategy = strategy
        if strategy == "fixed":
            table = torch.zeros(max_len, width)
            pos = torch.arange(0.0, max_len).unsqueeze(1)
            freq = torch.exp(torch.arange(0.0, width, 2) * -(math.log(10000.0) / width))
            table[:, 0::2] = torch.sin(pos * freq)
            table[:, 1::2] = torch.cos(pos * freq)
            self.register_buffer("pe", torch.cat((table, torch.zeros([1, width]))))
        else:
            self.pe = nn.Embedding(max_len + 1, width, padding_idx=-1)

    def forward(self, x, mask, indices):
        table = self.pe if self.strategy == "fixed" else self.pe.weight
        last = table.shape[0] - 1
        idx = indices.masked_fill(mask, last)
        idx = torch.where(idx > last, torch.full_like(idx, last), idx)
        looked_up = table[idx] if self.strategy == "fixed" else self.pe(idx)
        return math.sqrt(table.shape[1]) * x + looked_up


class Stack(nn.Module):
    """transformer.py:28-56"""

    def __init__(self, n_layers, width, heads, hidden, dropout, position=None):
        super().__init__()
        self.layers = nn.ModuleList([Block(width, heads, hidden, dropout) for _ in range(n_layers)])
        self.norm = RowNorm(width)
        self.position = position

    def forward(self, x, mask, indices):
        if self.position is not None:
            x = self.position(x, mask, indices)
        for blk in self.layers:
            x = blk(x, mask)
        return self.norm(x)


class InputFC(nn.Module):
    """model.py:12-44 (activation None -> identity; optional nn.LayerNorm on the input)."""

    def __init__(self, sizes, n_features, input_norm=False, activation=None, dropout=0.0):
        super().__init__()
        dims = [n_features] + list(sizes)
        self.input_norm = nn.LayerNorm(n_features) if input_norm else nn.Identity()
        self.activation = nn.Identity() if activation is None else getattr(nn, activation)()
        self.dropout = nn.Dropout(dropout or 0.0)
        self.layers = nn.ModuleList([nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])
        self.output_size = dims[-1]

    def forward(self, x):
        x = self.input_norm(x)
        for lin in self.layers:
            x = self.dropout(self.activation(lin(x)))
        return x


class Head(nn.Module):
    """model.py:95-128"""

    def __init__(self, width, d_output=1, output_activation=None):
        super().__init__()
        self.activation = nn.Identity() if output_activation is None else getattr(nn, output_activation)()
        self.d_output = d_output
        self.w_1 = nn.Linear(width, d_output)

    def forward(self, x):
        return self.activation(self.w_1(x).squeeze(dim=2))

    def score(self, x):
        y = self.forward(x)
        return y.sum(-1) if self.d_output > 1 else y


class RefLTRModel(nn.Module):
    """model.py:47-92"""

    def __init__(self, input_layer, encoder, output_layer):
        super().__init__()
        self.input_layer = input_layer
        self.encoder = encoder
        self.output_layer = output_layer

    def prepare_for_output(self, x, mask, indices):
        h = self.input_layer(x)
        return h if self.encoder is None else self.encoder(h, mask, indices)   # model.py:59: no encoder -> identity

    def forward(self, x, mask, indices):
        return self.output_layer(self.prepare_for_output(x, mask, indices))

    def score(self, x, mask, indices):
        return self.output_layer.score(self.prepare_for_output(x, mask, indices))


def make_ref_model(n_features, fc_sizes, n_layers, heads, d_ff, dropout=0.0, d_output=1,
                   output_activation=None, fc_activation=None, seed=None, positional=None, input_norm=False):
    """model.py:131-151: build + xavier_uniform_ on every parameter with dim > 1."""
    if seed is not None:
        torch.manual_seed(seed)
    fc = InputFC(fc_sizes, n_features, input_norm=input_norm, activation=fc_activation)
    width = fc.output_size
    position = Position(width, positional[1], positional[0]) if positional else None
    enc = Stack(n_layers, width, heads, d_ff, dropout, position) if n_layers > 0 else None   # transformer=None
    model = RefLTRModel(fc, enc, Head(width, d_output, output_activation))
    for p in model.parameters():
        if p.dim() > 1:
            nn.init.xavier_uniform_(p)
    return model
