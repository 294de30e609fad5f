"""Parameter containers of the graph-attention encoder, state_dict-compatible with the reference
(graph_encoder.py:18-207): `layers.L.0.module.{W_query,W_key,W_val,W_out}`, `layers.L.{1,3}.normalizer.*`,
`layers.L.2.module.{0,2}.{weight,bias}`.

The eval-mode forward pass does NOT run here: AttentionModel packs these parameters into the weight
blob and calls the CUDA encoder (agh_encode).  The torch forward below exists for the differentiable
training pass only (train-mode BatchNorm with batch statistics + autograd through cuBLAS/SDPA),
see DESIGN.md "training".
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn


class SkipConnection(nn.Module):
    def __init__(self, module: nn.Module):
        super().__init__()
        self.module = module

    def forward(self, x):
        return x + self.module(x)


class MultiHeadAttention(nn.Module):
    """Per-head projection tensors [H, D, d] / [H, d, D], uniform(+-1/sqrt(last dim)) init
    (graph_encoder.py:42-54); no biases."""

    def __init__(self, n_heads: int, input_dim: int, embed_dim: int):
        super().__init__()
        d = embed_dim // n_heads
        self.n_heads, self.input_dim, self.embed_dim, self.key_dim, self.val_dim = n_heads, input_dim, embed_dim, d, d
        self.W_query = nn.Parameter(torch.empty(n_heads, input_dim, d))
        self.W_key = nn.Parameter(torch.empty(n_heads, input_dim, d))
        self.W_val = nn.Parameter(torch.empty(n_heads, input_dim, d))
        self.W_out = nn.Parameter(torch.empty(n_heads, d, embed_dim))
        for p in self.parameters():
            bound = 1.0 / math.sqrt(p.size(-1))
            p.data.uniform_(-bound, bound)

    def forward(self, x):
        # x [B,n,D] -> per-head projections [B,H,n,d]; softmax(QK^T/sqrt(d)) V; concat heads; W_out
        q = torch.einsum('bnc,hck->bhnk', x, self.W_query)
        k = torch.einsum('bnc,hck->bhnk', x, self.W_key)
        v = torch.einsum('bnc,hck->bhnk', x, self.W_val)
        o = F.scaled_dot_product_attention(q, k, v)               # default scale 1/sqrt(d)
        return torch.einsum('bhnk,hkc->bnc', o, self.W_out)


class Normalization(nn.Module):
    def __init__(self, embed_dim: int, normalization: str = 'batch'):
        super().__init__()
        if normalization != 'batch':
            raise NotImplementedError("only normalization='batch' is on the AGH path (all shipped args.json)")
        self.normalizer = nn.BatchNorm1d(embed_dim, affine=True)

    def forward(self, x):
        return self.normalizer(x.reshape(-1, x.size(-1))).view_as(x)


class MultiHeadAttentionLayer(nn.Sequential):
    def __init__(self, n_heads: int, embed_dim: int, feed_forward_hidden: int = 512, normalization: str = 'batch'):
        super().__init__(
            SkipConnection(MultiHeadAttention(n_heads, embed_dim, embed_dim)),
            Normalization(embed_dim, normalization),
            SkipConnection(nn.Sequential(nn.Linear(embed_dim, feed_forward_hidden), nn.ReLU(),
                                         nn.Linear(feed_forward_hidden, embed_dim))),
            Normalization(embed_dim, normalization))


class GraphAttentionEncoder(nn.Module):
    def __init__(self, n_heads: int, embed_dim: int, n_layers: int, node_dim=None, normalization: str = 'batch',
                 feed_forward_hidden: int = 512):
        super().__init__()
        assert node_dim is None, 'AGH embeds nodes in AttentionModel._init_embed'
        self.layers = nn.Sequential(*[MultiHeadAttentionLayer(n_heads, embed_dim, feed_forward_hidden, normalization)
                                      for _ in range(n_layers)])

    def forward(self, x, mask=None):
        assert mask is None
        h = self.layers(x)
        return h, h.mean(dim=1)
