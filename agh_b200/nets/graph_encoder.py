"""Parameter containers of the graph-attention encoder, state_dict-compatible with the reference
(graph_encoder.py:18-207): `layers.L.0.module.{W_query,W_key,W_val,W_out}`, `layers.L.{1,3}.normalizer.*`,
`layers.L.2.module.{0,2}.{weight,bias}`.

No forward pass runs here: AttentionModel packs these parameters into the weight blob and calls the CUDA encoder
(agh_encode in eval mode; the training primitives of csrc/train.cu with batch-statistics BatchNorm and a hand-written
backward in train mode, nets/train_fn.py).  The modules exist so that `state_dict()` / `load_state_dict()` /
`parameters()` / pickling behave exactly like the reference's.
"""
from __future__ import annotations

import math

import torch
from torch import nn

_NO_FORWARD = 'parameter container only: the encoder runs inside libagh_b200 (AttentionModel.embed / forward)'


class SkipConnection(nn.Module):
    def __init__(self, module: nn.Module):
        super().__init__()
        self.module = module

    def forward(self, x):
        raise RuntimeError(_NO_FORWARD)


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
        raise RuntimeError(_NO_FORWARD)


class Normalization(nn.Module):
    def __init__(self, embed_dim: int, normalization: str = 'batch'):
        super().__init__()
        if normalization != 'batch':
            raise NotImplementedError("only normalization='batch' is on the AGH path (all shipped args.json)")
        self.normalizer = nn.BatchNorm1d(embed_dim, affine=True)

    def forward(self, x):
        raise RuntimeError(_NO_FORWARD)


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
        raise RuntimeError(_NO_FORWARD)
