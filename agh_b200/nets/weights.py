"""Pack a reference-compatible state_dict into the f32 weight blob libagh_b200 consumes.

Layout (floats) is declared in include/agh_b200.h and agh_b200/csrc/common.cuh.  Algebraic folds
applied here, in f64, once per weight version:
  * eval-mode BatchNorm1d (graph_encoder.py:135-138) -> per-channel scale/shift;
  * project_out (attention_model.py:152,568) is absorbed into the logit keys:
      logits_j = (W_out o) . LK_j / sqrt(128) = o . (W_out^T LK_j) / sqrt(128)
    so the decoder projection for LK' is (W_out^T W_node[256:384] / sqrt(128));
  * the node part of project_step_context (attention_model.py:149,432-433) becomes a per-node
    table P_sc = h W_sc[:, :128]^T computed once per instance; the capacity and time columns stay
    as two vectors; the 1/sqrt(16) glimpse scale (attention_model.py:559) is folded into K.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

D, H, FF, LAYERS = 128, 8, 512, 3
BN_EPS = 1e-5


def blob_floats() -> int:
    layer = D * 3 * D + D * D + 2 * D + D * FF + FF + FF * D + D + 2 * D
    tc = LAYERS * (3 * D * D + D * D + FF * D + D * FF) + 4 * D * D      # K-major copies for the tcgen05 GEMMs
    return 92 * D + 3 * D + D + LAYERS * layer + D * 4 * D + D * D + 10 * D + 2 * D + tc


def offsets() -> Dict[str, int]:
    """Float offsets of the blob sections (mirror of agh_b200/csrc/common.cuh); per-layer sections are relative to
    'layer0' + l * 'layer_size'."""
    o = {'loc_emb': 0}
    o['init_w'] = o['loc_emb'] + 92 * D
    o['init_b'] = o['init_w'] + 3 * D
    o['layer0'] = o['init_b'] + D
    l = {'wqkv': 0}
    l['wo'] = l['wqkv'] + D * 3 * D
    l['bn1s'] = l['wo'] + D * D
    l['bn1t'] = l['bn1s'] + D
    l['w1t'] = l['bn1t'] + D
    l['b1'] = l['w1t'] + D * FF
    l['w2t'] = l['b1'] + FF
    l['b2'] = l['w2t'] + FF * D
    l['bn2s'] = l['b2'] + D
    l['bn2t'] = l['bn2s'] + D
    o['layer_size'] = l['bn2t'] + D
    o.update({'l_' + k: v for k, v in l.items()})
    o['wdec'] = o['layer0'] + LAYERS * o['layer_size']
    o['wfc_t'] = o['wdec'] + D * 4 * D
    o['fleet'] = o['wfc_t'] + D * D
    o['w_cap'] = o['fleet'] + 10 * D
    o['w_time'] = o['w_cap'] + D
    # K-major ([out][in]) copies: per layer wqkv^T [384][128] | wo^T [128][128] | w1 [512][128] | w2 [128][512]; then wdec^T
    o['tc0'] = o['w_time'] + D
    o['t_wqkv'] = 0
    o['t_wo'] = o['t_wqkv'] + 3 * D * D
    o['t_w1'] = o['t_wo'] + D * D
    o['t_w2'] = o['t_w1'] + FF * D
    o['t_size'] = o['t_w2'] + D * FF
    o['tc_dec'] = o['tc0'] + LAYERS * o['t_size']
    return o


def _matmul_f64(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a @ b in float64.  On the GPU through the library's own kernel (agh_gemm_f64): the product path launches no
    cuBLAS kernel, not even for this 128^3 fold; on the CPU (host-logic tests) plain torch."""
    if not a.is_cuda:
        return a @ b
    from .. import _lib
    a, b = a.contiguous(), b.contiguous()
    c = torch.empty(a.size(0), b.size(1), dtype=torch.float64, device=a.device)
    _lib.check(_lib.lib().agh_gemm_f64(a.data_ptr(), a.size(1), 0, b.data_ptr(), b.size(1), 0, c.data_ptr(), c.size(1),
                                       a.size(0), b.size(1), a.size(1), 1.0, _lib.stream_ptr(a.device)))
    return c


def pack_weights(sd: Dict[str, torch.Tensor], device=None) -> torch.Tensor:
    g = lambda k: sd[k].detach().to(device=device, dtype=torch.float64)
    parts = [g('loc_embedding.weight'), g('init_embed.weight').t(), g('init_embed.bias')]
    tc_parts = []
    for l in range(LAYERS):
        pre = 'embedder.layers.%d.' % l
        heads = lambda k: g(pre + '0.module.' + k).permute(1, 0, 2).reshape(D, D)       # [c][h*16+k]
        parts.append(torch.cat((heads('W_query'), heads('W_key'), heads('W_val')), 1))   # [128][384]
        parts.append(g(pre + '0.module.W_out').reshape(D, D))                            # [h*16+k][c]
        tc_parts += [parts[-2].t(), parts[-1].t(), g(pre + '2.module.0.weight'), g(pre + '2.module.2.weight')]
        s1 = g(pre + '1.normalizer.weight') / torch.sqrt(g(pre + '1.normalizer.running_var') + BN_EPS)
        t1 = g(pre + '1.normalizer.bias') - g(pre + '1.normalizer.running_mean') * s1
        parts += [s1, t1]
        parts += [g(pre + '2.module.0.weight').t(), g(pre + '2.module.0.bias'),
                  g(pre + '2.module.2.weight').t(), g(pre + '2.module.2.bias')]
        s2 = g(pre + '3.normalizer.weight') / torch.sqrt(g(pre + '3.normalizer.running_var') + BN_EPS)
        t2 = g(pre + '3.normalizer.bias') - g(pre + '3.normalizer.running_mean') * s2
        parts += [s2, t2]
    wn = g('project_node_embeddings.weight')                 # [384][128]
    wout = g('project_out.weight')                           # [128][128]  glimpse = W_out o
    wsc = g('project_step_context.weight')                   # [128][130]
    lk_fold = _matmul_f64(wout.t(), wn[2 * D:3 * D]) / math.sqrt(D)    # [c][e]
    # K' carries 1/sqrt(16) and log2(e): the glimpse softmax runs in base 2 (exp2) in the kernel
    wdec = torch.cat((wn[0:D].t() * (math.log2(math.e) / math.sqrt(D // H)), wn[D:2 * D].t(), lk_fold.t(), wsc[:, :D].t()), 1)   # [128][512]
    parts += [wdec, g('project_fixed_context.weight').t(), g('fleets_embedding.weight'), wsc[:, D], wsc[:, D + 1]]
    parts += tc_parts + [wdec.t()]
    blob = torch.cat([p.contiguous().reshape(-1) for p in parts]).to(torch.float32).contiguous()
    assert blob.numel() == blob_floats(), (blob.numel(), blob_floats())
    return blob
