"""REINFORCE forward / backward of ONE fleet sub-problem on the CUDA path (train.py:159-219 differentiates
attention_model.py:159-196 -> graph_encoder.py:56-172 and attention_model.py:392-451,510-584).

Forward  = init embedding -> 3 encoder layers with batch-statistics BatchNorm (running statistics updated in place,
           graph_encoder.py:137-138) -> decoder projection -> the SAMPLING ROLLOUT itself (agh_decode), which returns the
           log-likelihood of the sampled tour and records, per step, the query, the glimpse attention weights, the glimpse
           and the log-probs.
Backward = hand-derived gradients of that log-likelihood: element-wise kernels + batched GEMMs over the recorded per-step
           tensors for the decoder, then the encoder layers in reverse (BatchNorm, feed-forward, self-attention), the
           init embedding, and the algebraic un-folding of the packed decoder weights (nets/weights.py) back to the
           reference's parameters.
Every contraction is an agh_gemm / agh_mha_* / agh_bn_* call into libagh_b200 (csrc/train.cu): no cuBLAS, no SDPA, no
ATen matmul.  torch supplies memory, views, a few element-wise / index_add ops and the autograd hook-up.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch

from .. import _lib, ops
from ..ops import FleetTask
from .weights import offsets, D, FF, LAYERS

H, HD = 8, 16
CK = math.log2(math.e) / math.sqrt(HD)       # scale folded into K' (nets/weights.py)


def _splits(K: int) -> int:
    return max(1, min(32, K // 256))


_OFF = None


def _layer_views(blob: torch.Tensor, l: int):
    global _OFF
    o = _OFF = _OFF or offsets()
    base = o['layer0'] + l * o['layer_size']
    v = lambda k, n: blob[base + o['l_' + k]: base + o['l_' + k] + n]
    tb = o['tc0'] + l * o['t_size']
    t = lambda k, n: blob[tb + o[k]: tb + o[k] + n]
    # every weight in both layouts, [in][out] and [out][in]: an activation-sized product C = A W is either a plain row-major
    # GEMM against the [K][N] copy (the register-tiled CUDA-core SGEMM) or, from TC_MIN_ROWS rows, a product against the K-major
    # [N][K] copy with transB, which agh_gemm runs on the tcgen05 kernel (_act_gemm below)
    return {'wqkv': v('wqkv', D * 3 * D), 'wo': v('wo', D * D), 'w1t': v('w1t', D * FF), 'b1': v('b1', FF),
            'w2t': v('w2t', FF * D), 'b2': v('b2', D),
            'wqkv_T': t('t_wqkv', 3 * D * D), 'wo_T': t('t_wo', D * D), 'w1': t('t_w1', FF * D), 'w2': t('t_w2', D * FF)}


TC_MIN_ROWS = 4096       # agh_gemm's threshold for the tensor-core path (AGH_TC_TRAIN_MIN_ROWS in csrc/train.cu)


def _act_gemm(A, w_kn, w_nk, C, M, N, K, lda, ldc, accumulate=False, mask=None, ldm=0, **kw):
    """C[M,N] = A[M,K] W (+ bias, ReLU, mask, accumulate): w_kn is W as [K][N], w_nk the same matrix as [N][K].  From TC_MIN_ROWS
    rows the product runs on the tcgen05 kernel (bias / ReLU fused); masking and accumulation are then one element-wise pass."""
    if M >= TC_MIN_ROWS and ldc == N:
        out = torch.empty_like(C) if accumulate else C
        ops.gemm(A, w_nk, out, M, N, K, lda, K, ldc, transB=True, **kw)
        if mask is not None:
            out.view(M, N).mul_(mask.view(M, N) > 0)
        if accumulate:
            C.add_(out)
        return C
    return ops.gemm(A, w_kn, C, M, N, K, lda, N, ldc, accumulate=accumulate, mask=mask, ldm=ldm, **kw)


def encoder_train_forward(model, blob, task: FleetTask, P, bufs, save: bool):
    """Init embedding + 3 encoder layers with batch-statistics BatchNorm (graph_encoder.py:137-138,146-172); updates the
    running statistics in place.  Returns (h [B*n,128], per-layer tensors for the backward when `save`)."""
    dev = blob.device
    B, N = task.B, task.N
    M = B * (N + 1)
    scratch = torch.empty(2 * D, dtype=torch.float64, device=dev)
    x = ops.init_embed(blob, task).view(M, D)
    saved = []
    for l in range(LAYERS):
        w = _layer_views(blob, l)
        pre = 'embedder.layers.%d.' % l
        qkv = torch.empty(M, 3 * D, dtype=torch.float32, device=dev)
        _act_gemm(x, w['wqkv'], w['wqkv_T'], qkv, M, 3 * D, D, D, 3 * D)
        heads = ops.mha_fwd(qkv, B, N)
        z1 = x.clone()
        _act_gemm(heads, w['wo'], w['wo_T'], z1, M, D, D, D, D, accumulate=True)
        y1, m1, r1 = ops.bn_train_fwd(z1, P[pre + '1.normalizer.weight'], P[pre + '1.normalizer.bias'],
                                      bufs[pre + '1.normalizer.running_mean'], bufs[pre + '1.normalizer.running_var'], scratch)
        hid = torch.empty(M, FF, dtype=torch.float32, device=dev)
        _act_gemm(y1, w['w1t'], w['w1'], hid, M, FF, D, D, FF, bias=w['b1'], relu=True)
        z2 = y1.clone()
        _act_gemm(hid, w['w2t'], w['w2'], z2, M, D, FF, FF, D, bias=w['b2'], accumulate=True)
        y2, m2, r2 = ops.bn_train_fwd(z2, P[pre + '3.normalizer.weight'], P[pre + '3.normalizer.bias'],
                                      bufs[pre + '3.normalizer.running_mean'], bufs[pre + '3.normalizer.running_var'], scratch)
        bufs[pre + '1.normalizer.num_batches_tracked'].add_(1)
        bufs[pre + '3.normalizer.num_batches_tracked'].add_(1)
        if save:
            saved.append((x, qkv, heads, z1, m1, r1, y1, hid, z2, m2, r2))
        x = y2
    model._bn_dirty = True             # the eval-mode BatchNorm fold of the packed weights is stale now
    return x, saved


class FleetTrainFn(torch.autograd.Function):
    """ll = FleetTrainFn.apply(model, task, names, *params); cost / serve / pi / steps ride along as non-differentiable
    outputs.  `params` are the model's parameters in named_parameters() order (so autograd routes the gradients)."""

    @staticmethod
    def forward(ctx, model, task: FleetTask, names: List[str], *params):
        P: Dict[str, torch.Tensor] = dict(zip(names, params))
        bufs = model._buffer_dict()
        B, n = task.B, task.N + 1
        blob = model.weight_blob()
        x, saved = encoder_train_forward(model, blob, task, P, bufs, save=True)
        h = x.view(B, n, D)
        keys, psc, qbase = ops.precompute(blob, h, task)
        out = ops.decode(blob, model.distance_table(), task, keys, psc, qbase, mode=model._mode(), temp=model.temp,
                         seed=model._seed() if model.decode_type == 'sampling' else 0, id_offset=model.id_offset,
                         want_dumps=True, want_logp_full=True, want_train_dumps=True)
        ctx.model, ctx.task, ctx.names = model, task, names
        ctx.saved, ctx.h, ctx.keys, ctx.out, ctx.blob = saved, h, keys, out, blob
        ctx.temp = float(model.temp)
        ctx.mark_non_differentiable(out['cost'], out['serve'], out['pi'], out['steps'], out['status'])
        return out['ll'], out['cost'], out['serve'], out['pi'], out['steps'], out['status']

    @staticmethod
    def backward(ctx, gll, *_unused):
        model, task, names = ctx.model, ctx.task, ctx.names
        P = model._param_dict()
        out, keys, h, blob = ctx.out, ctx.keys, ctx.h, ctx.blob
        dev = h.device
        B, n, _ = h.shape
        N = n - 1
        M = B * n
        T = out['pi'].size(1)
        o = offsets()
        G: Dict[str, torch.Tensor] = {}
        f32 = dict(dtype=torch.float32, device=dev)
        gll = gll.contiguous().to(torch.float32)

        # ---------------- decoder: d ll / d (K', V, LK', P_sc rows, qbase, w_cap, w_time) ----------------
        dU = ops.tf_dlogits(out['logp_full'], out['lse_dump'], out['pi'], out['steps'], gll, ctx.temp, N)     # [B,T,n]
        O, Q, A = out['o_dump'], out['q_dump'], out['attn_dump']
        dkeys = torch.empty(3, B, n, D, **f32)
        Tn, TD, nD = T * n, T * D, n * D
        # dLK'[b] = dU[b]^T O[b]
        ops.gemm(dU, O, dkeys[2], n, D, T, n, D, D, transA=True, batch=(B, 1), sA=(Tn, 0), sB=(TD, 0), sC=(nD, 0))
        # dO[b] = dU[b] LK'[b]
        dO = torch.empty(B, T, D, **f32)
        ops.gemm(dU, keys[2], dO, T, D, n, n, D, D, batch=(B, 1), sA=(Tn, 0), sB=(nD, 0), sC=(TD, 0))
        # dA[b,:,h,:] = dO_h V_h^T
        dA = torch.empty(B, T, H, n, **f32)
        ops.gemm(dO, keys[1], dA, T, n, HD, D, D, H * n, transB=True, batch=(B, H), sA=(TD, HD), sB=(nD, HD), sC=(T * H * n, n))
        ops.tf_dscore_(A, dA, N)                                   # dA -> dS' in place
        # dV_h = A_h^T dO_h
        ops.gemm(A, dO, dkeys[1], n, HD, T, H * n, D, D, transA=True, batch=(B, H), sA=(T * H * n, n), sB=(TD, HD), sC=(nD, HD))
        # dQ_h = dS'_h K'_h
        dQ = torch.empty(B, T, D, **f32)
        ops.gemm(dA, keys[0], dQ, T, HD, n, H * n, D, D, batch=(B, H), sA=(T * H * n, n), sB=(nD, HD), sC=(TD, HD))
        # dK'_h = dS'_h^T Q_h
        ops.gemm(dA, Q, dkeys[0], n, HD, T, H * n, D, D, transA=True, batch=(B, H), sA=(T * H * n, n), sB=(TD, HD), sC=(nD, HD))
        dwcap = torch.zeros(D, **f32)
        dwtime = torch.zeros(D, **f32)
        dqbase, dpsc = ops.tf_dquery(dQ, out['pi'], out['steps'], out['used_dump'], out['cft_dump'], N, dwcap, dwtime)

        # ---------------- decoder projection: [K'|V|LK'|P_sc] = h wdec, qbase = W_fc mean(h) + E_fleet[f] ----------------
        wdec = blob[o['wdec']: o['wdec'] + D * 4 * D]                 # [128][512]
        wdec_T = blob[o['tc_dec']: o['tc_dec'] + 4 * D * D]         # [512][128] = wdec^T
        dwdec = torch.zeros(D, 4 * D, **f32)
        dh = torch.zeros(M, D, **f32)
        hm = h.view(M, D)
        for k, dX in enumerate((dkeys[0], dkeys[1], dkeys[2], dpsc)):
            dXm = dX.view(M, D)
            ops.gemm(hm, dXm, dwdec[:, k * D:], D, D, M, D, D, 4 * D, transA=True, splits=_splits(M))
            ops.gemm(dXm, wdec_T[k * D * D:], dh, M, D, D, D, D, D, accumulate=True)         # dh += dX_k wdec_k^T
        mean = h.mean(1)                                             # [B,128]
        wfc = P['project_fixed_context.weight']
        G['project_fixed_context.weight'] = ops.gemm(dqbase, mean, torch.empty(D, D, **f32), D, D, B, D, D, D, transA=True)
        dmean = ops.gemm(dqbase, wfc, torch.empty(B, D, **f32), B, D, D, D, D, D)
        dh.view(B, n, D).add_(dmean[:, None, :] / n)
        fl = task.fleet_ids.long() if task.fleet_ids is not None else torch.full((B,), task.fleet, device=dev, dtype=torch.long)
        G['fleets_embedding.weight'] = torch.zeros(10, D, **f32).index_add_(0, fl, dqbase)
        # un-fold wdec (nets/weights.py): K' = ck W_k^T, V = W_v^T, LK' = (W_out^T W_lk / sqrt(D))^T, P_sc = W_sc[:, :128]^T
        wn = P['project_node_embeddings.weight']                     # [384][128]
        wout = P['project_out.weight']                               # [128 g][128 c]
        gwn = torch.empty(3 * D, D, **f32)
        gwn[0:D] = dwdec[:, 0:D].t() * CK
        gwn[D:2 * D] = dwdec[:, D:2 * D].t()
        inv = 1.0 / math.sqrt(D)
        blk = dwdec[:, 2 * D:]                                       # [e][c] view, ld = 512
        ops.gemm(wout, blk, gwn[2 * D:], D, D, D, D, 4 * D, D, transB=True, alpha=inv)       # dW_lk[g][e]
        G['project_node_embeddings.weight'] = gwn
        G['project_out.weight'] = ops.gemm(wn[2 * D:], blk, torch.empty(D, D, **f32), D, D, D, D, 4 * D, D, alpha=inv)
        gsc = torch.empty(D, D + 2, **f32)
        gsc[:, :D] = dwdec[:, 3 * D:].t()
        gsc[:, D] = dwcap
        gsc[:, D + 1] = dwtime
        G['project_step_context.weight'] = gsc

        # ---------------- encoder layers in reverse ----------------
        scratch = torch.empty(2 * D, dtype=torch.float64, device=dev)
        dy = dh
        for l in reversed(range(LAYERS)):
            x, qkv, heads, z1, m1, r1, y1, hid, z2, m2, r2 = ctx.saved[l]
            w = _layer_views(blob, l)
            pre = 'embedder.layers.%d.' % l
            dz2, G[pre + '3.normalizer.weight'], G[pre + '3.normalizer.bias'] = ops.bn_train_bwd(
                dy, z2, m2, r2, P[pre + '3.normalizer.weight'], scratch)
            G[pre + '2.module.2.bias'] = ops.colsum(dz2, M, D)
            # Linear2.weight [128][512] = dz2^T hid
            G[pre + '2.module.2.weight'] = ops.gemm(dz2, hid, torch.zeros(D, FF, **f32), D, FF, M, D, FF, FF, transA=True,
                                                    splits=_splits(M))
            dhid = torch.empty(M, FF, **f32)
            _act_gemm(dz2, w['w2'], w['w2t'], dhid, M, FF, D, D, FF, mask=hid, ldm=FF)          # (dz2 W2) o (hid > 0)
            G[pre + '2.module.0.bias'] = ops.colsum(dhid, M, FF)
            G[pre + '2.module.0.weight'] = ops.gemm(dhid, y1, torch.zeros(FF, D, **f32), FF, D, M, FF, D, D, transA=True,
                                                    splits=_splits(M))
            _act_gemm(dhid, w['w1'], w['w1t'], dz2, M, D, FF, FF, D, accumulate=True)            # dy1 = dz2 + dhid W1
            dz1, G[pre + '1.normalizer.weight'], G[pre + '1.normalizer.bias'] = ops.bn_train_bwd(
                dz2, z1, m1, r1, P[pre + '1.normalizer.weight'], scratch)
            G[pre + '0.module.W_out'] = ops.gemm(heads, dz1, torch.zeros(D, D, **f32), D, D, M, D, D, D, transA=True,
                                                 splits=_splits(M)).view(H, HD, D)
            dheads = _act_gemm(dz1, w['wo_T'], w['wo'], torch.empty(M, D, **f32), M, D, D, D, D)
            dqkv = ops.mha_bwd(qkv, heads, dheads, B, N)
            dwqkv = ops.gemm(x, dqkv, torch.zeros(D, 3 * D, **f32), D, 3 * D, M, D, 3 * D, 3 * D, transA=True, splits=_splits(M))
            for i, nm in enumerate(('W_query', 'W_key', 'W_val')):
                G[pre + '0.module.' + nm] = dwqkv[:, i * D:(i + 1) * D].reshape(D, H, HD).permute(1, 0, 2).contiguous()
            _act_gemm(dqkv, w['wqkv_T'], w['wqkv'], dz1, M, D, 3 * D, 3 * D, D, accumulate=True)  # dx = dz1 + dqkv Wqkv^T
            dy = dz1

        # ---------------- init embedding (attention_model.py:272-294) ----------------
        dh0 = dy.view(B, n, D)
        twr = task.tw_right.float() / 1440 if task.twr_f32 else (task.tw_right / 1440).float()
        feat = torch.zeros(B, n, 4, **f32)
        feat[:, 1:, 0] = task.demand
        feat[:, 1:, 1] = task.tw_left[:, 1:] / 1440
        feat[:, 1:, 2] = twr[:, 1:]
        feat[:, 1:, 3] = 1.0
        dwb = ops.gemm(feat.view(M, 4), dy, torch.zeros(4, D, **f32), 4, D, M, 4, D, D, transA=True, splits=_splits(M))
        G['init_embed.weight'] = dwb[:3].t().contiguous()
        G['init_embed.bias'] = dwb[3].contiguous()
        G['loc_embedding.weight'] = torch.zeros(92, D, **f32).index_add_(0, task.coord.view(-1).long(), dy)
        ctx.saved = ctx.out = ctx.keys = ctx.h = ctx.blob = None      # release the recorded tensors with the backward
        return (None, None, None) + tuple(G.get(nm) for nm in names)
