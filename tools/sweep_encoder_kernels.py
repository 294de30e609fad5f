#!/usr/bin/env python
"""Per-launch times of the two forms of the encoder attention (tcgen05 / mma.sync) and feed-forward sublayer (fused / two GEMMs)
over graph sizes and batch sizes: the data behind the size thresholds in csrc/encoder.cu.
    python tools/sweep_encoder_kernels.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch


def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    from agh_b200 import AttentionModel, AGH, ops, _lib
    torch.manual_seed(1234)
    model = AttentionModel(128, 128, AGH).cuda().eval()
    blob = model.weight_blob()
    lib = _lib.lib()
    for N in (20, 50, 100):
        for B in (64, 256, 1000, 2500, 10000):
            n = N + 1
            qkv = torch.randn(3, B * n, 128, device='cuda')
            x = torch.randn(B * n, 128, device='cuda')
            r = {}
            for on in (0, 1):
                lib.agh_set_mha_tc(2 * on)
                r['mha%d' % on] = timeit(lambda: ops.encoder_mha(qkv, B, N))
                lib.agh_set_ffn_fused(2 * on)
                r['ffn%d' % on] = timeit(lambda: ops.encoder_ffn(blob, 1, x))
            print('N=%3d B=%5d (M=%7d): attention mma.sync %7.1f us, tcgen05 %7.1f us | feed-forward two GEMMs %7.1f us, fused %7.1f us'
                  % (N, B, B * n, r['mha0'], r['mha1'], r['ffn0'], r['ffn1']), flush=True)
    lib.agh_set_mha_tc(1); lib.agh_set_ffn_fused(1)


if __name__ == '__main__':
    main()
