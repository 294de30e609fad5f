#!/usr/bin/env python
"""Fused feed-forward kernel (ffn_tc.cu) and tcgen05 attention (mha_tc.cu) against the two-GEMM / mma.sync forms and a float64
evaluation of the encoder, with timings.
    python tools/check_ffn.py [--time] [--mha]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch


def to_task(task_cpu, dev='cuda'):
    from agh_b200.ops import FleetTask
    return FleetTask.from_input({k: v.to(dev) for k, v in task_cpu.items() if k != 'distance'})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--time', action='store_true')
    ap.add_argument('--mha', action='store_true', help='A/B the attention kernels instead of the feed-forward kernels')
    ap.add_argument('--shapes', default='20x7,100x300,50x1000,200x40,300x12')
    args = ap.parse_args()
    from agh_b200 import AttentionModel, AGH, ops, _lib
    from agh_b200.problems.agh.problem_agh import generate_arrays
    from oracle import agh_oracle as O
    torch.manual_seed(1234)
    model = AttentionModel(128, 128, AGH).cuda().eval()
    # non-trivial BatchNorm statistics and biases
    with torch.no_grad():
        for k, v in model.state_dict().items():
            if 'running_mean' in k: v.normal_(0, 0.3)
            if 'running_var' in k: v.uniform_(0.5, 2.0)
    model.invalidate_blob() if hasattr(model, 'invalidate_blob') else None
    blob = model.weight_blob()
    lib = _lib.lib()
    tables = O.Tables.load()
    W64 = {k: v.double() for k, v in model.state_dict().items()}
    for shp in args.shapes.split(','):
        n, B = (int(a) for a in shp.split('x'))
        np.random.seed(3)
        inst = generate_arrays(B, n)
        task_cpu = O.fleet_task(tables, inst, 5, O.new_tw_left_levels(inst))
        task = to_task(task_cpu)
        sw = lib.agh_set_mha_tc if args.mha else lib.agh_set_ffn_fused
        sw(0)
        h_u = ops.encode(blob, task)
        sw(2)
        h_f = ops.encode(blob, task)
        sw(1)
        torch.cuda.synchronize()
        ref = O.encoder(W64, ops.init_embed(blob, task).double())
        eu = (h_u.double() - ref).abs(); ef = (h_f.double() - ref).abs()
        print('AGH-%d B=%d: |h| max %.2f; vs float64: old max %.3e mean %.3e | new max %.3e mean %.3e | new vs old max %.3e'
              % (n, B, ref.abs().max().item(), eu.max().item(), eu.mean().item(), ef.max().item(), ef.mean().item(),
                 (h_f - h_u).abs().max().item()), flush=True)
    if args.time and args.mha:
        qkv = torch.randn(3, 10000 * 101, 128, device='cuda')
        for on in (0, 2, 0, 2):
            lib.agh_set_mha_tc(on)
            for _ in range(3): ops.encoder_mha(qkv, 10000, 100)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): ops.encoder_mha(qkv, 10000, 100)
            e1.record(); torch.cuda.synchronize()
            print('attention alone, AGH-100 B=10000, tcgen05=%d: %.1f us' % (on, e0.elapsed_time(e1) * 50), flush=True)
        n, B = 100, 10000
        np.random.seed(3)
        inst = generate_arrays(B, n)
        task = to_task(O.fleet_task(tables, inst, 5, O.new_tw_left_levels(inst)))
        for on in (0, 1, 0, 1):
            lib.agh_set_mha_tc(on)
            for _ in range(3): ops.encode(blob, task)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): ops.encode(blob, task)
            e1.record(); torch.cuda.synchronize()
            print('encode AGH-100 B=10000 tcgen05 attention=%d: %.3f ms' % (on, e0.elapsed_time(e1) / 10), flush=True)
        lib.agh_set_mha_tc(1)
        return
    if args.time:
        x = torch.randn(1010000, 128, device='cuda')
        for fused in (0, 2, 0, 2):
            lib.agh_set_ffn_fused(fused)
            for _ in range(3): ops.encoder_ffn(blob, 1, x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): ops.encoder_ffn(blob, 1, x)
            e1.record(); torch.cuda.synchronize()
            print('FF sublayer alone, M = 1.01 M rows, fused=%d: %.1f us' % (fused, e0.elapsed_time(e1) * 50), flush=True)
        n, B = 100, 10000
        np.random.seed(3)
        inst = generate_arrays(B, n)
        task = to_task(O.fleet_task(tables, inst, 5, O.new_tw_left_levels(inst)))
        for fused in (0, 1, 0, 1):
            lib.agh_set_ffn_fused(fused)
            for _ in range(3): ops.encode(blob, task)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): ops.encode(blob, task)
            e1.record(); torch.cuda.synchronize()
            print('encode AGH-100 B=10000 fused=%d: %.3f ms' % (fused, e0.elapsed_time(e1) / 10), flush=True)
        lib.agh_set_ffn_fused(1)


if __name__ == '__main__':
    main()
