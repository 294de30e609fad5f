#!/usr/bin/env python
"""Micro-benchmark of agh_encode + agh_precompute (CUDA events): tensor-core vs CUDA-core GEMM back end."""
import argparse, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--flights', type=int, default=100)
    ap.add_argument('--batch', type=int, default=10000)
    ap.add_argument('--iters', type=int, default=3)
    args = ap.parse_args()
    from agh_b200 import AttentionModel, AGH, ops, _lib
    from agh_b200.problems.agh.problem_agh import generate_arrays
    from oracle import agh_oracle as O
    torch.manual_seed(1234)
    model = AttentionModel(128, 128, AGH).cuda().eval()
    blob = model.weight_blob()
    tb = O.Tables.load()
    np.random.seed(4321)
    inst = generate_arrays(args.batch, args.flights)
    task = ops.FleetTask.from_input({k: v.cuda() for k, v in O.fleet_task(tb, inst, 2, O.new_tw_left_levels(inst)).items() if k != 'distance'})
    F = (1278720 * (args.flights + 1) + 1536 * (args.flights + 1) ** 2 + 32768) * args.batch
    res = {}
    for name, be in (('simt', 0), ('tcgen05', 1)):
        _lib.lib().agh_set_gemm_backend(be)
        for _ in range(2):
            h = ops.encode(blob, task); k = ops.precompute(blob, h, task)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        per = [torch.cuda.Event(enable_timing=True) for _ in range(args.iters + 1)]
        e[0].record()
        per[0].record()
        for i in range(args.iters):
            h = ops.encode(blob, task)
            per[i + 1].record()
        e[1].record()
        for _ in range(args.iters):
            k = ops.precompute(blob, h, task)
        e[2].record()
        torch.cuda.synchronize()
        t_enc, t_pre = e[0].elapsed_time(e[1]) / args.iters, e[1].elapsed_time(e[2]) / args.iters
        res[name] = (h, k[0])
        print('         per-iteration encode ms: ' + ' '.join('%.2f' % per[i].elapsed_time(per[i + 1]) for i in range(args.iters)))
        print('%-8s N=%d B=%d: encode %.2f ms, precompute %.2f ms  -> %.1f TFLOP/s (F_enc)' % (name, args.flights, args.batch, t_enc, t_pre, F / (t_enc + t_pre) / 1e9))
    _lib.lib().agh_set_gemm_backend(1)
    # both back ends against a float64 evaluation of the same layers (first 256 instances)
    nb = min(256, args.batch)
    W64 = {k: v.double() for k, v in model.state_dict().items()}
    h0 = ops.init_embed(blob, task)[:nb].double()
    ref = O.encoder(W64, h0)
    for name in res:
        e = (res[name][0][:nb].double() - ref).abs()
        print('%-8s vs float64 layers: max abs err %.3e, mean abs err %.3e' % (name, e.max().item(), e.mean().item()))
    d =(res['simt'][0] - res['tcgen05'][0]).abs()
    print('h: max abs diff %.3e, mean abs diff %.3e, max |h| %.2f; keys max abs diff %.3e' % (d.max().item(), d.mean().item(), res['simt'][0].abs().max().item(), (res['simt'][1] - res['tcgen05'][1]).abs().max().item()))


if __name__ == '__main__':
    main()
