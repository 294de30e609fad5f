#!/usr/bin/env python
"""Micro-benchmark of the fused decode kernel alone (CUDA events), for kernel tuning.

    python tools/bench_decode.py --flights 100 --batch 2000 --smem-mats 4 2
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--flights', type=int, nargs='+', default=[100])
    ap.add_argument('--batch', type=int, default=2000)
    ap.add_argument('--smem-mats', type=int, nargs='+', default=[0])
    ap.add_argument('--fleet', type=int, default=2)
    ap.add_argument('--iters', type=int, default=5)
    ap.add_argument('--sampling', action='store_true')
    ap.add_argument('--reps', type=int, default=1, help='replicas per instance sharing one key set (sample_many)')
    ap.add_argument('--no-shared', action='store_true', help='force one CTA per replica')
    args = ap.parse_args()
    from agh_b200 import AttentionModel, AGH, ops, _lib
    from agh_b200.problems.agh.problem_agh import generate_arrays
    from oracle import agh_oracle as O
    torch.manual_seed(1234)
    model = AttentionModel(128, 128, AGH).cuda().eval()
    blob, dist = model.weight_blob(), model.distance_table()
    tb = O.Tables.load()
    for N in args.flights:
        np.random.seed(4321)
        inst = generate_arrays(args.batch, N)
        task = ops.FleetTask.from_input({k: v.cuda() for k, v in O.fleet_task(tb, inst, args.fleet, O.new_tw_left_levels(inst)).items()
                                         if k != 'distance'})
        h = ops.encode(blob, task)
        keys, psc, qbase = ops.precompute(blob, h, task)
        for sm in args.smem_mats:
            mode = _lib.AGH_SAMPLING if args.sampling else _lib.AGH_GREEDY
            for _ in range(2):
                out = ops.decode(blob, dist, task, keys, psc, qbase, mode=mode, smem_mats=sm, seed=1, reps=args.reps,
                                 shared_keys=not args.no_shared)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                out = ops.decode(blob, dist, task, keys, psc, qbase, mode=mode, smem_mats=sm, seed=1, reps=args.reps,
                                 shared_keys=not args.no_shared)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            steps = float(out['steps'].sum())
            print('N=%d B=%d reps=%d%s smem_mats=%d: %.3f ms  %.1f M instance-steps/s  %.2f M rollouts/s  cost %.2f'
                  % (N, args.batch, args.reps, ' (per-replica CTAs)' if args.no_shared else '', sm, ms, steps / ms / 1e3,
                     args.batch * args.reps / ms / 1e3, out['cost'].mean().item()))


if __name__ == '__main__':
    main()
