#!/usr/bin/env python
"""Timings of the other BASELINE.json configs (parity-test cases, not the headline bench line):
C1 AGH-20 greedy B=1000, C2 AGH-50 greedy B=1000 and sampling width 1280, C5 AGH-200/300 greedy.
Prints one JSON line per config; CUDA-event timed through the public API (agh_b200.eval._eval_dataset)."""
import json
import os
import sys
import types

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch


def main():
    from agh_b200 import AttentionModel, AGH, AGHDataset
    from agh_b200.eval import _eval_dataset
    from agh_b200.problems.agh.problem_agh import generate_arrays
    dev = torch.device('cuda')
    torch.manual_seed(1234)
    model = AttentionModel(128, 128, AGH).to(dev).eval()

    def dataset(n, size):
        np.random.seed(4321)
        ds = AGHDataset.__new__(AGHDataset)
        torch.utils.data.Dataset.__init__(ds)
        ds.arrays = {k: v.pin_memory() for k, v in generate_arrays(size, n).items()}
        ds.size = size
        return ds

    def run(name, n, size, strategy, width, bs, max_calc, reps=3):
        ds = dataset(n, size)
        opts = types.SimpleNamespace(decode_strategy=strategy, eval_batch_size=bs, max_calc_batch_size=max_calc,
                                     no_progress_bar=True, problem='agh', compress_mask=False)
        torch.manual_seed(1)
        res = _eval_dataset(model, ds, width, 1.0, opts, dev)           # warm-up
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            res = _eval_dataset(model, ds, width, 1.0, opts, dev)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(json.dumps({'config': name, 'flights': n, 'instances': size, 'strategy': strategy, 'width': width,
                          'ms': ms, 'instances_per_sec': size / ms * 1e3,
                          'rollouts_per_sec': size * max(width, 1) * 10 / ms * 1e3, 'avg_cost': float(res.mean())}))

    run('C1 AGH-20 greedy B=1000', 20, 1000, 'greedy', 0, 1000, 10000)
    run('C2 AGH-50 greedy B=1000', 50, 1000, 'greedy', 0, 1000, 10000)
    run('C2 AGH-50 sampling width 1280 (eval_batch_size 1)', 50, 8, 'sample', 1280, 1, 1280, reps=1)
    run('C2 AGH-50 sampling width 1280 (eval_batch_size 8)', 50, 64, 'sample', 1280, 8, 10240, reps=1)
    run('C5 AGH-200 greedy B=2000', 200, 2000, 'greedy', 0, 2000, 10000, reps=1)
    run('C5 AGH-300 greedy B=1000', 300, 1000, 'greedy', 0, 1000, 10000, reps=1)


if __name__ == '__main__':
    main()
