#!/usr/bin/env python
"""A/B of train.LEVEL_BATCH_LIMIT at the headline workload (AGH-100, 10,000 instances, end-to-end rollout)."""
import os, sys, types
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from agh_b200 import AttentionModel, AGH, AGHDataset, train
from agh_b200.problems.agh.problem_agh import generate_arrays

dev = torch.device('cuda')
torch.manual_seed(1234)
model = AttentionModel(128, 128, AGH).to(dev).eval()
model.set_decode_type('greedy')
np.random.seed(4321)
ds = AGHDataset.__new__(AGHDataset)
torch.utils.data.Dataset.__init__(ds)
ds.arrays = {k: v.pin_memory() for k, v in generate_arrays(10000, 100).items()}
ds.size = 10000
opts = types.SimpleNamespace(eval_batch_size=10000, device=dev, no_progress_bar=True)
for limit in (16384, 20000, 40000, 16384, 40000):
    train.LEVEL_BATCH_LIMIT = limit
    for _ in range(2):
        c = train.rollout(model, ds, opts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        c = train.rollout(model, ds, opts)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print('LEVEL_BATCH_LIMIT=%d: %.1f ms per rollout, %.0f inst/s, avg cost %.3f, peak mem %.1f GB' % (limit, ms, 1e7 / ms, c.sum(1).mean().item(), torch.cuda.max_memory_allocated() / 2**30))
