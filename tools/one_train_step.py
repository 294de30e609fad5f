#!/usr/bin/env python
"""One REINFORCE training batch (AGH-100, 64 instances) through train.train_batch_agh, for launch lists:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python tools/one_train_step.py
Every kernel of the batch must be the library's own (agh::*) or an element-wise / copy / fill / index ATen kernel:
no cuBLAS (gemm / gemv / cutlass), no SDPA / flash attention, no cuDNN."""
import os
import sys
import types

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch


def main():
    from agh_b200 import AttentionModel, AGH, _lib
    from agh_b200.train import train_batch_agh, rollout
    dev = torch.device('cuda')
    torch.manual_seed(1234); np.random.seed(1234)
    model = AttentionModel(128, 128, AGH).to(dev)
    opt = torch.optim.Adam([{'params': model.parameters(), 'lr': 1e-4}])
    ds = AGH.make_dataset(size=100, num_samples=64)
    opts = types.SimpleNamespace(eval_batch_size=64, device=dev, no_progress_bar=True, max_grad_norm=1.0, log_step=1e9,
                                 no_tensorboard=True, baseline='rollout')
    bl = rollout(model, ds, opts)
    batch = {'data': ds.arrays, 'baseline': bl}
    baseline = types.SimpleNamespace(unwrap_batch=lambda b: (b['data'], b['baseline']))
    model.train()
    steps = int(os.environ.get('AGH_TRAIN_STEPS', 2))
    l0 = _lib.lib().agh_launch_count()
    for i in range(steps):
        train_batch_agh(model, opt, baseline, 0, i, i + 1, batch, None, opts)
    torch.cuda.synchronize()
    print('train steps', steps, 'library launches per step', (_lib.lib().agh_launch_count() - l0) // steps)


if __name__ == '__main__':
    main()
