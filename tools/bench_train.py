#!/usr/bin/env python
"""Config C4: AGH-100 REINFORCE training with rollout baseline.  Times the three parts of an epoch through the
public API (agh_b200.train / reinforce_baselines): the baseline's greedy pass over the epoch's instances
(wrap_dataset), the sampling train batches (train_batch_agh) and validation.  Single GPU, or torchrun for N GPUs."""
import argparse
import json
import os
import sys
import time
import types

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--flights', type=int, default=100)
    ap.add_argument('--epoch-size', type=int, default=1280)
    ap.add_argument('--batch-size', type=int, default=64)
    ap.add_argument('--val-size', type=int, default=1000)
    args = ap.parse_args()
    import torch.distributed as dist
    from agh_b200 import AttentionModel, AGH, parallel
    from agh_b200.reinforce_baselines import RolloutBaseline
    from agh_b200.train import train_batch_agh, validate, iter_batches
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if 'RANK' in os.environ:
        dist.init_process_group('nccl', device_id=dev)
    rank, ws = parallel.world()
    torch.manual_seed(1234)
    np.random.seed(1234)
    model = AttentionModel(128, 128, AGH).to(dev)
    opts = types.SimpleNamespace(eval_batch_size=1000, device=dev, no_progress_bar=True, max_grad_norm=1.0, log_step=1e9,
                                 no_tensorboard=True, baseline='rollout', graph_size=args.flights, data_distribution=None,
                                 val_size=args.val_size, bl_alpha=0.05, epoch_size=args.epoch_size, batch_size=args.batch_size)
    baseline = RolloutBaseline(model, AGH, opts)
    optimizer = torch.optim.Adam([{'params': model.parameters(), 'lr': 1e-4}])
    val = AGH.make_dataset(size=args.flights, num_samples=args.val_size)

    def sync():
        torch.cuda.synchronize()
        if ws > 1:
            dist.barrier()

    sync(); t0 = time.perf_counter()
    training = baseline.wrap_dataset(AGH.make_dataset(size=args.flights, num_samples=args.epoch_size))
    sync(); t1 = time.perf_counter()
    model.train()
    nb = 0
    per_batch = []
    for bid, batch in enumerate(iter_batches(training, args.batch_size)):
        if ws > 1:
            lo, hi = parallel.shard_bounds(batch['baseline'].size(0), rank, ws)
            batch = {'data': {k: v[lo:hi] for k, v in batch['data'].items()}, 'baseline': batch['baseline'][lo:hi]}
        tb = time.perf_counter()
        train_batch_agh(model, optimizer, baseline, 0, bid, bid + 1, batch, None, opts)
        sync()
        per_batch.append(time.perf_counter() - tb)
        nb += 1
    sync(); t2 = time.perf_counter()
    steady = sorted(per_batch[2:])[len(per_batch[2:]) // 2] if len(per_batch) > 4 else (t2 - t1) / nb     # median after warm-up
    validate(model, val, opts)
    sync(); t3 = time.perf_counter()
    if rank == 0:
        print(json.dumps({'config': 'C4 AGH-%d REINFORCE' % args.flights, 'n_gpus': ws, 'epoch_size': args.epoch_size,
                          'batch_size': args.batch_size, 'baseline_wrap_s': t1 - t0, 'train_s': t2 - t1,
                          'ms_per_train_batch': 1e3 * steady, 'ms_per_train_batch_incl_warmup': 1e3 * (t2 - t1) / nb,
                          'validate_s': t3 - t2,
                          'projected_epoch_12800_s': (t1 - t0) * 12800 / args.epoch_size + steady * 12800 / args.batch_size + (t3 - t2)}))
    if ws > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
