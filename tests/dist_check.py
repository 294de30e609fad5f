#!/usr/bin/env python
"""Multi-GPU check, run under torchrun on 2+ GPUs (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py

1. rollout(): instances sharded over the ranks + NCCL all-gather gives exactly the single-rank cost table (C2).
2. train_batch_agh(): each rank trains on its slice, gradients are averaged through one flat all-reduce (C1);
   the parameters stay identical on all ranks after the optimizer step.
3. sampling with a fixed Philox seed and global instance ids is independent of the sharding.
"""
import os
import sys
import types

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch
import torch.distributed as dist


def main():
    from agh_b200 import AttentionModel, AGH, parallel
    from agh_b200.train import rollout, train_batch_agh
    from agh_b200.reinforce_baselines import BaselineDataset
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    rank, ws = dist.get_rank(), dist.get_world_size()
    torch.manual_seed(1234)
    model = AttentionModel(128, 128, AGH).to(dev)
    np.random.seed(4321)
    ds = AGH.make_dataset(size=50, num_samples=301)                 # odd size: uneven shards
    opts = types.SimpleNamespace(eval_batch_size=64, device=dev, no_progress_bar=True, max_grad_norm=1.0, log_step=1e9,
                                 no_tensorboard=True)
    sharded = rollout(model, ds, opts)
    opts.shard_across_ranks = False
    single = rollout(model, ds, opts)
    opts.shard_across_ranks = True
    assert sharded.shape == (301, 10)
    assert torch.equal(sharded, single), 'sharded rollout differs from the single-rank rollout'

    # training step on a per-rank slice of one global batch
    B = 8 * ws
    lo, hi = parallel.shard_bounds(B, rank, ws)
    batch = {'data': {k: v[lo:hi] for k, v in ds.arrays.items()}, 'baseline': single[lo:hi]}
    baseline = types.SimpleNamespace(unwrap_batch=lambda b: (b['data'], b['baseline']))
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    model.train()
    model.sampling_seed, model.id_offset = 99, lo
    train_batch_agh(model, opt, baseline, 0, 0, 1, batch, None, opts)
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    ref = flat.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(flat, ref), 'parameters diverged across ranks after the all-reduced step'
    bufs = torch.cat([b.detach().float().reshape(-1) for b in model.buffers()])
    ref = bufs.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(bufs, ref), 'BatchNorm running statistics diverged across ranks (they are averaged with the gradients)'

    # sharding-independent sampling: rank r samples its slice with global ids; compare with rank 0 doing all of it
    model.eval()
    model.set_decode_type('sampling')
    from agh_b200 import ops
    info = model.fleet_info
    x = {k: v[:B].to(dev) for k, v in ds.arrays.items()}
    lv = x['arrival'].repeat(6, 1, 1).contiguous()
    full = ops.fleet_task(x['loc'].contiguous(), x['type'].contiguous(), x['departure'].contiguous(), x['demand'].contiguous(),
                          lv, 0, 1, info['duration'][1], info['next_duration'][0], False)
    model.id_offset = 0
    with torch.no_grad():
        model.sampling_seed = 99            # (re)setting the seed restarts its per-rollout counter
        c_full = model(full)[0]
        xs = {k: v[lo:hi].contiguous() for k, v in x.items()}
        part = ops.fleet_task(xs['loc'], xs['type'], xs['departure'], xs['demand'], xs['arrival'].repeat(6, 1, 1).contiguous(),
                              0, 1, info['duration'][1], info['next_duration'][0], False)
        model.id_offset = lo
        model.sampling_seed = 99
        c_part = model(part)[0]
    assert torch.equal(c_part, c_full[lo:hi]), 'sampled tours depend on the sharding'
    dist.barrier()
    if rank == 0:
        print('dist_check OK: world_size=%d' % ws)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
