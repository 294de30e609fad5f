"""Training driver: `python -m agh_b200.run --graph_size 100 ...` (reference: run.py:22-185).  Single process per
GPU; launch with torchrun for multi-GPU (instances of the rollouts and of every training batch are sharded over the
ranks, gradients averaged through one flat all-reduce; every rank keeps identical weights)."""
from __future__ import annotations

import json
import os
import pprint

import numpy as np
import torch
import torch.optim as optim

from . import parallel
from .nets.attention_model import AttentionModel
from .options import get_options
from .reinforce_baselines import RolloutBaseline
from .train import train_epoch, validate
from .utils import load_problem, torch_load_cpu


def run(opts):
    if not opts.use_cuda:
        raise RuntimeError('agh_b200 has no CPU path: a CUDA device is required')
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))      # the C ABI launches on the current device
    if 'RANK' in os.environ and not torch.distributed.is_initialized():
        torch.distributed.init_process_group('nccl')
    rank, _ = parallel.world()
    opts.device = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
    if rank == 0:
        pprint.pprint(vars(opts))
    torch.manual_seed(opts.seed)          # identical on every rank: same init, same generated datasets
    np.random.seed(opts.seed)
    if rank == 0:
        os.makedirs(opts.save_dir, exist_ok=True)
        with open(os.path.join(opts.save_dir, 'args.json'), 'w') as f:
            json.dump({k: v for k, v in vars(opts).items() if k != 'device'}, f, indent=True)

    problem = load_problem(opts.problem)
    load_data = {}
    load_path = opts.load_path if opts.load_path is not None else opts.resume
    if load_path is not None:
        print('  [*] Loading data from {}'.format(load_path))
        load_data = torch_load_cpu(load_path)

    model = AttentionModel(opts.embedding_dim, opts.hidden_dim, problem, n_encode_layers=opts.n_encode_layers,
                           mask_inner=True, mask_logits=True, normalization=opts.normalization,
                           tanh_clipping=opts.tanh_clipping, checkpoint_encoder=opts.checkpoint_encoder,
                           shrink_size=opts.shrink_size, wo_time=opts.wo_time, rnn_time=opts.rnn_time).to(opts.device)
    model.load_state_dict({**model.state_dict(), **load_data.get('model', {})})

    baseline = RolloutBaseline(model, problem, opts)
    if 'baseline' in load_data:
        baseline.load_state_dict(load_data['baseline'])

    optimizer = optim.Adam([{'params': model.parameters(), 'lr': opts.lr_model}])
    if 'optimizer' in load_data:
        optimizer.load_state_dict(load_data['optimizer'])
        for state in optimizer.state.values():
            for k, v in state.items():
                if torch.is_tensor(v):
                    state[k] = v.to(opts.device)
    lr_scheduler = optim.lr_scheduler.LambdaLR(optimizer, lambda epoch: opts.lr_decay ** epoch)

    val_dataset = problem.make_dataset(size=opts.graph_size, num_samples=opts.val_size, filename=opts.val_dataset,
                                       distribution=opts.data_distribution)
    if opts.resume:
        epoch_resume = int(os.path.splitext(os.path.split(opts.resume)[-1])[0].split('-')[1])
        torch.set_rng_state(load_data['rng_state'])
        baseline.epoch_callback(model, epoch_resume)
        print('Resuming after {}'.format(epoch_resume))
        opts.epoch_start = epoch_resume + 1

    if opts.eval_only:
        validate(model, val_dataset, opts)
    else:
        for epoch in range(opts.epoch_start, opts.epoch_start + opts.n_epochs):
            train_epoch(model, optimizer, baseline, lr_scheduler, epoch, val_dataset, problem, None, opts)


if __name__ == '__main__':
    run(get_options())
