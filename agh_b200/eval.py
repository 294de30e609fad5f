"""Evaluation driver: the reference's eval.py (_eval_dataset :80-154, eval_dataset :52-77, CLI :157-192).

Greedy: one rollout per instance and fleet.  Sampling: best of `width` rollouts per instance and fleet
(greedy over fleets, eval.py:130-131); all `width` replicas of an instance share one staged key set on
the device.  Multi-GPU: launch with torchrun; instances are sharded over ranks (parallel.py) instead of
the reference's mp.Pool path.
"""
from __future__ import annotations

import argparse
import math
import os
import random
import time

import numpy as np
import torch

from . import ops, parallel
from .train import iter_batches, get_inner_model, fleet_rollout
from .utils import load_model, move_to
from .utils.functions import parse_softmax_temperature


def _eval_dataset(model, dataset, width, softmax_temp, opts, device):
    """Returns the total cost per instance, [len(dataset)] CPU f32."""
    model.to(device)
    model.eval()
    model.set_decode_type('greedy' if opts.decode_strategy in ('bs', 'greedy') else 'sampling', temp=softmax_temp)
    assert opts.decode_strategy in ('sample', 'greedy'), 'beam search is not supported for AGH (eval.py:140)'
    info = model.fleet_info
    rank, ws = parallel.world()
    lo, hi = parallel.shard_bounds(len(dataset), rank, ws)
    defer, model.defer_checks = model.defer_checks, True
    id_offset0 = model.id_offset
    cost = []
    try:
        done = 0
        for batch in iter_batches(dataset, opts.eval_batch_size, lo, hi, getattr(opts, 'no_progress_bar', True)):
            x = move_to(batch, device)
            B = x['loc'].size(0)
            model.id_offset = lo + done          # Philox counters carry the global instance index (shard-independent)
            done += B
            if opts.decode_strategy == 'greedy':
                assert width == 0, 'Do not set width when using greedy'
                assert opts.eval_batch_size <= opts.max_calc_batch_size, \
                    'eval_batch_size should be smaller than calc batch size'
                with torch.no_grad():                      # fleets of one precedence level share launches
                    costs, _ = fleet_rollout(model, batch, device)
                model.raise_pending()
                cost.append(costs)
                continue
            elif width * opts.eval_batch_size > opts.max_calc_batch_size:
                assert opts.eval_batch_size == 1
                assert width % opts.max_calc_batch_size == 0
                batch_rep, iter_rep = opts.max_calc_batch_size, width // opts.max_calc_batch_size
            else:
                batch_rep, iter_rep = width, 1
            assert batch_rep > 0
            loc, type_ = x['loc'].contiguous(), x['type'].contiguous()
            departure, demand_all = x['departure'].contiguous(), x['demand'].contiguous()
            lv = x['arrival'].repeat(len(info['next_duration']) + 1, 1, 1).contiguous()
            batch_cost = []
            with torch.no_grad():
                for f in info['order']:
                    p = info['precedence'][f]
                    nd = info['next_duration'][p]
                    task = ops.fleet_task(loc, type_, departure, demand_all, lv, p, f, info['duration'][f], nd,
                                          nd_is_f32=all(type(v) is float for v in nd))
                    _, costs, serve = model.sample_many(task, batch_rep=batch_rep, iter_rep=iter_rep)
                    batch_cost.append(costs.view(-1, 1))
                    ops.chain_tw_left(lv, p, serve.contiguous())
            model.raise_pending()
            cost.append(torch.cat(batch_cost, 1))
    finally:
        model.defer_checks, model.id_offset = defer, id_offset0
    local = torch.cat(cost, 0) if cost else torch.zeros(0, 10, device=device)
    if ws > 1:
        local = parallel.all_gather_rows(local, len(dataset))
    return local.sum(1).cpu()


def eval_dataset(dataset_path, width, softmax_temp, opts):
    start_time = time.time()
    model, _ = load_model(opts.model)
    assert model.is_agh
    if not torch.cuda.is_available() or opts.no_cuda:
        raise RuntimeError('agh_b200 has no CPU path: a CUDA device is required (reference flag --no_cuda is rejected)')
    rank, ws = parallel.world()
    device = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
    dataset = model.problem.make_dataset(filename=dataset_path, num_samples=opts.val_size, offset=opts.offset)
    results = _eval_dataset(model, dataset, width, softmax_temp, opts, device)
    if rank == 0:
        print(results.tolist())
        print('Using {} strategy: Average cost: {} +- {}'.format(opts.decode_strategy, torch.mean(results),
                                                                 torch.std(results) / math.sqrt(len(results))))
        print('>> End of validation within {:.2f}s'.format(time.time() - start_time))
    return results


def get_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('--datasets', nargs='+', default=['./data/agh/agh20_validation_seed4321.pkl'])
    parser.add_argument('--seed', type=int, default=1234)
    parser.add_argument('--val_size', type=int, default=1000)
    parser.add_argument('--offset', type=int, default=0)
    parser.add_argument('--eval_batch_size', type=int, default=1000)
    parser.add_argument('--width', type=int, nargs='+')
    parser.add_argument('--decode_strategy', type=str, default='greedy', choices=['sample', 'greedy', 'bs'])
    parser.add_argument('--softmax_temperature', type=parse_softmax_temperature, default=1)
    parser.add_argument('--model', type=str, default='./data/20')
    parser.add_argument('--no_cuda', action='store_true')
    parser.add_argument('--no_progress_bar', action='store_true')
    parser.add_argument('--compress_mask', action='store_true')
    parser.add_argument('--max_calc_batch_size', type=int, default=1000)
    parser.add_argument('--results_dir', default='results')
    parser.add_argument('--multiprocessing', action='store_true',
                        help='accepted for compatibility; use torchrun for multi-GPU')
    return parser


def main(argv=None):
    opts = get_parser().parse_args(argv)
    if torch.cuda.is_available():          # the C ABI launches on the current device: one rank = one GPU
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    if 'RANK' in os.environ and not torch.distributed.is_initialized():
        torch.distributed.init_process_group('nccl')
    random.seed(opts.seed)
    np.random.seed(opts.seed)
    torch.manual_seed(opts.seed)
    torch.cuda.manual_seed_all(opts.seed)
    for width in (opts.width if opts.width is not None else [0]):
        for dataset_path in opts.datasets:
            eval_dataset(dataset_path, width, opts.softmax_temperature, opts)


if __name__ == '__main__':
    main()
