"""Weight-free construction heuristics over the device environment: the reference's
`nearest_neighbor` (agh_baseline.py:80-101), `cws` (Clarke-Wright savings, agh_baseline.py:17-77) and the fleet
loop `val` (agh_baseline.py:290-335), with the same names, arguments and return values.

They are a second, independent consumer of make_state / get_mask / update / get_costs: every step's mask and state
update is one launch of agh_state_get_mask / agh_state_update (state_agh.py through the C ABI), the scores are a
few torch ops on the device, and the finished tours go through the device validator agh_get_costs.  The shipped
agh20 dataset gives two known answers that do not depend on any weights: NN 147099.953125, CWS 106354.4296875.

    python -m agh_b200.agh_baseline --val_method cws --graph_size 20
"""
from __future__ import annotations

import argparse
import math
import time

import numpy as np
import torch

from .nets.attention_model import load_fleet_info
from .problems.agh.problem_agh import AGH
from .train import iter_batches
from .utils import move_to

NODE_SIZE, SPEED = 92, 110.0


def _distance_row(input):
    d = input['distance']
    return d[0] if d.dim() == 2 else d              # every row of the reference's stride-0 [B,8464] view is the table


def nearest_neighbor(input, problem, return_state=False):
    """agh_baseline.py:80-101: go to the closest feasible node (the depot counts when it is allowed); masked nodes
    are pushed to distance 10000, ties go to the lowest index."""
    state = problem.make_state(input)
    table = _distance_row(input)
    sequences = []
    while not state.all_finished():
        mask = state.get_mask()[:, 0, :]
        here = state.coords.gather(1, state.prev_a)                         # [B,1] location id of the vehicle
        distance = table[state.NODE_SIZE * here + state.coords]             # [B,n] (advanced indexing copies)
        distance[mask.bool()] = 10000
        selected = distance.min(1)[1]
        state = state.update(selected)
        sequences.append(selected)
    cost, _ = problem.get_costs(input, torch.stack(sequences, 1))
    return (cost, state) if return_state else (cost, state.serve_time)


def cws(input_, problem, opt):
    """agh_baseline.py:17-77.  savings[i][j] = d(0,j) + d(i,0) - d(i,j) - 1.8 * (tw_left[j] - tw_left[i] -
    duration[i] - d(i,j)/110); start at the flight with the earliest tw_right, then always move to the unmasked node
    with the largest saving from the current one (the depot scores min - 1: a last resort).  Dtypes follow the
    reference: distances and d/110 in f32, the time term and the sum in f64.  At the depot the reference indexes the
    savings with prev_a - 1 = -1, i.e. the row of the LAST flight; kept."""
    loc = input_['loc']
    B, N = loc.shape
    assert N == opt.graph_size
    table = _distance_row(input_)
    d_from_depot = table[loc]                                               # [B,N]  d(0, loc_j)
    d_to_depot = table[NODE_SIZE * loc]                                     # [B,N]  d(loc_i, 0)
    d_ij = table[NODE_SIZE * loc[:, :, None] + loc[:, None, :]]             # [B,N,N] f32
    savings_distance = d_from_depot[:, None, :] + d_to_depot[:, :, None] - d_ij
    tw_left, tw_right = input_['tw_left'][:, 1:], input_['tw_right'][:, 1:]
    # a tensor divisor: torch's CUDA `tensor / python_scalar` multiplies by the reciprocal, one ulp off the
    # correctly rounded quotient the CPU reference computes
    travel = d_ij / torch.full((), SPEED, dtype=d_ij.dtype, device=d_ij.device)
    savings_time = (tw_left[:, None, :] - tw_left[:, :, None]) - input_['duration'][:, :, None] - travel
    savings = savings_distance - 0.03 * 60 * savings_time                   # f64

    state = problem.make_state(input_)
    # equal tw_right values are common (same aircraft type and departure minute) and the reference breaks the tie
    # with torch's (unstable) CPU sort; this one-off [B,N] sort per fleet runs there too, so the known answer holds
    selected = (1 + tw_right.cpu().sort(dim=1)[1][:, 0]).to(loc.device)
    state = state.update(selected)
    sequences = [selected]
    rows = torch.arange(B, device=loc.device)
    neg_inf = torch.tensor(-math.inf, dtype=savings.dtype, device=loc.device)
    while not state.all_finished():
        mask = state.get_mask()[:, 0, :]
        score = savings[rows, state.prev_a[:, 0] - 1]                       # prev_a = 0 wraps to the last row
        score = torch.cat((score.min(1, keepdim=True)[0] - 1, score), 1)
        score = torch.where(mask.bool(), neg_inf, score)
        selected = score.max(1)[1]
        state = state.update(selected)
        sequences.append(selected)
    cost, _ = problem.get_costs(input_, torch.stack(sequences, 1))
    return cost, state.serve_time


def val(dataset, opt, fleet_info, distance, problem):
    """agh_baseline.py:290-335: the ten fleets in order, tw_left chained through the precedence levels.
    Returns the per-instance total cost [len(dataset)] on the CPU and prints the reference's summary line."""
    cost = []
    for bat in iter_batches(dataset, getattr(opt, 'eval_batch_size', 1000), no_progress_bar=getattr(opt, 'no_progress_bar', True)):
        bat = move_to(bat, opt.device)
        B = bat['loc'].size(0)
        bat_cost = []
        bat_tw_left = bat['arrival'].repeat(len(fleet_info['next_duration']) + 1, 1, 1)
        for f in fleet_info['order']:
            p = fleet_info['precedence'][f]
            next_duration = torch.tensor(fleet_info['next_duration'][p], device=opt.device)     # f64, f32 at level 4
            tw_right = bat['departure'] - next_duration[bat['type']]
            tw_right = torch.cat((torch.full_like(tw_right[:, :1], 1441), tw_right), 1)
            tw_left = torch.cat((torch.zeros_like(bat_tw_left[p][:, :1]), bat_tw_left[p]), 1)
            duration = torch.tensor(fleet_info['duration'][f], device=opt.device)
            fleet_bat = {'loc': bat['loc'], 'demand': bat['demand'][:, f - 1, :],
                         'distance': distance.to(opt.device).expand(B, distance.numel()),
                         'duration': duration[bat['type']], 'tw_right': tw_right, 'tw_left': tw_left,
                         'fleet': torch.full((B, 1), f - 1, device=opt.device)}
            if opt.val_method == 'cws':
                fleet_cost, serve_time = cws(fleet_bat, problem, opt)
            elif opt.val_method == 'nearest_neighbor':
                fleet_cost, serve_time = nearest_neighbor(fleet_bat, problem)
            else:
                raise ValueError('unsupported val method %r (the insertion and annealing heuristics of the reference are '
                                 'host-side searches outside the accelerated path)' % opt.val_method)
            bat_cost.append(fleet_cost.view(-1, 1))
            bat_tw_left[p + 1] = torch.max(bat_tw_left[p + 1], serve_time[:, 1:])
        cost.append(torch.cat(bat_cost, 1))
    per_fleet = torch.cat(cost, 0).cpu()
    total = per_fleet.sum(1)
    print('Validation overall avg_cost: {} +- {}'.format(total.mean(), torch.std(total) / math.sqrt(len(total))))
    return total, per_fleet


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--filename', type=str, default=None, help='Dataset pkl to load (default: generate with --seed)')
    parser.add_argument('--problem', type=str, default='agh')
    parser.add_argument('--graph_size', type=int, default=20)
    parser.add_argument('--val_method', type=str, default='cws', choices=['cws', 'nearest_neighbor'])
    parser.add_argument('--val_size', type=int, default=1000)
    parser.add_argument('--offset', type=int, default=0)
    parser.add_argument('--seed', type=int, default=4321)
    parser.add_argument('--eval_batch_size', type=int, default=1000)
    parser.add_argument('--no_progress_bar', action='store_true')
    opts = parser.parse_args(argv)
    if not torch.cuda.is_available():
        raise RuntimeError('agh_b200.agh_baseline runs the environment kernels on a CUDA device; none is visible')
    opts.device = torch.device('cuda')
    np.random.seed(opts.seed)
    torch.manual_seed(opts.seed)
    dataset = AGH.make_dataset(filename=opts.filename, size=opts.graph_size, num_samples=opts.val_size, offset=opts.offset)
    fleet_info, distance = load_fleet_info()
    start = time.time()
    total, _ = val(dataset, opts, fleet_info, distance, AGH)
    torch.cuda.synchronize()
    print('>> End of validation within {:.2f}s'.format(time.time() - start))
    return total


if __name__ == '__main__':
    main()
