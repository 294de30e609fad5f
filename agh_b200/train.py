"""REINFORCE loop with rollout baseline for AGH: the reference's train.py entry points
(get_inner_model :16, validate :20-28, rollout :31-70, clip_grad_norms :84-101, train_epoch :104-156,
train_batch_agh :159-219) on top of the CUDA rollout path.

The 10 fleets are solved in fleet_info['order']; the serve times of precedence level p become the
time-window left edges of level p+1 (train.py:67).  Fleet-batch assembly (train.py:40-57) is one fused
kernel (agh_fleet_task) instead of ~15 torch ops and a B x 33.9 KB distance-table copy per fleet.
"""
from __future__ import annotations

import math
import os
import time

import torch
from torch.nn import DataParallel
from torch.utils.data import DataLoader
from tqdm import tqdm

from . import ops, parallel
from .nets.attention_model import set_decode_type
from .utils import move_to
from .utils.log_utils import log_values


def get_inner_model(model):
    return model.module if isinstance(model, (DataParallel, torch.nn.parallel.DistributedDataParallel)) else model


def iter_batches(dataset, batch_size, lo=0, hi=None, no_progress_bar=True):
    """Yield collated batches of dataset[lo:hi].  Datasets that keep one tensor per field (`arrays`,
    AGHDataset) are sliced directly; anything else goes through a DataLoader."""
    hi = len(dataset) if hi is None else hi
    arrays = getattr(dataset, 'arrays', None)
    if arrays is not None:
        for s in tqdm(range(lo, hi, batch_size), disable=no_progress_bar):
            yield {k: v[s:min(s + batch_size, hi)] for k, v in arrays.items()}
    else:
        sub = torch.utils.data.Subset(dataset, range(lo, hi)) if (lo, hi) != (0, len(dataset)) else dataset
        for bat in tqdm(DataLoader(sub, batch_size=batch_size), disable=no_progress_bar):
            yield bat


LEVEL_BATCH_LIMIT = 16384      # instance-fleets per launch up to which same-level fleets are solved together


def fleet_rollout(model, bat, device, per_fleet=None, level_batch=None):
    """Solve the 10 fleets of one batch in order.  `bat`: dict of loc/arrival/departure/type/demand on any
    device.  Returns (costs [B,10], lls list) on `device`; per_fleet(k, f, cost, ll, serve) is an optional hook.

    Fleets of one precedence level only read tw_left of their own level and max-accumulate into the next one
    (train.py:42,67), so they are independent of each other.  In greedy eval mode (instances do not interact:
    BatchNorm uses running statistics) they are concatenated along the batch and solved by ONE encode / decode
    launch sequence -- 5 dependent stages instead of 10, which matters when B alone does not fill
    the GPU.  Per-instance results are identical to the one-fleet-at-a-time order.  level_batch: None = automatic
    (small batches, greedy eval), False = never."""
    inner = get_inner_model(model)
    info = inner.fleet_info
    x = move_to(bat, device)
    loc, type_ = x['loc'].contiguous(), x['type'].contiguous()
    departure, demand_all = x['departure'].contiguous(), x['demand'].contiguous()
    B = loc.size(0)
    lv = x['arrival'].repeat(len(info['next_duration']) + 1, 1, 1).contiguous()         # [6,B,N]  (train.py:40)
    order = list(info['order'])
    if level_batch is None:
        level_batch = (not inner.training) and getattr(inner, 'decode_type', None) == 'greedy'
    groups = []                                    # consecutive fleets of `order` that share a precedence level
    for k, f in enumerate(order):
        p = info['precedence'][f]
        if level_batch and groups and groups[-1][0] == p and (len(groups[-1][1]) + 1) * B <= LEVEL_BATCH_LIMIT:
            groups[-1][1].append((k, f))
        else:
            groups.append((p, [(k, f)]))
    costs, lls = [None] * len(order), [None] * len(order)
    for p, members in groups:
        nd = info['next_duration'][p]
        tasks = [ops.fleet_task(loc, type_, departure, demand_all, lv, p, f, info['duration'][f], nd,
                                nd_is_f32=all(type(v) is float for v in nd)) for _, f in members]
        task = tasks[0] if len(tasks) == 1 else ops.cat_tasks(tasks)
        cost, ll, serve = model(task)
        for i, (k, f) in enumerate(members):
            c, l, s = cost[i * B:(i + 1) * B], ll[i * B:(i + 1) * B], serve[i * B:(i + 1) * B]
            costs[k], lls[k] = c, l
            ops.chain_tw_left(lv, p, s.contiguous())
            if per_fleet is not None:
                per_fleet(k, f, c, l, s)
    return torch.stack(costs, 1), lls


def validate(model, dataset, opts):
    print('Validating...')
    cost = rollout(model, dataset, opts)
    if get_inner_model(model).is_agh:
        cost = cost.sum(1)
    avg_cost = cost.mean()
    print('Validation overall avg_cost: {} +- {}'.format(avg_cost, torch.std(cost) / math.sqrt(len(cost))))
    return avg_cost


def rollout(model, dataset, opts):
    """Greedy, eval-mode rollout of the whole dataset -> per-fleet costs [len(dataset), 10] (CPU f32,
    columns in fleet_info['order']).  With torch.distributed initialised the instances are sharded in
    contiguous blocks over the ranks and the costs all-gathered (every rank returns the full table)."""
    set_decode_type(model, 'greedy')
    model.eval()
    inner = get_inner_model(model)
    rank, ws = parallel.world()
    if not getattr(opts, 'shard_across_ranks', True):     # every rank owns its whole dataset (weak-scaling benchmarks)
        rank, ws = 0, 1
    lo, hi = parallel.shard_bounds(len(dataset), rank, ws)
    defer, inner.defer_checks = inner.defer_checks, True      # tour validity is asserted once per batch
    out = []
    try:
        with torch.no_grad():
            for bat in iter_batches(dataset, opts.eval_batch_size, lo, hi, getattr(opts, 'no_progress_bar', True)):
                costs, _ = fleet_rollout(model, bat, opts.device)
                inner.raise_pending()
                out.append(costs)
    finally:
        inner.defer_checks = defer
    local = torch.cat(out, 0) if out else torch.zeros(0, 10, device=opts.device)
    if ws > 1:
        local = parallel.all_gather_rows(local, len(dataset))
    return local.cpu()


def clip_grad_norms(param_groups, max_norm=math.inf):
    """Clip each group's gradient 2-norm to max_norm; returns (norms, clipped norms)."""
    norms = [torch.nn.utils.clip_grad_norm_(g['params'], max_norm if max_norm > 0 else math.inf, norm_type=2)
             for g in param_groups]
    clipped = [min(n, max_norm) for n in norms] if max_norm > 0 else norms
    return norms, clipped


def train_epoch(model, optimizer, baseline, lr_scheduler, epoch, val_dataset, problem, tb_logger, opts):
    print('Start train epoch {}, lr={} for run {}'.format(epoch, optimizer.param_groups[0]['lr'], opts.run_name))
    step = epoch * (opts.epoch_size // opts.batch_size)
    start_time = time.time()
    if not opts.no_tensorboard and tb_logger is not None:
        tb_logger.log_value('learnrate_pg0', optimizer.param_groups[0]['lr'], step)

    # --device_data: the epoch's instances are generated on the GPU (one kernel, no host pass / H2D copy; Philox stream,
    # not the reference's numpy stream)
    gen_dev = opts.device if getattr(opts, 'device_data', False) else None
    training_dataset = baseline.wrap_dataset(
        problem.make_dataset(size=opts.graph_size, num_samples=opts.epoch_size, distribution=opts.data_distribution,
                             **({'device': gen_dev} if gen_dev is not None else {})))
    model.train()
    set_decode_type(model, 'sampling')
    rank, ws = parallel.world()
    id_offset0 = get_inner_model(model).id_offset
    for batch_id, batch in enumerate(iter_batches(training_dataset, opts.batch_size,
                                                  no_progress_bar=opts.no_progress_bar)):
        lo = 0
        if ws > 1:      # each rank trains on its slice of the global batch; gradients are averaged (C1)
            n_b = batch['baseline'].size(0)
            lo, hi = parallel.shard_bounds(n_b, rank, ws)
            batch = {'data': {k: v[lo:hi] for k, v in batch['data'].items()}, 'baseline': batch['baseline'][lo:hi]}
        # Philox counters carry the GLOBAL instance index: ranks draw independent noise, results do not depend on G
        get_inner_model(model).id_offset = batch_id * opts.batch_size + lo
        train_batch_agh(model, optimizer, baseline, epoch, batch_id, step, batch, tb_logger, opts)
        step += 1
    get_inner_model(model).id_offset = id_offset0

    epoch_duration = time.time() - start_time
    print('Finished epoch {}, took {} s'.format(epoch, time.strftime('%H:%M:%S', time.gmtime(epoch_duration))))
    if rank == 0 and ((opts.checkpoint_epochs != 0 and epoch % opts.checkpoint_epochs == 0) or epoch == opts.n_epochs - 1):
        print('Saving model and state...')
        torch.save({'model': get_inner_model(model).state_dict(), 'optimizer': optimizer.state_dict(),
                    'rng_state': torch.get_rng_state(), 'cuda_rng_state': torch.cuda.get_rng_state_all(),
                    'baseline': baseline.state_dict()},
                   os.path.join(opts.save_dir, 'epoch-{}.pt'.format(epoch)))

    avg_reward = validate(model, val_dataset, opts)
    if rank == 0:
        with open(os.path.join(opts.save_dir, 'validate_log.txt'), 'a') as f:
            f.write('Validating Epoch {}, Validation avg_cost: {}\n\n'.format(epoch, avg_reward))
    if not opts.no_tensorboard and tb_logger is not None:
        tb_logger.log_value('val_avg_reward', avg_reward, step)
    baseline.epoch_callback(model, epoch)
    lr_scheduler.step()


def train_batch_agh(model, optimizer, baseline, epoch, batch_id, step, batch, tb_logger, opts):
    """One REINFORCE step (train.py:159-219): 10 sampling forwards with grad, loss
    mean_f mean_b (cost_f - bl_f) * ll_f, backward, [C1 all-reduce], clip, optimizer step."""
    x, bl_val = baseline.unwrap_batch(batch)
    assert bl_val is not None
    bl_val = move_to(bl_val, opts.device)
    set_decode_type(model, 'sampling')
    inner = get_inner_model(model)
    defer, inner.defer_checks = inner.defer_checks, True                 # tour validity: one host sync per batch, not per fleet
    try:
        costs, lls = fleet_rollout(model, x, opts.device)                # costs [B,10], lls: 10 x [B] with grad
        inner.raise_pending()
    finally:
        inner.defer_checks = defer
    loss = sum(((costs[:, i] - bl_val[:, i]) * lls[i]).mean() for i in range(len(lls))) / len(lls)

    optimizer.zero_grad()
    loss.backward()
    parallel.allreduce_grads_([p for g in optimizer.param_groups for p in g['params']],
                              buffers=list(get_inner_model(model).buffers()))
    grad_norms = clip_grad_norms(optimizer.param_groups, opts.max_grad_norm)
    optimizer.step()

    if step % int(opts.log_step) == 0:
        log_values(costs.sum(1), grad_norms, epoch, batch_id, step, sum(lls), loss, 0, tb_logger, opts)
