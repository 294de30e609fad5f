"""Utilities on the AGH path: reference utils/functions.py (load_problem :13, move_to :32,
load_model :80-130, do_batch_rep :171-179, sample_many :182-223, parse_softmax_temperature :133)."""
from __future__ import annotations

import json
import os

import numpy as np
import torch
import torch.nn.functional as F


def load_problem(name):
    from ..problems import AGH
    assert name == 'agh', 'Currently unsupported problem: {}!'.format(name)
    return AGH


def torch_load_cpu(load_path):
    """torch.load on the CPU.  A checkpoint written by the reference pickles whole modules under the reference's
    module paths (reinforce_baselines.py:228-233): when those do not resolve, the load is retried with the paths
    aliased to this package's classes (utils/compat.py)."""
    try:
        return torch.load(load_path, map_location='cpu', weights_only=False)
    except (ModuleNotFoundError, AttributeError):
        from .compat import reference_module_aliases
        with reference_module_aliases():
            return torch.load(load_path, map_location='cpu', weights_only=False)


def move_to(var, device):
    if isinstance(var, dict):
        return {k: move_to(v, device) for k, v in var.items()}
    return var.to(device, non_blocking=True)


def load_args(filename):
    with open(filename, 'r') as f:
        args = json.load(f)
    args.setdefault('data_distribution', None)
    return args


def _load_model_file(load_path, model):
    """Merge a checkpoint into `model`; returns (model, optimizer_state_dict or None)."""
    print('  [*] Loading model from {}'.format(load_path))
    data = torch_load_cpu(os.path.join(os.getcwd(), load_path))
    opt_sd = None
    if isinstance(data, dict):
        opt_sd = data.get('optimizer', None)
        model_sd = data.get('model', data)
    else:
        model_sd = data.state_dict()
    merged = model.state_dict()
    merged.update(model_sd)
    model.load_state_dict(merged)
    return model, opt_sd


def load_model(path, epoch=None):
    """Load `epoch-N.pt` + `args.json` written by the reference (or by agh_b200.train): same
    state_dict keys, so reference checkpoints load unchanged."""
    from ..nets.attention_model import AttentionModel
    if os.path.isfile(path):
        model_filename, path = path, os.path.dirname(path)
    else:
        assert os.path.isdir(path), '{} is not a valid directory or file'.format(path)
        if epoch is None:
            epoch = max(int(os.path.splitext(fn)[0].split('-')[1]) for fn in os.listdir(path)
                        if os.path.splitext(fn)[1] == '.pt')
        model_filename = os.path.join(path, 'epoch-{}.pt'.format(epoch))
    args = load_args(os.path.join(path, 'args.json'))
    assert args.get('model', 'attention') == 'attention', 'only the attention model is on the AGH path'
    model = AttentionModel(args['embedding_dim'], args['hidden_dim'], load_problem(args['problem']),
                           n_encode_layers=args['n_encode_layers'], mask_inner=True, mask_logits=True,
                           normalization=args['normalization'], tanh_clipping=args['tanh_clipping'],
                           checkpoint_encoder=args.get('checkpoint_encoder', False),
                           shrink_size=args.get('shrink_size', None), wo_time=args.get('wo_time', False),
                           rnn_time=args.get('rnn_time', False))
    data = torch_load_cpu(model_filename)
    model.load_state_dict({**model.state_dict(), **data.get('model', {})})
    model.eval()
    return model, args


def parse_softmax_temperature(raw_temp):
    if os.path.isfile(raw_temp):
        return np.loadtxt(raw_temp)[-1, 0]
    return float(raw_temp)


def do_batch_rep(v, n):
    if isinstance(v, dict):
        return {k: do_batch_rep(v_, n) for k, v_ in v.items()}
    if isinstance(v, (list, tuple)):
        return type(v)(do_batch_rep(v_, n) for v_ in v)
    return v[None, ...].expand(n, *v.size()).contiguous().view(-1, *v.size()[1:])


def sample_many(inner_func, get_cost_func, input, batch_rep=1, iter_rep=1):
    """Generic best-of-k driver with the reference's signature.  AttentionModel.sample_many does not
    go through it (it shares the staged keys across replicas on the device); kept for callers that pass
    their own inner/cost functions."""
    input = do_batch_rep(input, batch_rep)
    costs, pis, serves = [], [], []
    for _ in range(iter_rep):
        _log_p, pi, serve_time = inner_func(input)
        cost, _ = get_cost_func(input, pi)
        costs.append(cost.view(batch_rep, -1).t())
        pis.append(pi.view(batch_rep, -1, pi.size(-1)).transpose(0, 1))
        serves.append(serve_time.view(batch_rep, -1, serve_time.size(-1)).transpose(0, 1))
    width = max(p.size(-1) for p in pis)
    pis = torch.cat([F.pad(p, (0, width - p.size(-1))) for p in pis], 1)
    costs, serves = torch.cat(costs, 1), torch.cat(serves, 1)
    mincosts, arg = costs.min(-1)
    rows = torch.arange(pis.size(0), device=arg.device)
    return pis[rows, arg], mincosts, serves[rows, arg]
