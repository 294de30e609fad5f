"""Utilities on the AGH path: reference utils/functions.py (load_problem :13, move_to :32, load_model :80-130,
parse_softmax_temperature :133).  The reference's do_batch_rep / sample_many (:171-223) have no counterpart here:
AttentionModel.sample_many runs the replicas on ONE staged key set per instance (csrc/decode_shared.cu) and reduces them with
agh_best_of instead of materialising batch_rep copies of every input."""
from __future__ import annotations

import json
import os

import numpy as np
import torch


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
