"""Training flags: same names, defaults and derived fields as the reference's options.py (:7-97), so that
`args.json` files written by either code base are interchangeable.  Table-driven; flags of the out-of-scope
ablations are accepted for compatibility and rejected when they would change behaviour."""
from __future__ import annotations

import argparse
import os
import time

import torch

_FLAGS = [
    # (name, type or 'flag', default, help)
    ('problem', str, 'agh', 'problem to solve (only agh)'),
    ('graph_size', int, 20, 'number of flights'),
    ('batch_size', int, 64, 'instances per training batch'),
    ('epoch_size', int, 12800, 'instances per epoch'),
    ('val_size', int, 1000, 'instances used for validation'),
    ('val_dataset', str, None, 'dataset file for validation'),
    ('model', str, 'attention', "model ('attention')"),
    ('embedding_dim', int, 128, 'embedding width'),
    ('hidden_dim', int, 128, 'hidden width'),
    ('n_encode_layers', int, 3, 'encoder layers'),
    ('tanh_clipping', float, 10., 'logit clipping'),
    ('normalization', str, 'batch', "normalisation ('batch')"),
    ('optimizer', str, 'Adam', 'optimizer'),
    ('lr_model', float, 1e-4, 'actor learning rate'),
    ('lr_critic', float, 1e-4, 'critic learning rate (unused)'),
    ('lr_decay', float, 1.0, 'learning-rate decay per epoch'),
    ('eval_only', 'flag', False, 'only evaluate'),
    ('n_epochs', int, 100, 'epochs to train'),
    ('seed', int, 1234, 'random seed'),
    ('max_grad_norm', float, 1.0, 'gradient-norm clip (0 disables)'),
    ('no_cuda', 'flag', False, 'rejected: there is no CPU path'),
    ('exp_beta', float, 0.8, 'unused'),
    ('baseline', str, 'rollout', "baseline ('rollout')"),
    ('bl_alpha', float, 0.05, 't-test significance for baseline updates'),
    ('bl_warmup_epochs', int, 0, 'unused (0)'),
    ('eval_batch_size', int, 1000, 'batch size of greedy rollouts'),
    ('checkpoint_encoder', 'flag', False, 'accepted, not needed'),
    ('shrink_size', int, None, 'accepted, not needed'),
    ('data_distribution', str, None, 'unused for agh'),
    ('log_step', int, 50, 'log every n steps'),
    ('log_dir', str, 'logs', 'unused'),
    ('run_name', str, 'run', 'run name'),
    ('output_dir', str, 'outputs', 'where checkpoints go'),
    ('epoch_start', int, 0, 'first epoch number'),
    ('checkpoint_epochs', int, 1, 'checkpoint every n epochs (0: never)'),
    ('load_path', str, None, 'load model/optimizer state from this file'),
    ('resume', str, None, 'resume from this epoch-N.pt'),
    ('no_tensorboard', 'flag', False, 'forced on'),
    ('no_progress_bar', 'flag', False, 'forced on'),
    ('fine_tune', 'flag', False, 'accepted'),
    ('wo_time', 'flag', False, 'ablation, out of scope'),
    ('rnn_time', 'flag', False, 'ablation, out of scope'),
    ('device_data', 'flag', False, "generate each epoch's training instances on the GPU (Philox stream, not the numpy one)"),
]


def get_options(args=None):
    parser = argparse.ArgumentParser(description='REINFORCE training of the AGH attention model on the B200 path')
    for name, typ, default, help_ in _FLAGS:
        if typ == 'flag':
            parser.add_argument('--' + name, action='store_true', help=help_)
        else:
            parser.add_argument('--' + name, type=typ, default=default, help=help_)
    opts = parser.parse_args(args)
    assert opts.baseline == 'rollout', "Only support for 'rollout' baseline now!"
    assert opts.problem == 'agh' and opts.model == 'attention'
    opts.use_cuda = torch.cuda.is_available() and not opts.no_cuda
    opts.run_name = '{}_{}'.format(opts.run_name, time.strftime('%Y%m%dT%H%M%S'))
    opts.no_progress_bar = opts.no_tensorboard = opts.log_stdout = True
    opts.save_dir = os.path.join(opts.output_dir, '{}_{}'.format(opts.problem, opts.graph_size), opts.run_name)
    assert opts.bl_warmup_epochs == 0
    assert opts.epoch_size % opts.batch_size == 0, 'Epoch size must be integer multiple of batch size!'
    return opts
