"""Greedy-rollout baseline for REINFORCE: the reference's reinforce_baselines.py RolloutBaseline
(:144-245) and BaselineDataset (:247-263).  The baseline model's rollouts run on the CUDA path
(train.rollout), sharded over ranks when torch.distributed is initialised."""
from __future__ import annotations

import copy

import torch
from scipy.stats import ttest_rel
from torch.utils.data import Dataset

from .train import rollout, get_inner_model


class Baseline(object):
    def wrap_dataset(self, dataset):
        return dataset

    def unwrap_batch(self, batch):
        return batch, None

    def eval(self, x, c):
        raise NotImplementedError('Override this method')

    def get_learnable_parameters(self):
        return []

    def epoch_callback(self, model, epoch):
        pass

    def state_dict(self):
        return {}

    def load_state_dict(self, state_dict):
        pass


class RolloutBaseline(Baseline):

    def __init__(self, model, problem, opts, epoch=0):
        super().__init__()
        self.problem = problem
        self.opts = opts
        self._update_model(model, epoch)

    def _update_model(self, model, epoch, dataset=None):
        self.model = copy.deepcopy(model)
        if dataset is not None:
            if len(dataset) != self.opts.val_size:
                print('Warning: not using saved baseline dataset since val_size does not match')
                dataset = None
            elif dataset[0]['loc'].size(0) != self.opts.graph_size:
                print('Warning: not using saved baseline dataset since graph_size does not match')
                dataset = None
        # a fresh evaluation set on every baseline update (prevents over-fitting to it)
        self.dataset = dataset if dataset is not None else self.problem.make_dataset(
            size=self.opts.graph_size, num_samples=self.opts.val_size, distribution=self.opts.data_distribution)
        print('Evaluating baseline model on evaluation dataset')
        self.bl_vals = rollout(self.model, self.dataset, self.opts).cpu().numpy().sum(1)
        self.mean = self.bl_vals.mean()
        self.epoch = epoch

    def wrap_dataset(self, dataset):
        print('Evaluating baseline on dataset...')
        return BaselineDataset(dataset, rollout(self.model, dataset, self.opts))      # per-fleet costs [len,10]

    def unwrap_batch(self, batch):
        return batch['data'], batch['baseline']

    def eval(self, x, c):
        with torch.no_grad():
            v = self.model(x)[0]
        return v, 0

    def epoch_callback(self, model, epoch):
        """Challenge the baseline with the current model; replace it when a one-sided paired t-test says
        the improvement is significant at bl_alpha (reinforce_baselines.py:202-226)."""
        print('Evaluating candidate model on evaluation dataset')
        candidate_vals = rollout(model, self.dataset, self.opts).cpu().numpy().sum(1)
        candidate_mean = candidate_vals.mean()
        print('Epoch {} candidate mean {}, baseline epoch {} mean {}, difference {}'.format(
            epoch, candidate_mean, self.epoch, self.mean, candidate_mean - self.mean))
        if candidate_mean - self.mean < 0:
            t, p = ttest_rel(candidate_vals, self.bl_vals)
            p_val = p / 2
            assert t < 0, 'T-statistic should be negative'
            print('p-value: {}'.format(p_val))
            if p_val < self.opts.bl_alpha:
                print('Update baseline')
                self._update_model(model, epoch)

    def state_dict(self):
        return {'model': self.model, 'dataset': self.dataset, 'epoch': self.epoch}

    def load_state_dict(self, state_dict):
        load_model = copy.deepcopy(self.model)
        inner = get_inner_model(load_model)
        inner.load_state_dict({**inner.state_dict(), **state_dict['model'].state_dict()})
        self._update_model(load_model, state_dict['epoch'], state_dict['dataset'])


class BaselineDataset(Dataset):
    """Items {'data': instance dict, 'baseline': f32 [10]}; keeps `arrays` so batches can be sliced."""

    def __init__(self, dataset=None, baseline=None):
        super().__init__()
        self.dataset = dataset
        self.baseline = baseline
        assert len(self.dataset) == len(self.baseline)
        inner = getattr(dataset, 'arrays', None)
        if inner is not None:
            self.arrays = _Nested(inner, baseline)

    def __getitem__(self, item):
        return {'data': self.dataset[item], 'baseline': self.baseline[item]}

    def __len__(self):
        return len(self.dataset)


class _Nested(dict):
    """`arrays` view of a BaselineDataset: slicing yields {'data': {...}, 'baseline': ...} batches."""

    def __init__(self, data_arrays, baseline):
        super().__init__(data=data_arrays, baseline=baseline)

    def items(self):
        return [('data', _SliceDict(self['data'])), ('baseline', self['baseline'])]


class _SliceDict(object):
    def __init__(self, arrays):
        self.arrays = arrays

    def __getitem__(self, sl):
        return {k: v[sl] for k, v in self.arrays.items()}
