"""AGH problem definition: drop-in for the reference's problems/agh/problem_agh.py
(get_costs :17-66, make_dataset :68-70, make_state :72-74, AGHDataset :93-122)."""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch
from torch.utils.data import Dataset

from ... import _lib, ops
from ...ops import FleetTask
from .state_agh import StateAGH

_TABLES = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'agh_tables.npz')
CAPACITIES = {10: 20., 20: 30., 50: 40., 100: 50., 200: 60., 300: 70.}


class AGH(object):
    NAME = 'agh'
    VEHICLE_CAPACITY = 1.0
    SPEED = 110.0
    NODE_SIZE = 92

    @staticmethod
    def get_costs(dataset, pi):
        """Validate the tours and return (cost [B] f32, None); raises AssertionError like the
        reference's asserts (problem_agh.py:24-27, :44, :64).  One agh_get_costs launch."""
        task = dataset if isinstance(dataset, FleetTask) else FleetTask.from_input(dataset)
        dist = dataset['distance'] if not isinstance(dataset, FleetTask) else None
        if dist is None:
            raise ValueError('FleetTask input needs the distance table: use ops.get_costs')
        dist1 = (dist[0] if dist.dim() == 2 else dist).to(torch.float32).contiguous()
        cost, status = ops.get_costs(pi.to(torch.int64).contiguous(), task, dist1)
        bad = int(status.max().item()) if status.numel() else 0
        assert not (bad & _lib.AGH_TOUR_INVALID), 'Invalid tour'
        assert not (bad & _lib.AGH_CAPACITY_EXCEEDED), 'Used more than capacity'
        assert not (bad & _lib.AGH_TW_VIOLATED), 'time window violated'
        return cost, None

    @staticmethod
    def make_dataset(*args, **kwargs):
        return AGHDataset(*args, **kwargs)

    @staticmethod
    def make_state(*args, **kwargs):
        return StateAGH.initialize(*args, **kwargs)

    @staticmethod
    def beam_search(*args, **kwargs):
        raise NotImplementedError('beam search is a stub in the reference too (problem_agh.py:76-79)')


def generate_arrays(num_samples, size, fleet_size=10, rng=np.random):
    """Same draws in the same order from the same RNG as problem_agh.py:104-113 / generate_data.py:8-18."""
    prob = np.load(_TABLES)['arrival_prob']
    loc = 1 + rng.choice(91, size=(num_samples, size))
    arrival = 60 * rng.choice(24, size=(num_samples, size), p=prob) + rng.randint(0, 60, size=(num_samples, size))
    type_ = rng.randint(0, 3, size=(num_samples, size))
    departure = arrival + np.asarray([30, 34, 33])[type_]
    demand = rng.randint(1, 10, size=(num_samples, fleet_size, size)) / CAPACITIES[size]
    return {'loc': torch.from_numpy(loc.astype(np.int64)), 'arrival': torch.from_numpy(arrival.astype(np.float32)),
            'departure': torch.from_numpy(departure.astype(np.float32)), 'type': torch.from_numpy(type_.astype(np.int64)),
            'demand': torch.from_numpy(demand).to(torch.float32)}


def generate_arrays_device(num_samples, size, device, seed=None, fleet_size=10, id_offset=0):
    """The same distributions drawn ON THE DEVICE by one kernel (agh_generate: Philox4x32-10 keyed by seed, instance id
    and flight): a 12,800-instance epoch set without the host numpy pass and the H2D copy.  NOT bit-compatible with the
    reference's numpy stream -- generate_arrays stays the parity generator.  `seed` defaults to a draw from torch's CPU
    generator, so datasets are reproducible under torch.manual_seed and identical on every rank."""
    device = torch.device(device)
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    cdf = torch.from_numpy(np.cumsum(np.load(_TABLES)['arrival_prob']).astype(np.float32))
    cdf[-1] = 1.0
    cdf = cdf.to(device)
    out = {'loc': torch.empty(num_samples, size, dtype=torch.int64, device=device),
           'arrival': torch.empty(num_samples, size, dtype=torch.float32, device=device),
           'departure': torch.empty(num_samples, size, dtype=torch.float32, device=device),
           'type': torch.empty(num_samples, size, dtype=torch.int64, device=device),
           'demand': torch.empty(num_samples, fleet_size, size, dtype=torch.float32, device=device)}
    with torch.cuda.device(device):
        _lib.check(_lib.lib().agh_generate(seed & (2 ** 64 - 1), int(id_offset), num_samples, size, fleet_size, float(CAPACITIES[size]),
                                           _lib.ptr(cdf), _lib.ptr(out['loc']), _lib.ptr(out['arrival']), _lib.ptr(out['departure']),
                                           _lib.ptr(out['type']), _lib.ptr(out['demand']), _lib.stream_ptr(device)))
    return out


class AGHDataset(Dataset):
    """Items are dicts {'loc' i64 [N], 'arrival' f32 [N], 'departure' f32 [N], 'type' i64 [N], 'demand' f32 [10,N]}.
    Storage is one tensor per field (`self.arrays`) so whole batches can be sliced without collation."""

    def __init__(self, filename=None, size=50, num_samples=1000000, offset=0, distribution=None, fleet_size=10,
                 device=None, seed=None):
        """device: a CUDA device generates the instances there (generate_arrays_device; not the numpy stream); the
        default (None) is the reference's host generator, draw for draw."""
        super().__init__()
        if filename is None and device is not None and torch.device(device).type == 'cuda':
            self.arrays = generate_arrays_device(num_samples, size, device, seed=seed, fleet_size=fleet_size, id_offset=offset)
        elif filename is not None:
            assert os.path.splitext(filename)[1] == '.pkl'
            with open(filename, 'rb') as f:
                data = pickle.load(f)[offset:offset + num_samples]
            self.arrays = {'loc': torch.tensor([d[0] for d in data], dtype=torch.long),
                           'arrival': torch.tensor([d[1] for d in data], dtype=torch.float),
                           'departure': torch.tensor([d[2] for d in data], dtype=torch.float),
                           'type': torch.tensor([d[3] for d in data], dtype=torch.long),
                           'demand': torch.tensor([d[4] for d in data], dtype=torch.float)}
        else:
            arr = generate_arrays(num_samples, size, fleet_size)
            self.arrays = {k: v[offset:offset + num_samples] for k, v in arr.items()}
        self.size = self.arrays['loc'].size(0)

    def __setstate__(self, state):
        """Datasets pickled by the reference (inside a checkpoint's baseline, reinforce_baselines.py:228-233) hold a
        list of per-instance dicts `data` (problem_agh.py:115): rebuild the per-field arrays from it."""
        self.__dict__.update(state)
        if 'arrays' not in state and 'data' in state:
            data = state['data']
            keys = ('loc', 'arrival', 'departure', 'type', 'demand')
            self.arrays = {k: torch.stack([d[k] for d in data]) for k in keys} if len(data) else {}
            self.size = len(data)

    def __len__(self):
        return self.size

    def __getitem__(self, idx):
        return {k: v[idx] for k, v in self.arrays.items()}
