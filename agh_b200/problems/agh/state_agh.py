"""StateAGH: the reference's environment state API (state_agh.py:7-178) backed by the C-ABI
single-step kernels agh_state_get_mask / agh_state_update.

Same field names and shapes as the reference's NamedTuple (prev_a [B,1], used_capacity [B,1],
visited_ [B,1,n] uint8, lengths [B,1], cur_free_time [B,1], serve_time [B,n], tour [B,steps], ...),
same functional style: update() returns a new state and, like the reference (state_agh.py:130),
writes serve_time in place.  Differences, by design: cur_free_time is float64 from step 0 (the
reference starts it as int64 -60; SURVEY 8a shows the step-0 arithmetic is identical), and the
distance table is one [8464] row shared by all instances rather than a stride-0 [B,8464] view.
"""
from __future__ import annotations

import ctypes as C

import torch

from ... import _lib
from ..._lib import lib, check, ptr, stream_ptr


class StateAGH(object):
    VEHICLE_CAPACITY = 1.0
    SPEED = 110.0
    NODE_SIZE = 92

    _FIELDS = ('coords', 'demand', 'tw_left', 'tw_right', 'distance', 'duration', 'ids', 'prev_a', 'used_capacity',
               'visited_', 'lengths', 'cur_coord', 'cur_free_time', 'serve_time', 'tour', 'i')

    def __init__(self, **kw):
        for k in self._FIELDS + ('_coord8', '_twr64', '_dur64', '_dist'):
            setattr(self, k, kw[k])

    def _replace(self, **kw):
        d = {k: getattr(self, k) for k in self._FIELDS + ('_coord8', '_twr64', '_dur64', '_dist')}
        d.update(kw)
        return StateAGH(**d)

    @property
    def visited(self):
        return self.visited_

    def __getitem__(self, key):
        assert torch.is_tensor(key) or isinstance(key, slice)
        return self._replace(**{k: getattr(self, k)[key] for k in
                                ('ids', 'prev_a', 'used_capacity', 'visited_', 'lengths', 'cur_coord',
                                 'cur_free_time', 'tour', 'serve_time', 'coords', 'demand', 'tw_left', 'tw_right',
                                 'duration', '_coord8', '_twr64', '_dur64')})

    @staticmethod
    def initialize(input, visited_dtype=torch.uint8):
        assert visited_dtype == torch.uint8, 'only the uint8 visited mask is on the AGH path (state_agh.py:61)'
        loc = input['loc']
        B, N = loc.shape
        dev = loc.device
        coords = torch.cat((torch.zeros_like(loc[:, :1]), loc), 1)
        duration = torch.cat((torch.zeros_like(loc[:, :1]), input['duration']), 1)
        dist = input['distance']
        dist1 = (dist[0] if dist.dim() == 2 else dist).to(torch.float32).contiguous()
        return StateAGH(
            coords=coords, demand=input['demand'].to(torch.float32).contiguous(),
            tw_left=input['tw_left'].to(torch.float32).contiguous(), tw_right=input['tw_right'], distance=dist,
            duration=duration, ids=torch.arange(B, dtype=torch.int64, device=dev)[:, None],
            prev_a=torch.zeros(B, 1, dtype=torch.long, device=dev),
            used_capacity=torch.zeros(B, 1, dtype=torch.float32, device=dev),
            visited_=torch.zeros(B, 1, N + 1, dtype=torch.uint8, device=dev),
            lengths=torch.zeros(B, 1, device=dev), cur_coord=torch.zeros_like(loc[:, :1]),
            cur_free_time=torch.full((B, 1), -60.0, dtype=torch.float64, device=dev),
            serve_time=torch.zeros(B, N + 1, device=dev), tour=torch.zeros(B, 1, dtype=torch.long, device=dev),
            i=torch.zeros(1, dtype=torch.int64, device=dev),
            _coord8=coords.to(torch.uint8).contiguous(), _twr64=input['tw_right'].to(torch.float64).contiguous(),
            _dur64=duration.to(torch.float64).contiguous(), _dist=dist1)

    def _args(self, prev_a, used, visited, lengths, cft, serve):
        B, n = self._coord8.shape
        return _lib.StateArgs(coord=ptr(self._coord8), demand=ptr(self.demand), tw_left=ptr(self.tw_left),
                              tw_right=ptr(self._twr64), duration=ptr(self._dur64), distance=ptr(self._dist),
                              prev_a=ptr(prev_a, torch.int64), used_capacity=ptr(used, torch.float32),
                              visited=ptr(visited, torch.uint8), lengths=ptr(lengths, torch.float32),
                              cur_free_time=ptr(cft, torch.float64), serve_time=ptr(serve, torch.float32), B=B, N=n - 1)

    def update(self, selected):
        assert self.i.size(0) == 1, 'Can only update if state represents single step'
        sel = selected.reshape(-1).to(torch.int64).contiguous()
        prev_a, used, visited = self.prev_a.clone(), self.used_capacity.clone(), self.visited_.clone()
        lengths, cft = self.lengths.clone(), self.cur_free_time.clone()
        a = self._args(prev_a, used, visited, lengths, cft, self.serve_time)       # serve_time in place
        check(lib().agh_state_update(C.byref(a), ptr(sel), stream_ptr(sel.device)))
        return self._replace(prev_a=prev_a, used_capacity=used, visited_=visited, lengths=lengths,
                             cur_coord=self.coords.gather(1, prev_a), i=self.i + 1, cur_free_time=cft,
                             tour=torch.cat((self.tour, sel[:, None]), 1))

    def all_finished(self):
        return self.i.item() >= self.demand.size(-1) and bool(self.visited_.all())

    def get_finished(self):
        return self.visited_.sum(-1) == self.visited_.size(-1)

    def get_current_node(self):
        return self.prev_a

    def get_mask(self):
        """[B,1,n] bool, True = infeasible (state_agh.py:148-174)."""
        B, n = self._coord8.shape
        mask = torch.empty(B, 1, n, dtype=torch.uint8, device=self._coord8.device)
        a = self._args(self.prev_a, self.used_capacity, self.visited_, self.lengths, self.cur_free_time, self.serve_time)
        check(lib().agh_state_get_mask(C.byref(a), ptr(mask), stream_ptr(mask.device)))
        return mask.bool()
