"""AttentionModel for the AGH problem: drop-in for the reference's nets/attention_model.py on the
AGH path (same constructor, attributes, state_dict keys and return tuples), with the work done by
libagh_b200's sm_100a kernels.

Reference call sites mirrored here (CB/nets/attention_model.py):
  forward :159-196, sample_many :358-370, set_decode_type :154-157 / module fn :16-19,
  _init_embed :272-294, _inner :298-356, _precompute :392-414, _get_log_p :429-451,
  _calc_log_likelihood :238-250.

Eval / no-grad forward:  agh_encode -> agh_precompute -> agh_decode  (3 C-ABI calls, no host sync).
Training forward (grad enabled, model.train()): nets/train_fn.py -- the encoder runs through the library's training
primitives with batch-statistics BatchNorm, the sampling rollout IS the forward pass (agh_decode returns the
log-likelihood and records the per-step tensors), and the backward is hand-written (csrc/train.cu): no cuBLAS, SDPA or
ATen contraction anywhere on the path.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import DataParallel

from .. import _lib, ops
from ..ops import FleetTask
from .graph_encoder import GraphAttentionEncoder
from .weights import pack_weights

_TABLES = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'problems', 'agh', 'agh_tables.npz')


def _on_model_device(fn):
    """The C ABI launches on the CURRENT CUDA device: run the method with the model's device selected."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *args, **kw):
        dev = self._device()
        if dev.type != 'cuda':
            raise _lib.AghError('agh_b200 kernels need the model on a CUDA device (it is on %s); there is no CPU path' % dev)
        with torch.cuda.device(dev):
            return fn(self, *args, **kw)
    return wrapper


def set_decode_type(model, decode_type):
    if isinstance(model, (DataParallel, torch.nn.parallel.DistributedDataParallel)):
        model = model.module
    model.set_decode_type(decode_type)


def load_fleet_info(path: str = _TABLES):
    """fleet_info.pkl / distance.pkl content (attention_model.py:78-88) from the packaged npz, with the
    reference's container types: dict of lists, numpy f64 scalars except precedence level 4."""
    z = np.load(path)
    order = [int(x) for x in z['order']]
    info = {'order': order, 'precedence': {f: int(z['precedence'][f]) for f in range(1, 11)},
            'duration': {f: [np.float64(x) for x in z['duration'][f]] for f in range(1, 11)},
            'next_duration': {p: ([float(x) for x in z['next_duration'][p]] if z['next_duration_is_f32'][p]
                                  else [np.float64(x) for x in z['next_duration'][p]]) for p in range(5)}}
    # keep the reference's key order (it only matters for printing)
    info['precedence'] = {f: info['precedence'][f] for f in [1, 2, 4, 8, 3, 5, 7, 9, 6, 10]}
    return info, torch.from_numpy(z['distance'].copy())


class AttentionModel(nn.Module):

    def __init__(self, embedding_dim, hidden_dim, problem, n_encode_layers=3, tanh_clipping=10., mask_inner=True,
                 mask_logits=True, normalization='batch', n_heads=8, checkpoint_encoder=False, shrink_size=None,
                 wo_time=False, rnn_time=False):
        super().__init__()
        assert problem.NAME == 'agh', 'agh_b200 implements the AGH path only'
        if (embedding_dim, n_encode_layers, n_heads, float(tanh_clipping)) != (128, 3, 8, 10.0):
            raise NotImplementedError('kernels are specialised for embedding_dim=128, 3 layers, 8 heads, tanh 10 '
                                      '(every shipped args.json)')
        if wo_time or rnn_time or not (mask_inner and mask_logits):
            raise NotImplementedError('ablation flags wo_time / rnn_time / unmasked variants are out of scope')
        self.embedding_dim, self.hidden_dim, self.n_encode_layers = embedding_dim, hidden_dim, n_encode_layers
        self.decode_type = None
        self.temp = 1.0
        self.allow_partial = self.is_vrp = self.is_orienteering = self.is_pctsp = False
        self.is_agh = True
        self.wo_time, self.rnn_time, self.pre_tw = wo_time, rnn_time, None
        self.fleet_info, self.distance = load_fleet_info()
        self.tanh_clipping, self.mask_inner, self.mask_logits = tanh_clipping, mask_inner, mask_logits
        self.problem, self.n_heads = problem, n_heads
        self.checkpoint_encoder, self.shrink_size = checkpoint_encoder, shrink_size   # accepted, not needed

        # parameters: same creation order as the reference so that a seed reproduces its init
        self.fleets_embedding = nn.Embedding(10, embedding_dim)
        self.loc_embedding = nn.Embedding(92, embedding_dim)
        self.init_embed = nn.Linear(3, embedding_dim)
        self.embedder = GraphAttentionEncoder(n_heads=n_heads, embed_dim=embedding_dim, n_layers=n_encode_layers,
                                              normalization=normalization)
        self.project_node_embeddings = nn.Linear(embedding_dim, 3 * embedding_dim, bias=False)
        self.project_fixed_context = nn.Linear(embedding_dim, embedding_dim, bias=False)
        self.project_step_context = nn.Linear(embedding_dim + 2, embedding_dim, bias=False)
        self.project_out = nn.Linear(embedding_dim, embedding_dim, bias=False)

        # host-side knobs of the CUDA path
        self.check_tours = True          # run agh_get_costs' validity asserts like problem.get_costs does
        self.defer_checks = False        # collect the status words; raise_pending() syncs once
        self._pending_status = []
        self._sampling_seed, self._seed_calls = None, 0   # fixed Philox seed (else drawn from torch's generator per call)
        self.id_offset = 0               # global id of instance 0 of the batch (sharding-independent sampling)
        self._blob = None
        self._blob_key = (None, None)
        self._bn_dirty = False
        self._plist = self._blist = self._pdict = self._bdict = None
        self._dist_dev = None
        self.last_steps = None           # [B] i32 steps per instance of the last rollout (device)

    def _apply(self, fn, *args, **kw):
        # .to() / .cuda() may replace the parameter tensors: drop the cached lists and the packed weights
        self._plist = self._blist = self._pdict = self._bdict = None
        self._blob = None
        return super()._apply(fn, *args, **kw)

    # ------------------------------------------------------------------ plumbing
    def set_decode_type(self, decode_type, temp=None):
        self.decode_type = decode_type
        if temp is not None:
            self.temp = temp

    def _device(self):
        return self.project_out.weight.device

    def _tensor_lists(self):
        """(parameters, buffers) as lists, cached: the module structure is fixed after construction."""
        if self._plist is None:
            self._plist = [p for _, p in self.named_parameters()]
            self._blist = [b for _, b in self.named_buffers()]
        return self._plist, self._blist

    def weight_blob(self) -> torch.Tensor:
        """Packed weights, re-packed whenever a parameter was modified in place or moved -- and, outside training, whenever
        a BatchNorm buffer changed: the packed BatchNorm fold (scale / shift from the running statistics) is only read by the
        eval-mode encoder, so the ten train-mode forwards of a batch, each of which updates the statistics, share ONE pack."""
        params, bufs = self._tensor_lists()
        pkey = (self._device(), tuple(t._version for t in params), tuple(t.data_ptr() for t in params))
        bkey = None if self.training else (tuple(t._version for t in bufs), tuple(t.data_ptr() for t in bufs))
        # the training kernels update the running statistics through raw pointers (no version bump): _bn_dirty marks the
        # fold stale until the first use outside training.  invalidate_blob() forces a re-pack after in-place edits through
        # `.data`, which do not bump versions either.
        stale = self._blob is None or pkey != self._blob_key[0]
        if not self.training:
            stale = stale or self._bn_dirty or bkey != self._blob_key[1]
        if stale:
            if not self.training:
                self._bn_dirty = False
            else:
                self._bn_dirty = True          # packed in training: the fold was built from whatever the buffers held
            self._blob = pack_weights(self.state_dict(), device=self._device())
            self._blob_key = (pkey, bkey)
        return self._blob

    def _param_dict(self):
        if self._pdict is None:
            self._pdict = dict(self.named_parameters())
            self._bdict = dict(self.named_buffers())
        return self._pdict

    def _buffer_dict(self):
        self._param_dict()
        return self._bdict

    def invalidate_blob(self):
        self._blob = None

    def distance_table(self) -> torch.Tensor:
        dev = self._device()
        if self._dist_dev is None or self._dist_dev.device != dev:
            self._dist_dev = self.distance.to(dev).contiguous()
        return self._dist_dev

    @property
    def sampling_seed(self) -> Optional[int]:
        return self._sampling_seed

    @sampling_seed.setter
    def sampling_seed(self, value: Optional[int]):
        self._sampling_seed, self._seed_calls = value, 0

    def _seed(self) -> int:
        """Philox key of one sampling rollout.  A fixed sampling_seed is mixed with the number of rollouts drawn since
        it was set (the 10 fleet forwards and the iter_rep iterations of sample_many must not reuse the same noise);
        the first call returns the seed itself."""
        if self._sampling_seed is not None:
            s = (int(self._sampling_seed) + self._seed_calls * 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
            self._seed_calls += 1
            return s
        return int(torch.randint(0, 2 ** 62, (1,)).item())      # CPU generator: reproducible under torch.manual_seed

    def _mode(self) -> int:
        assert self.decode_type in ('greedy', 'sampling'), 'Unknown decode type'
        return _lib.AGH_GREEDY if self.decode_type == 'greedy' else _lib.AGH_SAMPLING

    @staticmethod
    def _as_task(input) -> FleetTask:
        return input if isinstance(input, FleetTask) else FleetTask.from_input(input)

    def raise_pending(self):
        """Raise the deferred get_costs assertion (one host sync for all fleet calls since the last check)."""
        if self._pending_status:
            st, self._pending_status = torch.stack(self._pending_status), []
            self._raise_on_status(st)

    def _raise_on_status(self, status: torch.Tensor):
        bad = int(status.max().item())
        if bad & _lib.AGH_DECODE_STUCK:
            raise AssertionError('decode did not terminate within 2N-1 steps: an instance has unvisited nodes that '
                                 'are all infeasible (the reference loops forever on such an input)')
        if bad & _lib.AGH_TOUR_INVALID:
            raise AssertionError('Invalid tour')
        if bad & _lib.AGH_CAPACITY_EXCEEDED:
            raise AssertionError('Used more than capacity')
        if bad & _lib.AGH_TW_VIOLATED:
            raise AssertionError('time window violated')

    # ------------------------------------------------------------------ embeddings
    def _init_embed(self, input):
        return ops.init_embed(self.weight_blob(), self._as_task(input))

    @_on_model_device
    def embed(self, input) -> torch.Tensor:
        """a2+a3: node embeddings [B,n,128] (no grad).  Eval mode: running-statistics BatchNorm folded into the GEMM
        epilogues; train mode: batch statistics (and the running statistics are updated, as a torch forward would)."""
        task = self._as_task(input)
        with torch.no_grad():
            if self.training:
                from .train_fn import encoder_train_forward
                h, _ = encoder_train_forward(self, self.weight_blob(), task, self._param_dict(), self._buffer_dict(), save=False)
                return h.view(task.B, task.N + 1, -1)
            return ops.encode(self.weight_blob(), task)

    # ------------------------------------------------------------------ forward
    @_on_model_device
    def forward(self, input, return_pi=False):
        """input: the reference's fleet-batch dict (or an ops.FleetTask).  Returns (cost, ll, serve_time[, pi])."""
        task = self._as_task(input)
        need_grad = torch.is_grad_enabled() and self.training
        if need_grad:
            from .train_fn import FleetTrainFn
            named = self._param_dict()
            ll, cost, serve, pi16, steps, status = FleetTrainFn.apply(self, task, list(named.keys()), *named.values())
            out = {'pi': pi16, 'steps': steps}
        else:
            with torch.no_grad():
                blob = self.weight_blob()
                h = self.embed(task)
                keys, psc, qbase = ops.precompute(blob, h.contiguous(), task)
                out = ops.decode(blob, self.distance_table(), task, keys, psc, qbase, mode=self._mode(), temp=self.temp,
                                 seed=self._seed() if self.decode_type == 'sampling' else 0, id_offset=self.id_offset)
            ll, cost, serve, status = out['ll'], out['cost'], out['serve'], out['status']
        self.last_steps = out['steps']
        # status carries AGH_DECODE_STUCK: the step budget ran out with unvisited nodes
        if self.check_tours:
            # problem.get_costs' asserts (attention_model.py:183) on the zero-padded int16 tours: no host sync
            # needed for the width; the status sync happens here, or once per batch when defer_checks is set
            status = status | ops.get_costs(out['pi'], task, self.distance_table())[1]
        if self.defer_checks:
            self._pending_status.append(status.max())
        else:
            self._raise_on_status(status)
        if return_pi:
            T = int(out['steps'].max().item())
            return cost, ll, serve, out['pi'][:, :T].long()
        return cost, ll, serve

    # ------------------------------------------------------------------ reference-internal API kept for callers
    @_on_model_device
    def _inner(self, input, embeddings):
        """(log_p [B,T,n], pi [B,T], serve_time) like attention_model.py:298-356 (materialises log_p: debugging only)."""
        task = self._as_task(input)
        with torch.no_grad():
            blob = self.weight_blob()
            keys, psc, qbase = ops.precompute(blob, embeddings.detach().contiguous(), task)
            out = ops.decode(blob, self.distance_table(), task, keys, psc, qbase, mode=self._mode(), temp=self.temp,
                             seed=self._seed() if self.decode_type == 'sampling' else 0, id_offset=self.id_offset,
                             want_logp_full=True)
        T = int(out['steps'].max().item())
        log_p = out['logp_full'][:, :T]
        # steps after an instance finished: only the depot is feasible, log_p(depot) = 0
        done = torch.arange(T, device=log_p.device)[None, :] >= out['steps'][:, None]
        pad = torch.full_like(log_p[0, 0], -math.inf)
        pad[0] = 0.0
        log_p = torch.where(done[:, :, None], pad, log_p)
        return log_p, out['pi'][:, :T].long(), out['serve']

    def _calc_log_likelihood(self, _log_p, a, mask):
        log_p = _log_p.gather(2, a.unsqueeze(-1)).squeeze(-1)
        if mask is not None:
            log_p[mask] = 0
        assert (log_p > -1000).data.all(), 'Logprobs should not be -inf, check sampling procedure!'
        return log_p.sum(1)

    @_on_model_device
    def sample_many(self, input, batch_rep=1, iter_rep=1):
        """Best of batch_rep*iter_rep rollouts per instance (attention_model.py:358-370,
        utils/functions.py:182-223).  The encoder and the key projection run once per instance; the
        batch_rep replicas share the staged keys (key_mod) instead of being materialised by do_batch_rep."""
        task = self._as_task(input)
        B = task.B
        with torch.no_grad():
            blob = self.weight_blob()
            h = self.embed(task)
            keys, psc, qbase = ops.precompute(blob, h.contiguous(), task)
            costs, pis, serves, stuck = [], [], [], []
            for it in range(iter_rep):
                out = ops.decode(blob, self.distance_table(), task, keys, psc, qbase, mode=self._mode(), temp=self.temp,
                                 seed=self._seed() if self.decode_type == 'sampling' else 0,
                                 id_offset=self.id_offset, reps=batch_rep)
                costs.append(out['cost'].view(batch_rep, B)); pis.append(out['pi']); serves.append(out['serve'])
                stuck.append(out['status'].max())
            R = batch_rep * iter_rep
            if R == 1:
                mc, mp, ms = costs[0].view(B), pis[0], serves[0]
            else:
                mc, mp, ms, _ = ops.best_of(torch.cat(costs, 0).contiguous(), torch.cat(pis, 0).contiguous(),
                                            torch.cat(serves, 0).contiguous(), R, B)
            T = int((mp != 0).int().cumsum(1).argmax(1).max().item()) + 1 if mp.numel() else 0
            pi = mp[:, :max(T, 1)].long()
            status = torch.stack(stuck).max()
            if self.check_tours:
                status = torch.maximum(status, ops.get_costs(mp, task, self.distance_table())[1].max())
            if self.defer_checks:
                self._pending_status.append(status)
            else:
                self._raise_on_status(status)
        return pi, mc, ms
