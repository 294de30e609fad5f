"""CPU oracle for the AGH attention-model rollout -- TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement of the reference algorithm for the hot path in
SURVEY.md section 8 (rows a1-a20).  It exists to CHECK the CUDA path and to
serve as the timed CPU baseline of bench.py.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it; nothing in
agh_b200/ does, and the product path raises when its CUDA library is missing.

Parity status: PINNED.  tests/test_oracle_golden.py checks every function here
against fixtures produced by the unmodified reference (tests/golden/make_golden.py):
per-step masks / log-probs / state scalars, tours, costs, log-likelihoods,
serve times, sampled-action replay, one REINFORCE batch (loss + gradients) and
the env-only heuristic KATs.  The arithmetic the reference executes lives in
PyTorch ATen (third party, pinned by the reference as pytorch=1.8.1,
environment.yml:257; torch 2.11 in this image); the oracle issues the same ATen
contractions in the same order so that greedy tours reproduce bit-for-bit.

Citations are to /root/reference/Construction_based (CB/).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

NODE_SIZE = 92          # state_agh.py:33
SPEED = 110.0           # state_agh.py:32
CAPACITY = 1.0          # state_agh.py:31
N_HEADS = 8
CAPACITIES = {10: 20., 20: 30., 50: 40., 100: 50., 200: 60., 300: 70.}   # generate_data.py:9
STAY = (30, 34, 33)                                                     # generate_data.py:14

_TABLES = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'agh_b200', 'problems', 'agh',
                       'agh_tables.npz')


@dataclass
class Tables:
    """fleet_info.pkl / distance.pkl / arrival_prob.npy (attention_model.py:78-88, problem_agh.py:105)."""
    order: list
    precedence: dict
    duration: torch.Tensor            # [11,3] f64, row = fleet id
    next_duration: torch.Tensor       # [5,3]  f64, row = precedence level
    nd_is_f32: list                   # level 4 is a python-float list -> f32 tensor (train.py:44-46)
    distance: torch.Tensor            # [8464] f32, row-major (i,j) -> 92*i+j
    arrival_prob: np.ndarray

    @staticmethod
    def load(path: str = _TABLES) -> 'Tables':
        z = np.load(path)
        return Tables(order=[int(x) for x in z['order']],
                      precedence={f: int(z['precedence'][f]) for f in range(1, 11)},
                      duration=torch.from_numpy(z['duration']),
                      next_duration=torch.from_numpy(z['next_duration']),
                      nd_is_f32=[bool(x) for x in z['next_duration_is_f32']],
                      distance=torch.from_numpy(z['distance']),
                      arrival_prob=z['arrival_prob'])


# --------------------------------------------------------------------------------------
# instance generator (generate_data.py:8-18 == problem_agh.py:104-113)
# --------------------------------------------------------------------------------------
def generate_instances(num: int, n: int, prob: np.ndarray, fleet_size: int = 10, rng=np.random) -> Dict[str, torch.Tensor]:
    """Same draws, in the same order, from the same (global) numpy RNG as the reference."""
    loc = 1 + rng.choice(91, size=(num, n))
    arrival = 60 * rng.choice(24, size=(num, n), p=prob) + rng.randint(0, 60, size=(num, n))
    type_ = rng.randint(0, 3, size=(num, n))
    departure = arrival + np.asarray(STAY)[type_]
    demand = rng.randint(1, 10, size=(num, fleet_size, n)) / CAPACITIES[n]
    # make_instance (problem_agh.py:82-90): python lists -> torch tensors of these dtypes
    return {'loc': torch.from_numpy(loc.astype(np.int64)),
            'arrival': torch.from_numpy(arrival.astype(np.float32)),
            'departure': torch.from_numpy(departure.astype(np.float32)),
            'type': torch.from_numpy(type_.astype(np.int64)),
            'demand': torch.tensor(demand.tolist(), dtype=torch.float)}


# --------------------------------------------------------------------------------------
# a1: fleet-batch assembly (train.py:40-57 == train.py:167-184 == eval.py:97-113)
# --------------------------------------------------------------------------------------
def new_tw_left_levels(inst) -> torch.Tensor:
    return inst['arrival'].repeat(6, 1, 1)              # train.py:40


def fleet_task(tb: Tables, inst, f: int, tw_left_levels: torch.Tensor) -> Dict[str, torch.Tensor]:
    p = tb.precedence[f]
    B = inst['loc'].size(0)
    nd = tb.next_duration[p]
    if tb.nd_is_f32[p]:
        nd = nd.float()
    tw_right = inst['departure'] - nd[inst['type']]                       # f32 - f64 -> f64 (f32 at level 4)
    tw_right = torch.cat((torch.full_like(tw_right[:, :1], 1441), tw_right), 1)
    tw_left = tw_left_levels[p]
    tw_left = torch.cat((torch.zeros_like(tw_left[:, :1]), tw_left), 1)
    return {'loc': inst['loc'], 'demand': inst['demand'][:, f - 1, :],
            'distance': tb.distance.expand(B, tb.distance.numel()),
            'duration': tb.duration[f][inst['type']],                   # f64 [B,N]
            'tw_right': tw_right, 'tw_left': tw_left,
            'fleet': torch.full((B, 1), f - 1)}


def chain_tw_left(tb: Tables, tw_left_levels: torch.Tensor, f: int, serve_time: torch.Tensor) -> None:
    p = tb.precedence[f]
    tw_left_levels[p + 1] = torch.max(tw_left_levels[p + 1], serve_time[:, 1:])     # train.py:67


# --------------------------------------------------------------------------------------
# a2-a4: embedding, encoder, precompute
# --------------------------------------------------------------------------------------
def init_embed(W, task) -> torch.Tensor:
    """attention_model.py:272-294."""
    loc0 = torch.cat((torch.zeros_like(task['loc'][:, :1]), task['loc']), 1)
    x = W['loc_embedding.weight'][loc0]
    feat = torch.cat((task['demand'][:, :, None], task['tw_left'][:, 1:, None] / 1440,
                      task['tw_right'][:, 1:, None] / 1440), -1).float()
    x[:, 1:, :] = x[:, 1:, :] + F.linear(feat, W['init_embed.weight'], W['init_embed.bias'])
    return x


def _batch_norm(W, prefix, x, training, stats_out):
    """graph_encoder.py:135-138: BatchNorm1d over the flattened B*n rows."""
    flat = x.reshape(-1, x.size(-1))
    if training:
        y = F.batch_norm(flat, None, None, W[prefix + '.weight'], W[prefix + '.bias'], True, 0.1, 1e-5)
        if stats_out is not None:
            stats_out[prefix] = (flat.detach().mean(0), flat.detach().var(0, unbiased=True))
    else:
        y = F.batch_norm(flat, W[prefix + '.running_mean'], W[prefix + '.running_var'],
                         W[prefix + '.weight'], W[prefix + '.bias'], False, 0.1, 1e-5)
    return y.view_as(x)


def encoder(W, x: torch.Tensor, training: bool = False, n_layers: int = 3, stats_out=None) -> torch.Tensor:
    """graph_encoder.py:56-111,146-172,195-207 (3 x [MHA+skip, BN, FF+skip, BN])."""
    B, n, D = x.shape
    for L in range(n_layers):
        pre = 'embedder.layers.%d.' % L
        flat = x.contiguous().view(-1, D)
        shp = (N_HEADS, B, n, -1)
        Q = torch.matmul(flat, W[pre + '0.module.W_query']).view(shp)
        K = torch.matmul(flat, W[pre + '0.module.W_key']).view(shp)
        V = torch.matmul(flat, W[pre + '0.module.W_val']).view(shp)
        attn = torch.softmax((1 / math.sqrt(D // N_HEADS)) * torch.matmul(Q, K.transpose(2, 3)), dim=-1)
        heads = torch.matmul(attn, V)
        out = torch.mm(heads.permute(1, 2, 0, 3).contiguous().view(-1, D),
                       W[pre + '0.module.W_out'].view(-1, D)).view(B, n, D)
        x = _batch_norm(W, pre + '1.normalizer', x + out, training, stats_out)
        ff = F.linear(F.relu(F.linear(x, W[pre + '2.module.0.weight'], W[pre + '2.module.0.bias'])),
                      W[pre + '2.module.2.weight'], W[pre + '2.module.2.bias'])
        x = _batch_norm(W, pre + '3.normalizer', x + ff, training, stats_out)
    return x


def precompute(W, h: torch.Tensor, fleet: torch.Tensor):
    """attention_model.py:392-414,604-611.  Returns (fixed_ctx, fleet_emb, gK, gV, LK) with
    gK/gV as the reference's permuted head views (8,B,1,n,16)."""
    B, n, D = h.shape
    fixed_ctx = F.linear(h.mean(1), W['project_fixed_context.weight'])[:, None, :]
    fleet_emb = W['fleets_embedding.weight'][fleet]                   # [B,1,D]
    gk, gv, lk = F.linear(h[:, None, :, :], W['project_node_embeddings.weight']).chunk(3, dim=-1)

    def heads(v):
        return v.contiguous().view(B, 1, n, N_HEADS, -1).permute(3, 0, 1, 2, 4)
    return fixed_ctx, fleet_emb, heads(gk), heads(gv), lk.contiguous()


# --------------------------------------------------------------------------------------
# a5, a8, a12: environment
# --------------------------------------------------------------------------------------
class Env:
    """Plain-tensor restatement of StateAGH (state_agh.py:60-174).

    cur_free_time is held in f64 from the start.  The reference starts it as int64 -60
    (state_agh.py:89) so its step-0 arithmetic is f32; that is unobservable because on a fresh
    vehicle max(cft + travel, tw_left) == tw_left (travel <= 22.75 < 60 <= 60 + tw_left) and
    -60/1440 rounds to the same f32 either way (checked against the golden step dumps).
    """

    def __init__(self, task):
        loc = task['loc']
        B, N = loc.shape
        self.B, self.N = B, N
        self.coords = torch.cat((torch.zeros_like(loc[:, :1]), loc), 1)              # [B,n] i64
        self.demand = task['demand']                                                  # [B,N] f32
        self.tw_left, self.tw_right = task['tw_left'], task['tw_right']
        self.distance = task['distance']
        self.duration = torch.cat((torch.zeros_like(loc[:, :1]), task['duration']), 1)  # i64 0 ++ f64 -> f64
        self.prev = torch.zeros(B, 1, dtype=torch.long)
        self.used = self.demand.new_zeros(B, 1)
        self.visited = torch.zeros(B, N + 1, dtype=torch.uint8)
        self.lengths = torch.zeros(B, 1)
        self.cft = torch.full((B, 1), -60.0, dtype=torch.float64)
        self.serve = torch.zeros(B, N + 1)
        self.i = 0

    def travel(self, to_coord):
        """f32 table lookup and the f32 IEEE division by SPEED (state_agh.py:118,167)."""
        idx = NODE_SIZE * self.coords.gather(1, self.prev) + to_coord
        d = self.distance.gather(1, idx)
        return d, d / SPEED

    def mask(self) -> torch.Tensor:
        """state_agh.py:148-174.  [B,n] bool, True = infeasible."""
        cap = (self.demand + self.used) > CAPACITY
        _, tt = self.travel(self.coords[:, 1:])
        fin = torch.max(self.cft + tt, self.tw_left[:, 1:]) + self.duration[:, 1:]
        tw = fin > self.tw_right[:, 1:]
        m_loc = self.visited[:, 1:].bool() | cap | tw
        m_dep = (self.prev == 0) & ((m_loc == 0).int().sum(-1, keepdim=True) > 0)
        return torch.cat((m_dep, m_loc), -1)

    def update(self, sel: torch.Tensor) -> None:
        """state_agh.py:99-137."""
        sel = sel[:, None]
        d, tt = self.travel(self.coords.gather(1, sel))
        self.lengths = self.lengths + d
        nz = (sel != 0).float()
        dem = self.demand.gather(1, torch.clamp(sel - 1, 0, self.N - 1))
        self.used = (self.used + dem) * nz
        self.cft = (torch.max(self.cft + tt, self.tw_left.gather(1, sel)) + self.duration.gather(1, sel)) * nz \
            - 60 * (sel == 0).float()
        self.visited = self.visited.scatter(1, sel, 1)
        self.serve = self.serve.scatter(1, sel, self.cft.float())
        self.prev = sel
        self.i += 1

    def all_finished(self) -> bool:
        return self.i >= self.N and bool(self.visited.all())                 # state_agh.py:139-140

    def finished(self) -> torch.Tensor:
        return self.visited.sum(-1) == self.visited.size(-1)                 # state_agh.py:142-143


# --------------------------------------------------------------------------------------
# a6-a11, a14: decode loop
# --------------------------------------------------------------------------------------
def step_log_p(W, h, fixed, env: Env, temp: float = 1.0, inject_logits: Optional[torch.Tensor] = None,
               logits_out: Optional[list] = None):
    """attention_model.py:429-451,510-522,550-584 for one step.  Returns (log_p [B,n], mask [B,n]).

    inject_logits [B,n]: replace the clipped logits (attention_model.py:580, before the mask and the
    temperature) by these values -- the "given identical logits" experiments; logits_out collects the
    clipped logits of the step."""
    fixed_ctx, fleet_emb, gK, gV, LK = fixed
    B, n, D = h.shape
    ctx = torch.cat((h.gather(1, env.prev[:, :, None].expand(B, 1, D)),
                     CAPACITY - env.used[:, :, None], env.cft[:, :, None] / 1440), -1).float().to(h.dtype)
    query = fixed_ctx + F.linear(ctx, W['project_step_context.weight']) + fleet_emb
    mask = env.mask()
    gQ = query.view(B, 1, N_HEADS, 1, D // N_HEADS).permute(2, 0, 1, 3, 4)
    compat = torch.matmul(gQ, gK.transpose(-2, -1)) / math.sqrt(gQ.size(-1))
    compat[mask[None, :, None, None, :].expand_as(compat)] = -math.inf
    heads = torch.matmul(torch.softmax(compat, dim=-1), gV)
    glimpse = F.linear(heads.permute(1, 2, 3, 0, 4).contiguous().view(-1, 1, 1, D), W['project_out.weight'])
    logits = torch.matmul(glimpse, LK.transpose(-2, -1)).squeeze(-2) / math.sqrt(D)
    logits = torch.tanh(logits) * 10.0
    if logits_out is not None:
        logits_out.append(logits[:, 0, :].detach().clone())
    if inject_logits is not None:
        logits = inject_logits[:, None, :].clone()
    logits[mask[:, None, :]] = -math.inf
    return torch.log_softmax(logits / temp, dim=-1)[:, 0, :], mask


def decode(W, task, h, mode: str = 'greedy', actions: Optional[torch.Tensor] = None, temp: float = 1.0,
           record: Optional[dict] = None, inject_log_p: Optional[torch.Tensor] = None,
           inject_logits: Optional[torch.Tensor] = None):
    """attention_model.py:298-356.  mode: 'greedy' | 'sampling' | 'replay' (follow `actions` [B,T]).

    inject_log_p [B,T,n]: use these log-probs for SELECTION instead of the computed ones
    ("identical logits" experiments); inject_logits [B,T,n]: replace the clipped logits of step t, the
    log-softmax / exp / argmax that follow are the reference's.  record['logits'] collects the clipped logits.
    Returns dict(pi, ll, serve, lengths, logp_sel [B,T]).
    """
    env = Env(task)
    fixed = precompute(W, h, task['fleet'])
    pis, lps = [], []
    t = 0
    while True:
        if mode == 'replay':
            if t >= actions.size(1):
                break
        elif env.all_finished():
            break
        log_p, mask = step_log_p(W, h, fixed, env, temp,
                                 inject_logits=None if inject_logits is None else inject_logits[:, t],
                                 logits_out=None if record is None else record.setdefault('logits', []))
        if record is not None:
            record.setdefault('log_p', []).append(log_p.detach())
            record.setdefault('mask', []).append(mask)
            record.setdefault('used', []).append(env.used[:, 0].clone())
            record.setdefault('cft', []).append(env.cft[:, 0].clone())
            record.setdefault('prev', []).append(env.prev[:, 0].clone())
        sel_src = inject_log_p[:, t] if inject_log_p is not None else log_p
        if mode == 'greedy':
            sel = sel_src.exp().max(1)[1]                                    # attention_model.py:333,377
        elif mode == 'sampling':
            sel = sel_src.exp().multinomial(1).squeeze(1)                    # attention_model.py:381
        else:
            sel = actions[:, t].long()
        assert not mask.gather(1, sel[:, None]).any(), 'infeasible action selected'
        env.update(sel)
        pis.append(sel)
        lps.append(log_p.gather(1, sel[:, None]).squeeze(1))
        t += 1
    pi = torch.stack(pis, 1)
    logp_sel = torch.stack(lps, 1)
    return {'pi': pi, 'll': logp_sel.sum(1), 'serve': env.serve, 'lengths': env.lengths[:, 0],
            'logp_sel': logp_sel, 'env': env}


def get_costs(task, pi: torch.Tensor) -> torch.Tensor:
    """problem_agh.py:18-66: validity asserts + closed-tour length (torch.sum over T+1 f32 terms)."""
    B, N = task['demand'].shape
    sp = pi.sort(1)[0]
    assert (torch.arange(1, N + 1).view(1, -1).expand(B, N) == sp[:, -N:]).all() and (sp[:, :-N] == 0).all(), \
        'Invalid tour'
    dwd = torch.cat((torch.full_like(task['demand'][:, :1], -CAPACITY), task['demand']), 1).gather(1, pi)
    used = torch.zeros_like(task['demand'][:, 0])
    for i in range(pi.size(1)):
        used = used + dwd[:, i]
        used[used < 0] = 0
        assert (used <= CAPACITY + 1e-5).all(), 'Used more than capacity'
    loc = torch.cat((torch.zeros_like(task['loc'][:, :1]), task['loc']), 1).gather(1, pi)
    z = torch.zeros_like(loc[:, :1])
    idx = NODE_SIZE * torch.cat((z, loc), 1) + torch.cat((loc, z), 1)
    tt = (task['distance'] / SPEED).gather(1, idx)
    cur = torch.full_like(pi[:, 0:1], -60)
    dur = torch.cat((torch.zeros_like(task['duration'][:, :1]), task['duration']), 1)
    for i in range(pi.size(1)):
        a = pi[:, i:i + 1]
        cur = (torch.max(cur + tt[:, i:i + 1], task['tw_left'].gather(1, a)) + dur.gather(1, a)) * (a != 0).float() \
            - 60 * (a == 0).float()
        assert (cur <= task['tw_right'].gather(1, a) + 1e-5).all(), 'time window violated'
    return task['distance'].gather(1, idx).sum(1)


def forward(W, task, mode='greedy', training=False, actions=None, temp=1.0, record=None, stats_out=None):
    """a15: AttentionModel.forward (attention_model.py:159-196) -> dict(cost, ll, serve, pi)."""
    h = encoder(W, init_embed(W, task), training=training, stats_out=stats_out)
    out = decode(W, task, h, mode=mode, actions=actions, temp=temp, record=record)
    out['cost'] = get_costs(task, out['pi'])
    out['h'] = h
    return out


def rollout(W, tb: Tables, inst, mode='greedy', actions=None, collect=None, temp=1.0):
    """a17: train.py:31-70 greedy 10-fleet loop (eval mode).  Returns per-fleet costs [B,10]."""
    lv = new_tw_left_levels(inst)
    costs = []
    with torch.no_grad():
        for k, f in enumerate(tb.order):
            task = fleet_task(tb, inst, f, lv)
            out = forward(W, task, mode=mode, actions=None if actions is None else actions[k], temp=temp)
            costs.append(out['cost'].view(-1, 1))
            if collect is not None:
                collect.append({'pi': out['pi'], 'll': out['ll'], 'serve': out['serve'], 'h': out['h'], 'task': task})
            chain_tw_left(tb, lv, f, out['serve'])
    return torch.cat(costs, 1)


def sample_many(W, task, batch_rep: int, iter_rep: int, temp: float = 1.0):
    """a16: attention_model.py:358-370 + utils/functions.py:171-223 (best of batch_rep*iter_rep)."""
    h = encoder(W, init_embed(W, task))
    rep = lambda v: v[None, ...].expand(batch_rep, *v.size()).contiguous().view(-1, *v.size()[1:])
    task_r = {k: rep(v) for k, v in task.items()}
    h_r = rep(h)
    costs, pis, serves = [], [], []
    for _ in range(iter_rep):
        out = decode(W, task_r, h_r, mode='sampling', temp=temp)
        c = get_costs(task_r, out['pi'])
        costs.append(c.view(batch_rep, -1).t())
        pis.append(out['pi'].view(batch_rep, -1, out['pi'].size(-1)).transpose(0, 1))
        serves.append(out['serve'].view(batch_rep, -1, out['serve'].size(-1)).transpose(0, 1))
    L = max(p.size(-1) for p in pis)
    pis = torch.cat([F.pad(p, (0, L - p.size(-1))) for p in pis], 1)
    costs, serves = torch.cat(costs, 1), torch.cat(serves, 1)
    mc, am = costs.min(-1)
    ar = torch.arange(pis.size(0))
    return pis[ar, am], mc, serves[ar, am]


def reinforce_loss(W, tb: Tables, inst, bl_val: torch.Tensor, actions, stats_out=None):
    """a18: train.py:159-209 with the sampled actions replayed (train-mode BatchNorm, autograd on W)."""
    lv = new_tw_left_levels(inst)
    loss = 0.0
    outs = []
    for k, f in enumerate(tb.order):
        task = fleet_task(tb, inst, f, lv)
        so = {} if stats_out is not None else None
        out = forward(W, task, mode='replay', training=True, actions=actions[k], stats_out=so)
        if stats_out is not None:
            stats_out.append(so)
        loss = loss + ((out['cost'] - bl_val[:, k]) * out['ll']).mean()
        chain_tw_left(tb, lv, f, out['serve'].detach())
        outs.append(out)
    return loss / len(tb.order), outs


# --------------------------------------------------------------------------------------
# model-free heuristic used as an env KAT (agh_baseline.py:80-101)
# --------------------------------------------------------------------------------------
def nearest_neighbor(task):
    env = Env(task)
    seq = []
    while not env.all_finished():
        m = env.mask()
        d, _ = env.travel(env.coords)
        d = d.clone()
        d[m] = 10000
        sel = d.min(1)[1]
        env.update(sel)
        seq.append(sel)
    pi = torch.stack(seq, 1)
    return get_costs(task, pi), env.serve, pi


def cws(task):
    """Clarke-Wright savings heuristic, agh_baseline.py:17-77: second weight-free consumer of the environment
    (known answer on the shipped agh20 set: 106354.4296875).  savings f64 = (d(0,j) + d(i,0) - d(i,j)) [f32]
    - 0.03*60 * (tw_left[j] - tw_left[i] - duration[i] - d(i,j)/110 [f32]); at the depot the reference indexes the
    savings with prev_a - 1 = -1, the row of the last flight."""
    loc = task['loc']
    B, N = loc.shape
    table = task['distance'][0]
    d_ij = table[NODE_SIZE * loc[:, :, None] + loc[:, None, :]]
    savings_distance = table[loc][:, None, :] + table[NODE_SIZE * loc][:, :, None] - d_ij
    tw_left, tw_right = task['tw_left'][:, 1:], task['tw_right'][:, 1:]
    savings_time = tw_left[:, None, :] - tw_left[:, :, None]
    savings_time = savings_time - task['duration'][:, :, None] - d_ij / 110.0
    savings = savings_distance - 0.03 * 60 * savings_time
    env = Env(task)
    sel = 1 + tw_right.sort(dim=1)[1][:, 0]
    env.update(sel)
    seq = [sel]
    rows = torch.arange(B)
    while not env.all_finished():
        score = savings[rows, env.prev[:, 0] - 1]
        score = torch.cat((score.min(1, keepdim=True)[0] - 1, score), 1)
        score[env.mask().bool()] = -math.inf
        sel = score.sort(descending=True)[1][:, 0]
        env.update(sel)
        seq.append(sel)
    pi = torch.stack(seq, 1)
    return get_costs(task, pi), env.serve, pi


def random_init_weights(seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Same parameter creation order and initialisers as AttentionModel.__init__
    (attention_model.py:115-152, graph_encoder.py:42-54,124,165-168) so that torch.manual_seed(seed)
    reproduces the reference's random-init state_dict (sha pinned in tests/golden/meta.json)."""
    from torch import nn
    torch.manual_seed(seed)
    W = {}
    W['fleets_embedding.weight'] = nn.Embedding(10, 128).weight.data
    W['loc_embedding.weight'] = nn.Embedding(92, 128).weight.data
    lin = nn.Linear(3, 128)
    W['init_embed.weight'], W['init_embed.bias'] = lin.weight.data, lin.bias.data
    for L in range(3):
        pre = 'embedder.layers.%d.' % L
        for name, shape in (('W_query', (8, 128, 16)), ('W_key', (8, 128, 16)), ('W_val', (8, 128, 16)),
                            ('W_out', (8, 16, 128))):
            W[pre + '0.module.' + name] = torch.empty(shape).uniform_(-1 / math.sqrt(shape[-1]), 1 / math.sqrt(shape[-1]))
        for idx in ('1', '3'):
            if idx == '3':
                l1, l2 = nn.Linear(128, 512), nn.Linear(512, 128)
                W[pre + '2.module.0.weight'], W[pre + '2.module.0.bias'] = l1.weight.data, l1.bias.data
                W[pre + '2.module.2.weight'], W[pre + '2.module.2.bias'] = l2.weight.data, l2.bias.data
            W[pre + idx + '.normalizer.weight'] = torch.ones(128)
            W[pre + idx + '.normalizer.bias'] = torch.zeros(128)
            W[pre + idx + '.normalizer.running_mean'] = torch.zeros(128)
            W[pre + idx + '.normalizer.running_var'] = torch.ones(128)
            W[pre + idx + '.normalizer.num_batches_tracked'] = torch.tensor(0)
    W['project_node_embeddings.weight'] = nn.Linear(128, 384, bias=False).weight.data
    W['project_fixed_context.weight'] = nn.Linear(128, 128, bias=False).weight.data
    W['project_step_context.weight'] = nn.Linear(130, 128, bias=False).weight.data
    W['project_out.weight'] = nn.Linear(128, 128, bias=False).weight.data
    return W
