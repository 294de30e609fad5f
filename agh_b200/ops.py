"""Thin host wrappers over the C ABI: allocate outputs with torch, pass raw pointers + stream.

Nothing here computes; every function is one (or a fixed few) calls into libagh_b200.so on the
current CUDA stream.  CPU tensors are rejected (agh_b200._lib.ptr raises): no fallback path.
"""
from __future__ import annotations

import ctypes as C
import ctypes as _ct
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr

D, KS = 128, _lib.KEY_STRIDE


@dataclass
class FleetTask:
    """Node vectors of one fleet sub-problem in the layout the kernels read."""
    coord: torch.Tensor      # [B,n] u8, coord[:,0] = 0 (depot)
    demand: torch.Tensor     # [B,N] f32
    tw_left: torch.Tensor    # [B,n] f32
    tw_right: torch.Tensor   # [B,n] f64
    duration: torch.Tensor   # [B,n] f64, duration[:,0] = 0
    twr_f32: bool            # tw_right carried f32 values at the caller (fleet 10)
    fleet: int               # 0..9, or -1 when fleet_ids is given
    fleet_ids: Optional[torch.Tensor] = None   # [B] i32

    @property
    def B(self) -> int:
        return self.coord.size(0)

    @property
    def N(self) -> int:
        return self.demand.size(1)

    @staticmethod
    def from_input(inp: Dict[str, torch.Tensor]) -> 'FleetTask':
        """Accept the reference's fleet-batch dict (train.py:53-57): loc i64 [B,N], demand f32 [B,N],
        duration f64 [B,N], tw_right f64|f32 [B,n], tw_left f32 [B,n], fleet i64 [B,1].  'distance' is ignored:
        the table is model state, not a per-instance stride-0 tensor."""
        loc = inp['loc']
        B, N = loc.shape
        coord = torch.zeros(B, N + 1, dtype=torch.uint8, device=loc.device)
        coord[:, 1:] = loc.to(torch.uint8)
        duration = torch.zeros(B, N + 1, dtype=torch.float64, device=loc.device)
        duration[:, 1:] = inp['duration'].to(torch.float64)
        fl = inp['fleet'].reshape(-1)
        # a batch normally holds one fleet; keep per-instance ids for generality without a host sync
        return FleetTask(coord=coord, demand=inp['demand'].to(torch.float32).contiguous(),
                         tw_left=inp['tw_left'].to(torch.float32).contiguous(),
                         tw_right=inp['tw_right'].to(torch.float64).contiguous(), duration=duration,
                         twr_f32=inp['tw_right'].dtype == torch.float32, fleet=-1,
                         fleet_ids=fl.to(torch.int32).contiguous())


def fleet_task(loc, type_, departure, demand_all, tw_left_levels, level: int, fleet: int, duration3, next_duration3,
               nd_is_f32: bool) -> FleetTask:
    """a1 fused (train.py:40-57): one launch builds every node vector of fleet `fleet` (1..10)."""
    B, N = loc.shape
    dev = loc.device
    coord = torch.empty(B, N + 1, dtype=torch.uint8, device=dev)
    demand = torch.empty(B, N, dtype=torch.float32, device=dev)
    tw_left = torch.empty(B, N + 1, dtype=torch.float32, device=dev)
    tw_right = torch.empty(B, N + 1, dtype=torch.float64, device=dev)
    duration = torch.empty(B, N + 1, dtype=torch.float64, device=dev)
    d3 = (C.c_double * 3)(*[float(x) for x in duration3])
    nd3 = (C.c_double * 3)(*[float(x) for x in next_duration3])
    check(lib().agh_fleet_task(ptr(loc, torch.int64), ptr(type_, torch.int64), ptr(departure, torch.float32),
                               ptr(demand_all, torch.float32), ptr(tw_left_levels, torch.float32), level, fleet,
                               C.cast(d3, C.c_void_p), C.cast(nd3, C.c_void_p), int(nd_is_f32), B, N, ptr(coord),
                               ptr(demand), ptr(tw_left), ptr(tw_right), ptr(duration), stream_ptr(dev)))
    return FleetTask(coord, demand, tw_left, tw_right, duration, bool(nd_is_f32), fleet - 1)


def cat_tasks(tasks) -> FleetTask:
    """Concatenate the sub-problems of several fleets of ONE precedence level along the batch (per-instance fleet
    ids; the tw_right dtype flag is a property of the level, so it is shared)."""
    assert len({t.twr_f32 for t in tasks}) == 1
    dev = tasks[0].coord.device
    ids = torch.cat([t.fleet_ids if t.fleet_ids is not None else
                     torch.full((t.B,), t.fleet, dtype=torch.int32, device=dev) for t in tasks])
    return FleetTask(coord=torch.cat([t.coord for t in tasks]), demand=torch.cat([t.demand for t in tasks]),
                     tw_left=torch.cat([t.tw_left for t in tasks]), tw_right=torch.cat([t.tw_right for t in tasks]),
                     duration=torch.cat([t.duration for t in tasks]), twr_f32=tasks[0].twr_f32, fleet=-1, fleet_ids=ids)


def chain_tw_left(tw_left_levels, level: int, serve) -> None:
    _, B, N = tw_left_levels.shape
    check(lib().agh_chain_tw_left(ptr(tw_left_levels, torch.float32), level, ptr(serve, torch.float32), B, N,
                                  stream_ptr(serve.device)))


def init_embed(wblob, task: FleetTask) -> torch.Tensor:
    B, N = task.B, task.N
    h0 = torch.empty(B, N + 1, D, dtype=torch.float32, device=wblob.device)
    check(lib().agh_init_embed(ptr(wblob, torch.float32), ptr(task.coord), ptr(task.demand), ptr(task.tw_left),
                               ptr(task.tw_right), int(task.twr_f32), B, N, ptr(h0), stream_ptr(wblob.device)))
    return h0


def _workspace(B, N, dev):
    nbytes = lib().agh_encoder_workspace_bytes(B, N)
    return torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=dev)


def encode(wblob, task: FleetTask) -> torch.Tensor:
    """a2+a3: init embedding + 3 encoder layers (eval-mode BatchNorm)."""
    B, N = task.B, task.N
    dev = wblob.device
    h = torch.empty(B, N + 1, D, dtype=torch.float32, device=dev)
    ws = _workspace(B, N, dev)
    check(lib().agh_encode(ptr(wblob, torch.float32), ptr(task.coord), ptr(task.demand), ptr(task.tw_left),
                           ptr(task.tw_right), int(task.twr_f32), B, N, ptr(h), ptr(ws), stream_ptr(dev)))
    return h


def encode_layers(wblob, h0) -> torch.Tensor:
    B, n, _ = h0.shape
    h = torch.empty_like(h0)
    ws = _workspace(B, n - 1, h0.device)
    check(lib().agh_encode_layers(ptr(wblob, torch.float32), ptr(h0, torch.float32), B, n - 1, ptr(h), ptr(ws),
                                  stream_ptr(h0.device)))
    return h


def encoder_mha(qkv, B: int, N: int) -> torch.Tensor:
    """Self-attention of one encoder layer alone: qkv [3, B*(N+1), 128] (Q, K, V stacked) -> heads [B*(N+1), 128] (graph_encoder.py:83-104)."""
    qkv = qkv.contiguous()
    heads = torch.empty(B * (N + 1), D, dtype=torch.float32, device=qkv.device)
    check(lib().agh_encoder_mha(ptr(qkv, torch.float32), int(B), int(N), ptr(heads), stream_ptr(qkv.device)))
    return heads


def encoder_ffn(wblob, layer: int, x) -> torch.Tensor:
    """Feed-forward sublayer of encoder layer `layer` alone: BN2(x + W2 relu(W1 x + b1) + b2), x [M,128] (graph_encoder.py:159-172)."""
    x = x.contiguous()
    M = x.shape[0]
    y = torch.empty_like(x)
    ws = torch.empty(int(lib().agh_encoder_ffn_workspace_bytes(M)) // 4 + 1, dtype=torch.float32, device=x.device)
    check(lib().agh_encoder_ffn(ptr(wblob, torch.float32), int(layer), ptr(x, torch.float32), M, ptr(y), ptr(ws), stream_ptr(x.device)))
    return y


def precompute(wblob, h, task: FleetTask):
    """a4: keys [3,B,n,128] (K', V, LK' as plain row-major matrices), psc [B,n,128], qbase [B,128]."""
    B, n, _ = h.shape
    dev = h.device
    keys = torch.empty(3, B, n, KS, dtype=torch.float32, device=dev)
    psc = torch.empty(B, n, D, dtype=torch.float32, device=dev)
    qbase = torch.empty(B, D, dtype=torch.float32, device=dev)
    check(lib().agh_precompute(ptr(wblob, torch.float32), ptr(h, torch.float32),
                               ptr(task.fleet_ids, torch.int32) if task.fleet_ids is not None else None,
                               max(task.fleet, 0), B, n - 1, ptr(keys), ptr(psc), ptr(qbase), stream_ptr(dev)))
    return keys, psc, qbase


def travel_table(distance: torch.Tensor) -> torch.Tensor:
    """distance / 110 as the reference computes it (state_agh.py:118,167: an f32 IEEE division), once per table: the
    decode kernel's mask path loads the quotient instead of dividing every step.  Computed on the CPU (correctly
    rounded f32 division) and kept ON the distance tensor object (with the version it was computed from), so a cache can
    never outlive or be confused with another table."""
    cached = getattr(distance, '_agh_travel', None)
    if cached is not None and cached[0] == distance._version and cached[1].device == distance.device:
        return cached[1]
    t = (distance.detach().cpu().to(torch.float32) / 110.0).to(distance.device).contiguous()
    distance._agh_travel = (distance._version, t)
    return t


def decode(wblob, distance, task: FleetTask, keys, psc, qbase, mode: int = _lib.AGH_GREEDY, temp: float = 1.0,
           seed: int = 0, id_offset: int = 0, actions: Optional[torch.Tensor] = None, replay_steps: int = 0,
           reps: int = 1, inject_logp: Optional[torch.Tensor] = None, inject_z: Optional[torch.Tensor] = None,
           want_logp_sel: bool = False,
           want_logp_full: bool = False, want_dumps: bool = False, want_train_dumps: bool = False,
           t_stride: Optional[int] = None, smem_mats: int = 0, shared_keys: bool = True):
    """a5-a12 + a14.  `reps` > 1 runs reps*B rollouts sharing the B key sets (sample_many): in sampling mode they take
    the shared-key kernel (one staged key set per CTA, one warp per replica; csrc/decode_shared.cu) unless
    shared_keys=False forces one CTA per replica."""
    if not shared_keys:
        smem_mats = int(smem_mats) | 64
    Bk, N = task.B, task.N
    n = N + 1
    B = Bk * reps
    dev = keys.device
    T = max(2 * N - 1, 2) if t_stride is None else t_stride      # N = 1: the node, then the depot
    out = {
        'pi': torch.empty(B, T, dtype=torch.int16, device=dev),
        'steps': torch.empty(B, dtype=torch.int32, device=dev),
        'status': torch.empty(B, dtype=torch.int32, device=dev),
        'll': torch.empty(B, dtype=torch.float32, device=dev),
        'cost': torch.empty(B, dtype=torch.float32, device=dev),
        'serve': torch.empty(B, n, dtype=torch.float32, device=dev),
    }
    if want_logp_sel:
        out['logp_sel'] = torch.empty(B, T, dtype=torch.float32, device=dev)
    if want_logp_full:
        out['logp_full'] = torch.zeros(B, T, n, dtype=torch.float32, device=dev)
    if want_dumps:
        out['mask_dump'] = torch.zeros(B, T, (n + 31) // 32, dtype=torch.int32, device=dev)
        out['used_dump'] = torch.zeros(B, T, dtype=torch.float32, device=dev)
        out['cft_dump'] = torch.zeros(B, T, dtype=torch.float64, device=dev)
    if want_train_dumps:       # what the REINFORCE backward differentiates (nets/train_fn.py)
        out['q_dump'] = torch.zeros(B, T, D, dtype=torch.float32, device=dev)
        out['attn_dump'] = torch.zeros(B, T, 8, n, dtype=torch.float32, device=dev)
        out['o_dump'] = torch.zeros(B, T, D, dtype=torch.float32, device=dev)
        out['lse_dump'] = torch.zeros(B, T, dtype=torch.float32, device=dev)
    if actions is not None:
        assert actions.dtype == torch.int16 and actions.shape == (B, T), (actions.dtype, actions.shape, (B, T))
    a = _lib.DecodeArgs(
        keys=ptr(keys, torch.float32), psc=ptr(psc, torch.float32), qbase=ptr(qbase, torch.float32),
        coord=ptr(task.coord, torch.uint8), demand=ptr(task.demand, torch.float32),
        tw_left=ptr(task.tw_left, torch.float32), tw_right=ptr(task.tw_right, torch.float64),
        duration=ptr(task.duration, torch.float64), distance=ptr(distance, torch.float32),
        travel=ptr(travel_table(distance), torch.float32),
        wblob=ptr(wblob, torch.float32), key_mod=Bk if reps > 1 else 0, smem_mats=int(smem_mats), mode=mode, temp=float(temp),
        seed=int(seed) & (2 ** 64 - 1), id_offset=int(id_offset), actions=ptr(actions), replay_steps=int(replay_steps),
        inject_logp=ptr(inject_logp, torch.float32), inject_z=ptr(inject_z, torch.float32), t_stride=T, pi=ptr(out['pi']),
        steps=ptr(out['steps']), status=ptr(out['status']), ll=ptr(out['ll']),
        cost=ptr(out['cost']), serve=ptr(out['serve']), logp_sel=ptr(out.get('logp_sel')),
        logp_full=ptr(out.get('logp_full')), mask_dump=ptr(out.get('mask_dump')), used_dump=ptr(out.get('used_dump')),
        cft_dump=ptr(out.get('cft_dump')), q_dump=ptr(out.get('q_dump')), attn_dump=ptr(out.get('attn_dump')),
        o_dump=ptr(out.get('o_dump')), lse_dump=ptr(out.get('lse_dump')), B=B, N=N)
    check(lib().agh_decode(C.byref(a), stream_ptr(dev)))
    return out


def get_costs(pi, task: FleetTask, distance):
    """a13: returns (cost [B] f32, status [B] i32)."""
    B, T = pi.shape
    dev = pi.device
    cost = torch.empty(B, dtype=torch.float32, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    assert pi.dtype in (torch.int64, torch.int16)
    check(lib().agh_get_costs(ptr(pi), pi.element_size(), T, ptr(task.coord), ptr(task.demand), ptr(task.tw_left),
                              ptr(task.tw_right), ptr(task.duration), ptr(distance, torch.float32), B, task.N,
                              ptr(cost), ptr(status), stream_ptr(dev)))
    return cost, status


def best_of(costs, pi, serve, R: int, B: int):
    """a16: first-minimum replica per instance (utils/functions.py:216-219)."""
    dev = costs.device
    T, n = pi.size(1), serve.size(1)
    bc = torch.empty(B, dtype=torch.float32, device=dev)
    bp = torch.empty(B, T, dtype=torch.int16, device=dev)
    bs = torch.empty(B, n, dtype=torch.float32, device=dev)
    br = torch.empty(B, dtype=torch.int32, device=dev)
    check(lib().agh_best_of(ptr(costs, torch.float32), ptr(pi, torch.int16), ptr(serve, torch.float32), R, B, n - 1, T,
                            ptr(bc), ptr(bp), ptr(bs), ptr(br), stream_ptr(dev)))
    return bc, bp, bs, br


# ------------------------------------------------------------------------------------------------
# a18: REINFORCE training primitives (csrc/train.cu)
# ------------------------------------------------------------------------------------------------
def _vptr(t: torch.Tensor, dtype=torch.float32) -> int:
    """Device pointer of a CUDA tensor VIEW (strides are described by the caller's ld / stride arguments)."""
    if not t.is_cuda:
        raise _lib.AghError('agh_b200 kernels need CUDA tensors; there is no CPU path')
    if t.dtype != dtype:
        raise _lib.AghError('expected dtype %s, got %s' % (dtype, t.dtype))
    return t.data_ptr()


def gemm(A, B, C, M, N, K, lda, ldb, ldc, transA=False, transB=False, alpha=1.0, accumulate=False, bias=None, relu=False,
         mask=None, ldm=0, splits=1, batch=(1, 1), sA=(0, 0), sB=(0, 0), sC=(0, 0)):
    """C = alpha op(A) op(B) [+ bias] [relu] [masked] [+ C]; op(A) is M x K, op(B) is K x N (see include/agh_b200.h)."""
    d = _lib.GemmDesc(A=_vptr(A), lda=lda, transA=int(transA), B=_vptr(B), ldb=ldb, transB=int(transB), C=_vptr(C), ldc=ldc,
                      M=M, N=N, K=K, batch1=batch[0], batch2=batch[1], sA1=sA[0], sA2=sA[1], sB1=sB[0], sB2=sB[1],
                      sC1=sC[0], sC2=sC[1], alpha=float(alpha), accumulate=int(accumulate),
                      bias=None if bias is None else _vptr(bias), relu=int(relu),
                      mask=None if mask is None else _vptr(mask), ldm=ldm, splits=splits)
    check(lib().agh_gemm(_ct.byref(d), stream_ptr(C.device)))
    return C


def colsum(X, M, N, ld=None, out=None):
    out = torch.zeros(N, dtype=torch.float32, device=X.device) if out is None else out
    check(lib().agh_colsum(_vptr(X), N if ld is None else ld, M, N, _vptr(out), stream_ptr(X.device)))
    return out


def bn_train_fwd(x, gamma, beta, running_mean, running_var, scratch, eps=1e-5, momentum=0.1):
    M = x.numel() // D
    y = torch.empty_like(x)
    mean = torch.empty(D, dtype=torch.float32, device=x.device)
    rstd = torch.empty(D, dtype=torch.float32, device=x.device)
    check(lib().agh_bn_train_fwd(ptr(x, torch.float32), ptr(gamma, torch.float32), ptr(beta, torch.float32), M, eps, momentum,
                                 ptr(y), ptr(mean), ptr(rstd), ptr(running_mean, torch.float32), ptr(running_var, torch.float32),
                                 ptr(scratch, torch.float64), stream_ptr(x.device)))
    return y, mean, rstd


def bn_train_bwd(dy, x, mean, rstd, gamma, scratch):
    M = x.numel() // D
    dx = torch.empty_like(x)
    dgamma = torch.empty(D, dtype=torch.float32, device=x.device)
    dbeta = torch.empty(D, dtype=torch.float32, device=x.device)
    check(lib().agh_bn_train_bwd(ptr(dy, torch.float32), ptr(x, torch.float32), ptr(mean), ptr(rstd), ptr(gamma, torch.float32), M,
                                 ptr(dx), ptr(dgamma), ptr(dbeta), ptr(scratch, torch.float64), stream_ptr(x.device)))
    return dx, dgamma, dbeta


def mha_fwd(qkv, B, N):
    heads = torch.empty(B * (N + 1), D, dtype=torch.float32, device=qkv.device)
    check(lib().agh_mha_fwd(ptr(qkv, torch.float32), B, N, ptr(heads), stream_ptr(qkv.device)))
    return heads


def mha_bwd(qkv, heads, dheads, B, N):
    dqkv = torch.empty_like(qkv)
    check(lib().agh_mha_bwd(ptr(qkv, torch.float32), ptr(heads, torch.float32), ptr(dheads, torch.float32), B, N, ptr(dqkv),
                            stream_ptr(qkv.device)))
    return dqkv


def tf_dlogits(logp_full, lse, pi, steps, grad_ll, temp, N):
    B, T, n = logp_full.shape
    dU = torch.empty_like(logp_full)
    check(lib().agh_tf_dlogits(ptr(logp_full, torch.float32), ptr(lse, torch.float32), ptr(pi, torch.int16), ptr(steps, torch.int32),
                               ptr(grad_ll, torch.float32), float(temp), B, T, N, ptr(dU), stream_ptr(dU.device)))
    return dU


def tf_dscore_(attn, dattn, N):
    rows = attn.numel() // (N + 1)
    check(lib().agh_tf_dscore(ptr(attn, torch.float32), ptr(dattn, torch.float32), rows, N, stream_ptr(attn.device)))
    return dattn


def tf_dquery(dQ, pi, steps, used, cft, N, dwcap, dwtime):
    B, T, _ = dQ.shape
    dqbase = torch.empty(B, D, dtype=torch.float32, device=dQ.device)
    dpsc = torch.zeros(B, N + 1, D, dtype=torch.float32, device=dQ.device)
    check(lib().agh_tf_dquery(ptr(dQ, torch.float32), ptr(pi, torch.int16), ptr(steps, torch.int32), ptr(used, torch.float32),
                              ptr(cft, torch.float64), B, T, N, ptr(dqbase), ptr(dpsc), ptr(dwcap, torch.float32),
                              ptr(dwtime, torch.float32), stream_ptr(dQ.device)))
    return dqbase, dpsc
