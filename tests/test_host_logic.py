"""CPU-only tests of the host side: the C-ABI library loads and exports every symbol the header declares,
the weight folds are algebraically right (checked against the oracle in f64), the host mirrors of the
reference API behave, the N>1 path works over gloo.  No kernel is launched here."""
import ctypes
import hashlib
import math
import os
import pickle
import re
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import ROOT, make_instances
from oracle import agh_oracle as O


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode()); h.update(str(a.shape).encode()); h.update(a.tobytes())
    return h.hexdigest()


# ------------------------------------------------------------------ C ABI
def test_library_exports_every_header_symbol():
    from agh_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'agh_b200.h')).read()
    names = set(re.findall(r'^(?:int|size_t|const char\*|unsigned long long)\s+(agh_\w+)\s*\(', hdr, flags=re.M))
    assert len(names) >= 18
    lib = _lib.lib()
    for nm in names:
        assert hasattr(lib, nm), 'libagh_b200.so does not export %s' % nm
    assert names == set(_lib.SYMBOLS.keys()), names ^ set(_lib.SYMBOLS.keys())
    assert lib.agh_version() >= 100
    assert lib.agh_sizeof_decode_args() == ctypes.sizeof(_lib.DecodeArgs)
    assert lib.agh_sizeof_state() == ctypes.sizeof(_lib.StateArgs)


def test_blob_size_matches_library():
    from agh_b200 import _lib
    from agh_b200.nets.weights import blob_floats
    assert _lib.lib().agh_wblob_floats() == blob_floats()


def test_cpu_tensors_are_rejected_loudly(weights):
    """No CPU fallback: handing host tensors to the product path raises."""
    from agh_b200 import ops, AghError, AttentionModel, AGH
    from agh_b200.nets.weights import pack_weights
    inst = {k: v[:2] for k, v in make_instances(20).items()}
    tables = O.Tables.load()
    task = ops.FleetTask.from_input(O.fleet_task(tables, inst, 1, O.new_tw_left_levels(inst)))
    with pytest.raises(AghError, match='CUDA'):
        ops.encode(pack_weights(weights), task)
    model = AttentionModel(128, 128, AGH)
    model.eval(); model.set_decode_type('greedy')
    with pytest.raises(AghError):
        model(O.fleet_task(tables, inst, 1, O.new_tw_left_levels(inst)))


def test_bad_arguments_return_error_codes():
    from agh_b200 import _lib
    lib = _lib.lib()
    a = _lib.DecodeArgs()
    assert lib.agh_decode(ctypes.byref(a), None) == -1
    assert b'argument check failed' in lib.agh_last_error()
    assert lib.agh_get_costs(None, 8, 5, None, None, None, None, None, None, 1, 20, None, None, None) == -1


# ------------------------------------------------------------------ weight folds
def _split_blob(blob):
    D, FF = 128, 512
    it = iter([92 * D, 3 * D, D] + [D * 3 * D, D * D, D, D, D * FF, FF, FF * D, D, D, D] * 3 + [D * 4 * D, D * D, 10 * D, D, D] +
              [3 * D * D, D * D, FF * D, D * FF] * 3 + [4 * D * D])
    out, off = [], 0
    for sz in it:
        out.append(blob[off:off + sz]); off += sz
    assert off == blob.numel()
    return out


def test_weight_folds_reproduce_the_oracle(weights, tables):
    """Emulate the kernels' arithmetic (folded BN, fused QKV, W_out folded into the logit keys, P_sc table)
    in f64 torch from the packed blob; log-probs must match the oracle's to 1e-5."""
    from agh_b200.nets.weights import pack_weights
    W = {k: v.clone() for k, v in weights.items()}
    # make BatchNorm non-trivial so the fold is exercised
    g = torch.Generator().manual_seed(3)
    for k in W:
        if k.endswith('running_mean'): W[k] = torch.randn(128, generator=g) * 0.3
        if k.endswith('running_var'): W[k] = torch.rand(128, generator=g) + 0.5
        if k.endswith('normalizer.weight'): W[k] = torch.rand(128, generator=g) + 0.5
        if k.endswith('normalizer.bias'): W[k] = torch.randn(128, generator=g) * 0.1
    parts = [p.double() for p in _split_blob(pack_weights(W))]
    loc_emb, init_w, init_b = parts[0].view(92, 128), parts[1].view(3, 128), parts[2]
    for l in range(3):          # K-major copies used by the tcgen05 GEMMs are exact transposes
        assert torch.equal(parts[38 + 4 * l].view(384, 128), parts[3 + 10 * l].view(128, 384).t())
        assert torch.equal(parts[39 + 4 * l].view(128, 128), parts[4 + 10 * l].view(128, 128).t())
        assert torch.equal(parts[40 + 4 * l].view(512, 128), parts[7 + 10 * l].view(128, 512).t())
        assert torch.equal(parts[41 + 4 * l].view(128, 512), parts[9 + 10 * l].view(512, 128).t())
    assert torch.equal(parts[50].view(512, 128), parts[33].view(128, 512).t())
    inst = {k: v[:6] for k, v in make_instances(20).items()}
    task = O.fleet_task(tables, inst, 3, O.new_tw_left_levels(inst))
    B, N = task['loc'].shape
    n = N + 1
    # a2
    coord = torch.cat((torch.zeros_like(task['loc'][:, :1]), task['loc']), 1)
    x = loc_emb[coord].clone()
    feat = torch.stack((task['demand'].double(), (task['tw_left'][:, 1:] / 1440).double(), (task['tw_right'][:, 1:] / 1440).float().double()), -1)
    x[:, 1:] += feat @ init_w + init_b
    assert torch.allclose(x.float(), O.init_embed(W, task), atol=1e-6)
    # a3
    for l in range(3):
        wqkv, wo, s1, t1, w1t, b1, w2t, b2, s2, t2 = parts[3 + 10 * l: 13 + 10 * l]
        qkv = x @ wqkv.view(128, 384)
        q, k, v = [t.view(B, n, 8, 16).transpose(1, 2) for t in qkv.split(128, -1)]
        att = torch.softmax(q @ k.transpose(-1, -2) * 0.25, -1) @ v
        x1 = (x + att.transpose(1, 2).reshape(B, n, 128) @ wo.view(128, 128)) * s1 + t1
        x = (x1 + torch.relu(x1 @ w1t.view(128, 512) + b1) @ w2t.view(512, 128) + b2) * s2 + t2
    h_ref = O.encoder(W, O.init_embed(W, task))
    assert torch.allclose(x.float(), h_ref, rtol=1e-4, atol=1e-4)      # oracle is f32, emulation f64
    # a4 + one decode step per state along the oracle's greedy tour
    wdec, wfc_t, fleet_emb, w_cap, w_time = parts[33].view(128, 512), parts[34].view(128, 128), parts[35].view(10, 128), parts[36], parts[37]
    proj = h_ref.double() @ wdec
    Kp, V, Lp, Psc = proj.split(128, -1)
    Kp = Kp / math.log2(math.e)          # the blob's K' is pre-scaled for the kernel's exp2 softmax
    qbase = h_ref.double().mean(1) @ wfc_t + fleet_emb[task['fleet'][:, 0]]
    rec = {}
    with torch.no_grad():
        O.decode(W, task, h_ref, record=rec)
    for t in range(0, len(rec['log_p']), 3):
        prev, used, cft, mask = rec['prev'][t], rec['used'][t], rec['cft'][t], rec['mask'][t]
        q = qbase + Psc[torch.arange(B), prev] + w_cap * (1.0 - used.double())[:, None] + w_time * (cft / 1440).float().double()[:, None]
        s = torch.einsum('bhk,bjhk->bhj', q.view(B, 8, 16), Kp.view(B, n, 8, 16)).masked_fill(mask[:, None], -math.inf)
        o = torch.einsum('bhj,bjhk->bhk', torch.softmax(s, -1), V.view(B, n, 8, 16)).reshape(B, 128)
        u = (10 * torch.tanh(torch.einsum('bc,bjc->bj', o, Lp))).masked_fill(mask, -math.inf)
        lp = torch.log_softmax(u, -1)
        ref = rec['log_p'][t].double()
        fin = torch.isfinite(ref)
        assert torch.equal(torch.isfinite(lp), fin)
        assert ((lp[fin] - ref[fin]).abs() / ref[fin].abs().clamp(min=1.0)).max() < 5e-5    # f32 noise of the oracle itself


# ------------------------------------------------------------------ host mirrors of the reference API
def test_model_init_and_state_dict_compat(meta, weights):
    from agh_b200 import AttentionModel, AGH
    torch.manual_seed(1234)
    m = AttentionModel(128, 128, AGH)
    sd = m.state_dict()
    assert list(sd.keys()) == list(meta['state_dict_keys'].keys())
    assert sha(*[v.numpy() for v in sd.values()]) == meta['state_dict_sha256']
    assert sum(p.numel() for p in m.parameters()) == meta['n_trainable'] == 705408
    m2 = AttentionModel(128, 128, AGH)
    m2.load_state_dict(weights)                       # a reference-format state_dict loads unchanged
    assert m.is_agh and m.distance.shape == (8464,) and m.distance.dtype == torch.float32
    fi = m.fleet_info
    assert fi['order'] == [1, 2, 4, 8, 3, 5, 7, 9, 6, 10]
    assert all(type(v) is float for v in fi['next_duration'][4]) and type(fi['next_duration'][0][0]) is np.float64
    with pytest.raises(NotImplementedError):
        AttentionModel(128, 128, AGH, rnn_time=True)


@pytest.mark.parametrize('n', [20, 100, 300])
def test_dataset_generator(meta, n):
    from agh_b200 import AGH
    np.random.seed(4321)
    ds = AGH.make_dataset(size=n, num_samples=1000)
    a = ds.arrays
    assert sha(a['loc'].numpy(), a['arrival'].long().numpy(), a['departure'].long().numpy(), a['type'].numpy(),
               a['demand'].numpy()) == meta['dataset_sha256_%d' % n]
    item = ds[3]
    assert item['loc'].dtype == torch.int64 and item['demand'].shape == (10, n) and item['arrival'].dtype == torch.float32
    assert len(ds) == 1000


def test_dataset_from_pkl(tmp_path):
    from agh_b200 import AGH
    a = make_instances(20, size=12)
    data = list(zip(a['loc'].tolist(), a['arrival'].long().tolist(), a['departure'].long().tolist(), a['type'].tolist(),
                    a['demand'].double().tolist()))
    fn = tmp_path / 'agh20.pkl'
    pickle.dump(data, open(fn, 'wb'))
    ds = AGH.make_dataset(filename=str(fn), num_samples=5, offset=2)
    assert len(ds) == 5
    for k in a:
        assert torch.equal(ds.arrays[k], a[k][2:7]), k


def test_baseline_dataset_batches():
    from agh_b200 import AGH
    from agh_b200.reinforce_baselines import BaselineDataset
    from agh_b200.train import iter_batches
    np.random.seed(1)
    ds = AGH.make_dataset(size=20, num_samples=10)
    bl = torch.arange(100.).view(10, 10)
    wrapped = BaselineDataset(ds, bl)
    batches = list(iter_batches(wrapped, 4))
    assert [b['baseline'].size(0) for b in batches] == [4, 4, 2]
    assert torch.equal(batches[1]['data']['loc'], ds.arrays['loc'][4:8])
    assert torch.equal(wrapped[7]['baseline'], bl[7]) and torch.equal(wrapped[7]['data']['loc'], ds[7]['loc'])


def test_fleet_task_from_input_dtypes(tables):
    from agh_b200.ops import FleetTask
    inst = {k: v[:3] for k, v in make_instances(20).items()}
    t1 = FleetTask.from_input(O.fleet_task(tables, inst, 1, O.new_tw_left_levels(inst)))
    t10 = FleetTask.from_input(O.fleet_task(tables, inst, 10, O.new_tw_left_levels(inst)))
    assert not t1.twr_f32 and t10.twr_f32                        # fleet 10's tw_right is f32
    assert t1.coord.dtype == torch.uint8 and t1.coord[:, 0].eq(0).all() and t1.duration[:, 0].eq(0).all()
    assert t1.tw_right.dtype == torch.float64 and t1.tw_right[0, 0] == 1441


# ------------------------------------------------------------------ N > 1 over gloo
def _dist_worker(rank, ws, port, q):
    import torch.distributed as dist
    from agh_b200 import parallel
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=ws)
    n_total = 11
    lo, hi = parallel.shard_bounds(n_total, rank, ws)
    local = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 10)
    full = parallel.all_gather_rows(local, n_total)
    ok = torch.equal(full[:, 0], torch.arange(n_total, dtype=torch.float32))
    p = torch.nn.Parameter(torch.zeros(5)); p2 = torch.nn.Parameter(torch.zeros(2, 3))
    p.grad = torch.full((5,), float(rank + 1)); p2.grad = torch.full((2, 3), float(10 * (rank + 1)))
    parallel.allreduce_grads_([p, p2])
    ok = ok and torch.allclose(p.grad, torch.full((5,), 1.5)) and torch.allclose(p2.grad, torch.full((2, 3), 15.0))
    # BatchNorm running statistics ride in the same bucket (integer buffers such as num_batches_tracked are skipped)
    p.grad = torch.full((5,), float(rank + 1))
    run_mean, run_var, tracked = torch.full((4,), float(rank)), torch.full((4,), 2.0 + 2 * rank), torch.tensor(7)
    parallel.allreduce_grads_([p, p2], buffers=[run_mean, run_var, tracked])
    ok = ok and torch.allclose(run_mean, torch.full((4,), 0.5)) and torch.allclose(run_var, torch.full((4,), 3.0)) and int(tracked) == 7
    ok = ok and torch.allclose(p.grad, torch.full((5,), 1.5)) and torch.allclose(p2.grad, torch.full((2, 3), 15.0))
    q.put((rank, bool(ok), (lo, hi)))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert res[0][1] and res[1][1]
    assert res[0][2] == (0, 6) and res[1][2] == (6, 11)


def test_shard_bounds_cover_everything():
    from agh_b200.parallel import shard_bounds
    for n in (0, 1, 7, 1000, 10000):
        for ws in (1, 2, 4, 8):
            b = [shard_bounds(n, r, ws) for r in range(ws)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


def test_options_match_reference_flags():
    """Every key of the reference's args.json (CB/data/100/args.json) is produced by get_options with the same default."""
    from agh_b200.options import get_options
    ref = {"problem": "agh", "graph_size": 100, "batch_size": 64, "epoch_size": 12800, "val_size": 1000, "val_dataset": None,
           "model": "attention", "embedding_dim": 128, "hidden_dim": 128, "n_encode_layers": 3, "tanh_clipping": 10.0,
           "normalization": "batch", "optimizer": "Adam", "lr_model": 0.0001, "lr_critic": 0.0001, "lr_decay": 1.0,
           "eval_only": False, "n_epochs": 100, "seed": 1234, "max_grad_norm": 1.0, "no_cuda": False, "exp_beta": 0.8,
           "baseline": "rollout", "bl_alpha": 0.05, "bl_warmup_epochs": 0, "eval_batch_size": 1000,
           "checkpoint_encoder": False, "shrink_size": None, "data_distribution": None, "log_step": 50, "log_dir": "logs",
           "output_dir": "outputs", "epoch_start": 0, "checkpoint_epochs": 1, "load_path": None, "resume": None,
           "no_tensorboard": True, "no_progress_bar": True, "fine_tune": False, "wo_time": False, "rnn_time": False,
           "log_stdout": True}
    o = vars(get_options(['--graph_size', '100']))
    for k, v in ref.items():
        assert k in o and o[k] == v, k
    assert o['save_dir'].startswith(os.path.join('outputs', 'agh_100', 'run_'))
    with pytest.raises(AssertionError):
        get_options(['--epoch_size', '100', '--batch_size', '64'])


def test_fp16_two_term_split_keeps_22_bits():
    """The operand split of the tensor-core GEMM / attention kernels (csrc/gemm_tc.cu split_f16_row, csrc/mha_mma.cu
    split_pair), restated in numpy: x = hi + lo'/2048 with hi = rn_f16(x), lo' = rn_f16((x - hi) * 2048).  Claims
    checked: (1) the reconstruction error is <= 2^-22 |x| over the fp16 normal range and <= 2^-36 absolute below it,
    (2) x - hi is exact in fp32, (3) the dropped lo.lo term of a product is <= 2^-22 |a b|, so a split dot product
    agrees with the float64 one to fp32 level."""
    rng = np.random.default_rng(7)
    mag = np.exp(rng.uniform(np.log(1e-9), np.log(6e4), 200000)).astype(np.float32)
    x = (mag * rng.choice([-1.0, 1.0], mag.shape)).astype(np.float32)
    hi = x.astype(np.float16)
    d = (x - hi.astype(np.float32)).astype(np.float32)
    assert np.array_equal(d.astype(np.float64), x.astype(np.float64) - hi.astype(np.float64))        # (2)
    lo = (d * np.float32(2048.0)).astype(np.float16)
    rec = hi.astype(np.float64) + lo.astype(np.float64) / 2048.0
    err = np.abs(rec - x.astype(np.float64))
    normal = np.abs(x) >= 2.0 ** -14
    assert (err[normal] <= 2.0 ** -22 * np.abs(x[normal])).all()                                    # (1)
    assert (err[~normal] <= 2.0 ** -36).all()
    assert np.isfinite(lo.astype(np.float32)).all() and np.isfinite(hi.astype(np.float32)).all()
    # (3) a K = 128 dot product from the three products the kernels issue, accumulated in float64
    a = rng.standard_normal((64, 128)).astype(np.float32) * 3
    w = rng.standard_normal((128, 128)).astype(np.float32) * 0.3

    def split(v):
        h = v.astype(np.float16)
        l = ((v - h.astype(np.float32)) * np.float32(2048.0)).astype(np.float16)
        return h.astype(np.float64), l.astype(np.float64)
    ah, al = split(a)
    wh, wl = split(w)
    got = ah @ wh.T + (ah @ wl.T + al @ wh.T) / 2048.0
    ref = a.astype(np.float64) @ w.astype(np.float64).T
    scale = np.abs(a).astype(np.float64) @ np.abs(w).astype(np.float64).T
    assert (np.abs(got - ref) <= 2.0 ** -21 * scale).all()


def test_greedy_near_tie_rule_equals_the_reference_selection():
    """The selection contract of the decode kernels, restated in numpy and checked against the reference's own arithmetic
    (torch: argmax_j exp(log_softmax(z))_j, first index on ties; attention_model.py:333,377,446-447): select the largest
    logit unless a second one lies within NEAR_TIE = 2e-6 of it; then evaluate p_j = exp((z_j - max) - log(sum exp(z - max)))
    for the candidates in that window and take the first index of the maximum.  Logits: equal values, 1-3 ulp apart far
    below the rounding grid of log_p (the reference's exp(log_p) collapses them into a tie), and decisive gaps."""
    rng = np.random.RandomState(3)
    near_tie = np.float32(2e-6)

    def kernel_rule(z):
        zmax = z.max()
        cand = np.nonzero(z >= zmax - near_tie)[0]
        if len(cand) == 1:
            return int(cand[0])
        lse = np.log(np.exp((z - zmax).astype(np.float32)).sum(dtype=np.float32)).astype(np.float32)
        p = np.exp(((z[cand] - zmax).astype(np.float32) - lse).astype(np.float32)).astype(np.float32)
        return int(cand[np.argmax(p)])

    not_largest = 0
    for trial in range(3000):
        n = int(rng.randint(2, 102))
        base = np.float32(rng.uniform(2.0 ** -6, 2.0 ** -5 * 0.99))
        kind = trial % 3
        near = (np.full(n, base, np.float32).view(np.int32) + rng.randint(0, 4, n).astype(np.int32)).view(np.float32)
        far = (base + rng.uniform(-5, 5, n)).astype(np.float32)
        z = near if kind == 0 else far if kind == 1 else np.where(rng.rand(n) < 0.3, near + np.float32(6.0), far).astype(np.float32)
        mask = rng.rand(n) < 0.4
        mask[rng.randint(n)] = False
        zt = torch.from_numpy(np.where(mask, -np.inf, z).astype(np.float32))
        ref = int(torch.log_softmax(zt, -1).exp().max(0)[1])
        mine = kernel_rule(zt.numpy())
        assert mine == ref, (trial, kind, mine, ref)
        not_largest += int(zt[ref] < zt.max())
    assert not_largest > 100          # the reference often does NOT pick the largest logit: the rule matters
