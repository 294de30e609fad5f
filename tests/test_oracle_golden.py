"""Pin the CPU oracle (oracle/agh_oracle.py) to fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import load_golden, make_instances
from oracle import agh_oracle as O


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode()); h.update(str(a.shape).encode()); h.update(a.tobytes())
    return h.hexdigest()


def test_random_init_matches_reference(meta, weights):
    assert list(weights.keys()) == list(meta['state_dict_keys'].keys())
    assert {k: list(v.shape) for k, v in weights.items()} == meta['state_dict_keys']
    assert sha(*[v.numpy() for v in weights.values()]) == meta['state_dict_sha256']
    assert sum(v.numel() for k, v in weights.items() if 'running' not in k and 'num_batches' not in k) == meta['n_trainable']


@pytest.mark.parametrize('n', [20, 50, 100, 200, 300])
def test_generator_matches_reference(meta, tables, n):
    np.random.seed(4321)
    inst = O.generate_instances(1000, n, tables.arrival_prob)
    got = sha(inst['loc'].numpy(), inst['arrival'].long().numpy(), inst['departure'].long().numpy(),
              inst['type'].numpy(), inst['demand'].numpy())
    assert got == meta['dataset_sha256_%d' % n]
    assert meta['agh20_pkl_equals_regenerated']


def _collect(W, tables, inst):
    col = []
    costs = O.rollout(W, tables, inst, collect=col)
    return costs, col


def test_greedy_agh20_bit_exact(weights, tables):
    """Config C1: AGH-20, batch 1000, greedy: avg 253139.109375, every tour bit-exact."""
    g = load_golden('greedy_agh20.npz')
    inst = make_instances(20)
    costs, col = _collect(weights, tables, inst)
    assert np.array_equal(costs.numpy(), g['costs'])
    assert costs.sum(1).mean().item() == pytest.approx(253139.109375, abs=0)
    T = g['pi'].shape[2]
    pi = np.stack([np.pad(c['pi'].numpy(), ((0, 0), (0, T - c['pi'].shape[1]))) for c in col], 1)
    assert np.array_equal(pi, g['pi'])
    assert np.array_equal(np.stack([c['ll'].numpy() for c in col], 1), g['ll'])
    gs = g['serve'].shape[0]
    assert np.array_equal(np.stack([c['serve'].numpy()[:gs] for c in col], 1), g['serve'])
    for k in range(2):
        assert np.array_equal(col[k]['h'][:8].numpy(), g['h'][:, k])
        assert np.array_equal(O.init_embed(weights, col[k]['task'])[:8].numpy(), g['h0'][:, k])


def test_step_dumps_agh20(weights, tables):
    """Per-step mask / log_p / used_capacity / cur_free_time of the reference, all 10 fleets."""
    g = load_golden('greedy_agh20.npz')
    S = g['step_log_p'].shape[0]
    inst = make_instances(20)
    lv = O.new_tw_left_levels(inst)
    with torch.no_grad():
        for k, f in enumerate(tables.order):
            task = O.fleet_task(tables, inst, f, lv)
            rec = {}
            out = O.forward(weights, task, record=rec)
            T = int(g['T'][k])
            lp = torch.stack(rec['log_p'], 1)[:S].numpy()
            assert np.array_equal(lp, g['step_log_p'][:, k, :T])
            mask = np.packbits(torch.stack(rec['mask'], 1)[:S].numpy(), axis=-1)
            assert np.array_equal(mask, g['step_mask'][:, k, :T])
            assert np.array_equal(torch.stack(rec['used'], 1)[:S].numpy(), g['step_used'][:, k, :T])
            assert np.array_equal(torch.stack(rec['cft'], 1)[:S].numpy(), g['step_cft'][:, k, :T])
            assert np.array_equal(torch.stack(rec['prev'], 1)[:S].numpy(), g['step_prev'][:, k, :T])
            O.chain_tw_left(tables, lv, f, out['serve'])


@pytest.mark.parametrize('n,B', [(50, 1000)])
def test_greedy_agh50_bit_exact(weights, tables, n, B):
    g = load_golden('greedy_agh%d.npz' % n)
    inst = {k: v[:B] for k, v in make_instances(n).items()}
    costs, col = _collect(weights, tables, inst)
    assert np.array_equal(costs.numpy(), g['costs'][:B])
    G, T = g['pi'].shape[0], g['pi'].shape[2]
    pi = np.stack([np.pad(c['pi'].numpy()[:G], ((0, 0), (0, T - c['pi'].shape[1]))) for c in col], 1)
    assert np.array_equal(pi, g['pi'])
    assert costs.sum(1).mean().item() == pytest.approx(621320.0625, abs=0)


def test_sampling_replay_agh20(weights, tables):
    """Replaying the reference's sampled actions reproduces its cost / ll / serve_time bit-for-bit."""
    g = load_golden('sampling_agh20.npz')
    inst = {k: v[:64] for k, v in make_instances(20).items()}
    acts = [torch.from_numpy(g['pi'][:, k, :int(g['T'][k])].astype(np.int64)) for k in range(10)]
    col = []
    costs = O.rollout(weights, tables, inst, mode='replay', actions=acts, collect=col)
    assert np.array_equal(costs.numpy(), g['costs'])
    assert np.array_equal(np.stack([c['ll'].numpy() for c in col], 1), g['ll'])
    assert np.array_equal(np.stack([c['serve'].numpy() for c in col], 1), g['serve'])


def test_reinforce_batch_agh20(weights, tables):
    """train.py:159-219 with replayed actions: loss and gradients of the reference (train-mode BatchNorm)."""
    g = load_golden('train_agh20.npz')
    B = g['pi'].shape[0]
    inst = {k: v[:B] for k, v in make_instances(20).items()}
    W = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v.clone())
         for k, v in weights.items()}
    acts = [torch.from_numpy(g['pi'][:, k, :int(g['T'][k])].astype(np.int64)) for k in range(10)]
    loss, outs = O.reinforce_loss(W, tables, inst, torch.from_numpy(g['bl_val']), acts)
    assert loss.item() == pytest.approx(float(g['loss']), rel=1e-6)
    assert np.allclose(np.stack([o['cost'].detach().numpy() for o in outs], 1), g['cost'], rtol=0, atol=0)
    assert np.allclose(np.stack([o['ll'].detach().numpy() for o in outs], 1), g['ll'], rtol=1e-6)
    loss.backward()
    keys = [str(k) for k in g['grad_keys']]
    rs = np.random.RandomState(11)
    for k, ref in zip(keys, g['grad_summ']):
        gr = W[k].grad.numpy()
        proj = rs.standard_normal(gr.size).astype(np.float32)
        assert np.linalg.norm(gr.astype(np.float64)) == pytest.approx(ref[0], rel=2e-4, abs=1e-3), k
        assert (gr.ravel().astype(np.float64) * proj).sum() == pytest.approx(ref[2], rel=2e-3, abs=ref[0] * 2e-3 + 1e-3), k
    assert np.allclose(W['project_step_context.weight'].grad.numpy(), g['grad_step_ctx'], rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize('method,kat', [('nearest_neighbor', 147099.953125), ('cws', 106354.4296875)])
def test_env_kat_heuristics(tables, method, kat):
    """Weight-free KATs: the reference's nearest-neighbour (agh_baseline.py:80-101) and Clarke-Wright savings
    (agh_baseline.py:17-77) heuristics on the shipped agh20 set: avg cost 147099.953125 / 106354.4296875, per-fleet
    costs and serve times equal to the unmodified reference's."""
    g = load_golden('kat_agh20.npz')
    inst = make_instances(20)
    lv = O.new_tw_left_levels(inst)
    costs, serves = [], []
    for f in tables.order:
        task = O.fleet_task(tables, inst, f, lv)
        c, serve, _ = getattr(O, method)(task)
        costs.append(c.numpy())
        serves.append(serve[:16].numpy())
        O.chain_tw_left(tables, lv, f, serve)
    costs = np.stack(costs, 1)
    assert np.array_equal(costs, g[method + '_costs'])
    assert np.array_equal(np.stack(serves, 1), g[method + '_serve'])
    assert torch.from_numpy(costs).sum(1).mean().item() == pytest.approx(kat, abs=0)


def test_get_costs_rejects_bad_tours(weights, tables):
    inst = {k: v[:4] for k, v in make_instances(20).items()}
    task = O.fleet_task(tables, inst, 1, O.new_tw_left_levels(inst))
    _, _, pi = O.nearest_neighbor(task)
    O.get_costs(task, pi)
    bad = pi.clone(); bad[0, 0] = bad[0, 1]
    with pytest.raises(AssertionError):
        O.get_costs(task, bad)
    # all nodes in one route: capacity must trip
    one = torch.arange(1, 21).repeat(4, 1)
    with pytest.raises(AssertionError):
        O.get_costs(task, one)


def test_c_env_matches_golden(tables):
    """oracle/agh_env.c (scalar C) against the reference's nearest-neighbour KAT and the torch oracle's masks."""
    import ctypes as C
    import os
    import subprocess
    from conftest import ROOT
    subprocess.run(['make', '-C', os.path.join(ROOT, 'oracle'), '-s'], check=True)
    lib = C.CDLL(os.path.join(ROOT, 'oracle', '_build', 'libagh_env.so'))
    g = load_golden('kat_agh20.npz')
    inst = make_instances(20)
    B, N = inst['loc'].shape
    lv = O.new_tw_left_levels(inst)
    P = lambda t: C.c_void_p(t.data_ptr())
    costs = []
    for k, f in enumerate(tables.order):
        task = O.fleet_task(tables, inst, f, lv)
        loc, dem, twl = task['loc'].contiguous(), task['demand'].contiguous(), task['tw_left'].contiguous()
        twr, dur, dist = task['tw_right'].double().contiguous(), task['duration'].contiguous(), tables.distance.contiguous()
        pi = torch.zeros(B, 2 * N, dtype=torch.int64)
        steps = torch.zeros(B, dtype=torch.int32)
        serve = torch.zeros(B, N + 1)
        cost = torch.zeros(B)
        masks = torch.zeros(B, 2 * N, N + 1, dtype=torch.uint8) if k in (0, 9) else None
        lib.agh_c_nearest_neighbor(B, N, P(loc), P(dem), P(twl), P(twr), P(dur), P(dist), P(pi), P(steps), P(serve), P(cost),
                                   P(masks) if masks is not None else None)
        ref_cost, ref_serve, ref_pi = O.nearest_neighbor(task)
        T = ref_pi.size(1)
        assert torch.equal(pi[:, :T], ref_pi) and int(steps.max()) == T
        assert torch.equal(serve, ref_serve)
        assert np.allclose(cost.numpy(), g['nearest_neighbor_costs'][:, k], rtol=1e-6, atol=0)
        if masks is not None:       # masks of the torch oracle along the same tours
            env = O.Env(task)
            for t in range(T):
                alive = steps > t
                assert torch.equal(masks[alive, t].bool(), env.mask()[alive])
                env.update(ref_pi[:, t])
        costs.append(cost)
        O.chain_tw_left(tables, lv, f, serve)
    avg = torch.stack(costs, 1).sum(1).mean().item()
    assert abs(avg - 147099.953125) / 147099.953125 < 1e-6
