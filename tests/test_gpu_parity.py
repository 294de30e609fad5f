"""GPU parity tests: the CUDA path (through the C ABI) against the golden fixtures of the unmodified
reference and against the CPU oracle.  Run on the B200 box: pytest -m gpu.

Tolerances (BASELINE.json north_star): integer mask / state transitions and tours given identical
logits bit-exact; log-probs 1e-4 relative; tour costs 1e-5 relative.
"""
import types

import numpy as np
import pytest
import torch

from conftest import load_golden, make_instances
from oracle import agh_oracle as O

pytestmark = pytest.mark.gpu

LOGP_RTOL = 1e-4
COST_RTOL = 1e-5


def to_task(task_cpu, dev='cuda'):
    from agh_b200.ops import FleetTask
    return FleetTask.from_input({k: v.to(dev) for k, v in task_cpu.items() if k != 'distance'})


def golden_tasks(tables, inst, g, G):
    """CPU fleet tasks of the first G instances with the reference's own serve-time chain."""
    sub = {k: v[:G] for k, v in inst.items()}
    lv = O.new_tw_left_levels(sub)
    tasks = []
    for k, f in enumerate(tables.order):
        tasks.append(O.fleet_task(tables, sub, f, lv))
        O.chain_tw_left(tables, lv, f, torch.from_numpy(g['serve'][:G, k]))
    return tasks


def cuda_forward(model, task, **kw):
    from agh_b200 import ops
    blob = model.weight_blob()
    h = ops.encode(blob, task)
    keys, psc, qbase = ops.precompute(blob, h, task)
    out = ops.decode(blob, model.distance_table(), task, keys, psc, qbase, **kw)
    out['h'] = h
    return out


def unpack_mask(words, n):
    bits = torch.arange(32, device=words.device, dtype=torch.int32)
    return ((words.unsqueeze(-1) >> bits) & 1).bool().reshape(*words.shape[:-1], -1)[..., :n]


def assert_logp_close(mine, ref, what='', n=0):
    """|log_p - ref| <= 1e-4 * max(|ref|, 1) for every entry.  N = 300 (outside the headline configs): two fp32
    evaluations of 301-node softmaxes differ by up to 1.2e-4 in a handful of entries -- measured (tools/diag_logp.py):
    kernel vs float64 1.9e-4 abs, the fp32 CPU reference itself vs float64 0.6e-4 abs -- so there the bound holds for
    all but 1e-5 of the entries and no entry exceeds 2e-4."""
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(mine), fin), 'mask pattern of log_p differs ' + what
    err = np.abs(mine[fin] - ref[fin]) / np.maximum(np.abs(ref[fin]), 1.0)
    if n >= 300:
        assert (err > LOGP_RTOL).mean() <= 1e-5 and err.max() <= 2 * LOGP_RTOL, '%s: log_p rel err max %.3e, above 1e-4: %d' % (
            what, err.max(), (err > LOGP_RTOL).sum())
    else:
        assert err.max() <= LOGP_RTOL, '%s: max log_p rel err %.3e' % (what, err.max())


def test_weights_init_and_keys(cuda_model, meta, weights):
    sd = cuda_model.state_dict()
    assert list(sd.keys()) == list(meta['state_dict_keys'].keys())
    for k, v in sd.items():
        assert torch.equal(v.cpu(), weights[k]), k


@pytest.mark.parametrize('n', [20, 50])
def test_encoder_matches_reference(cuda_model, tables, n):
    """a2+a3 against the reference's own embeddings (golden h0 / h of fleets 1 and 2)."""
    from agh_b200 import ops
    g = load_golden('greedy_agh%d.npz' % n)
    S = g['h'].shape[0]
    tasks = golden_tasks(tables, make_instances(n), g, S)
    for k in range(2):
        t = to_task(tasks[k])
        h0 = ops.init_embed(cuda_model.weight_blob(), t).cpu().numpy()
        assert np.allclose(h0, g['h0'][:, k], rtol=1e-6, atol=1e-6)
        h = ops.encode(cuda_model.weight_blob(), t).cpu().numpy()
        err = np.abs(h - g['h'][:, k]).max()
        assert err < 5e-5, 'encoder max abs err %.3e' % err


@pytest.mark.parametrize('n', [20, 50, 100, 200])
def test_replay_reference_actions(cuda_model, tables, n):
    """Teacher-force the reference's greedy tours: masks, used capacity, cur_free_time, serve_time
    bit-exact; log-probs within 1e-4; cost within 1e-5."""
    from agh_b200 import _lib
    g = load_golden('greedy_agh%d.npz' % n)
    S = g['step_log_p'].shape[0]
    G = min(g['serve'].shape[0], 32)
    inst = make_instances(n)            # golden sets are the first rows of the 1000-instance seed-4321 draw
    tasks = golden_tasks(tables, inst, g, G)
    Tmax = 2 * n - 1
    for si, k in enumerate(g['step_fleets']):
        T = int(g['T'][k])
        acts = torch.zeros(G, Tmax, dtype=torch.int16)
        acts[:, :T] = torch.from_numpy(g['pi'][:G, k, :T].astype(np.int16))
        out = cuda_forward(cuda_model, to_task(tasks[k]), mode=_lib.AGH_REPLAY, actions=acts.cuda(), replay_steps=T,
                           want_logp_full=True, want_dumps=True, want_logp_sel=True)
        torch.cuda.synchronize()
        # environment: bit-exact
        mask = unpack_mask(out['mask_dump'][:S, :T], n + 1).cpu().numpy()
        ref_mask = np.unpackbits(g['step_mask'][:, si, :T], axis=-1)[..., :n + 1].astype(bool)
        assert np.array_equal(mask, ref_mask), 'feasibility mask differs (fleet idx %d)' % k
        assert np.array_equal(out['used_dump'][:S, :T].cpu().numpy(), g['step_used'][:, si, :T])
        assert np.array_equal(out['cft_dump'][:S, :T].cpu().numpy(), g['step_cft'][:, si, :T])
        assert np.array_equal(out['serve'].cpu().numpy(), g['serve'][:G, k]), 'serve_time differs'
        # numerics
        assert_logp_close(out['logp_full'][:S, :T].cpu().numpy(), g['step_log_p'][:, si, :T], 'N=%d fleet idx %d' % (n, k))
        ll = out['ll'].cpu().numpy()
        assert np.allclose(ll, g['ll'][:G, k], rtol=LOGP_RTOL, atol=1e-4)
        assert np.allclose(out['cost'].cpu().numpy(), g['costs'][:G, k], rtol=COST_RTOL, atol=0)


@pytest.mark.parametrize('n', [20, 50, 100])
def test_greedy_bit_exact_given_identical_logits(cuda_model, tables, n):
    """Inject the reference's log-probs: the selected tours must equal the reference's exactly."""
    from agh_b200 import _lib
    g = load_golden('greedy_agh%d.npz' % n)
    S = g['step_log_p'].shape[0]
    tasks = golden_tasks(tables, make_instances(n), g, S)
    Tmax = 2 * n - 1
    for si, k in enumerate(g['step_fleets']):
        T = int(g['T'][k])
        inj = torch.full((S, Tmax, n + 1), -np.inf)
        inj[:, :T] = torch.from_numpy(g['step_log_p'][:, si, :T])
        out = cuda_forward(cuda_model, to_task(tasks[k]), mode=_lib.AGH_GREEDY, inject_logp=inj.cuda())
        pi = out['pi'][:, :T].cpu().numpy()
        assert np.array_equal(pi, g['pi'][:S, k, :T].astype(np.int16))
        assert np.array_equal(out['serve'].cpu().numpy(), g['serve'][:S, k])


def oracle_fleet_records(tables, weights, inst, fleets=None):
    """Greedy oracle rollout of `inst` (all 10 fleets, the oracle's own serve-time chain) with per-step dumps.
    Yields (k, task_cpu, out, rec) for the fleets in `fleets` (positions in tables.order; default all)."""
    lv = O.new_tw_left_levels(inst)
    with torch.no_grad():
        for k, f in enumerate(tables.order):
            task = O.fleet_task(tables, inst, f, lv)
            rec = {}
            out = O.forward(weights, task, record=rec)
            if fleets is None or k in fleets:
                yield k, task, out, rec
            O.chain_tw_left(tables, lv, f, out['serve'])


@pytest.mark.parametrize('n,B', [(100, 256), (200, 64), (300, 64)])
def test_replay_oracle_actions_wide(cuda_model, tables, weights, n, B):
    """Bit-exact environment at the headline size and beyond: hundreds of instances x all 10 fleets.  The pinned
    oracle (bit-equal to the reference, tests/test_oracle_golden.py) produces the per-step dumps at test time; the
    kernel replays its tours: masks, used capacity, cur_free_time and serve_time must be identical in every bit,
    log-probs within 1e-4, costs within 1e-5."""
    from agh_b200 import _lib
    inst = {k: v[:B] for k, v in make_instances(n).items()}
    Tmax = 2 * n - 1
    checked = 0
    for k, task_cpu, out, rec in oracle_fleet_records(tables, weights, inst):
        T = out['pi'].size(1)
        acts = torch.zeros(B, Tmax, dtype=torch.int16)
        acts[:, :T] = out['pi'].to(torch.int16)
        mine = cuda_forward(cuda_model, to_task(task_cpu), mode=_lib.AGH_REPLAY, actions=acts.cuda(), replay_steps=T,
                            want_logp_full=True, want_dumps=True)
        mask = unpack_mask(mine['mask_dump'][:, :T], n + 1).cpu()
        assert torch.equal(mask, torch.stack(rec['mask'], 1)), 'feasibility mask differs (N=%d fleet idx %d)' % (n, k)
        assert torch.equal(mine['used_dump'][:, :T].cpu(), torch.stack(rec['used'], 1))
        assert torch.equal(mine['cft_dump'][:, :T].cpu(), torch.stack(rec['cft'], 1).double())
        assert torch.equal(mine['serve'].cpu(), out['serve']), 'serve_time differs'
        assert_logp_close(mine['logp_full'][:, :T].cpu().numpy(), torch.stack(rec['log_p'], 1).numpy(),
                          'N=%d fleet idx %d' % (n, k), n=n)
        assert np.allclose(mine['ll'].cpu().numpy(), out['ll'].numpy(), rtol=LOGP_RTOL, atol=1e-4)
        assert np.allclose(mine['cost'].cpu().numpy(), out['cost'].numpy(), rtol=COST_RTOL, atol=0)
        checked += mask.numel()
    print('N=%d: %d mask bits, %d instance-fleets bit-exact' % (n, checked, 10 * B))


@pytest.mark.parametrize('n,B', [(20, 128), (50, 64), (100, 32)])
def test_greedy_production_branch_given_identical_logits(cuda_model, tables, weights, n, B):
    """north_star: "greedy tours bit-exact given identical logits".  The reference's clipped logits (from the pinned
    oracle) are injected into the kernel's PRODUCTION greedy branch (inject_z: same sum-exp, same selection code as a
    normal rollout, including the near-tie re-decision): tours and serve times must equal the reference's exactly."""
    from agh_b200 import _lib
    inst = {k: v[:B] for k, v in make_instances(n).items()}
    Tmax = 2 * n - 1
    for k, task_cpu, out, rec in oracle_fleet_records(tables, weights, inst, fleets=(0, 3, 9)):
        T = out['pi'].size(1)
        z = torch.zeros(B, Tmax, n + 1)
        z[:, :T] = torch.stack(rec['logits'], 1)
        mine = cuda_forward(cuda_model, to_task(task_cpu), mode=_lib.AGH_GREEDY, inject_z=z.cuda())
        assert torch.equal(mine['pi'][:, :T].cpu().long(), out['pi']), 'tour differs (N=%d fleet idx %d)' % (n, k)
        assert torch.equal(mine['serve'].cpu(), out['serve'])
        assert int(mine['status'].max()) == 0


@pytest.mark.parametrize('temp', [1.0, 2.0])
def test_greedy_near_ties_and_duplicates(cuda_model, tables, weights, temp):
    """Constructed logits: at every step the feasible nodes carry values that are equal (duplicate flights), 1-3 ulp
    apart at |z| < 1/32 (far below the rounding grid of log_p, so the reference's exp(log_p) collapses them into a tie
    and takes the FIRST index -- selecting on z itself would take the largest), or well separated.  The production
    branch must follow the reference's argmax(exp(log_softmax(z)))."""
    from agh_b200 import _lib
    n, B = 20, 256
    inst = {k: v[:B] for k, v in make_instances(n).items()}
    task_cpu = O.fleet_task(tables, inst, 2, O.new_tw_left_levels(inst))
    Tmax = 2 * n - 1
    rng = np.random.RandomState(11)
    base = rng.uniform(2.0 ** -6, 2.0 ** -5 * 0.99, size=(B, Tmax, 1)).astype(np.float32)
    ulps = rng.randint(0, 4, size=(B, Tmax, n + 1))
    kind = rng.randint(0, 3, size=(B, Tmax, 1))
    near = (base.view(np.int32) + ulps.astype(np.int32)).view(np.float32)                         # 0..3 ulp above base
    far = (base + rng.uniform(-5, 5, size=(B, Tmax, n + 1))).astype(np.float32)  # decisive gaps
    top = rng.rand(B, Tmax, n + 1) < 0.3                                         # a near-tied cluster on top of a far field
    mixed = np.where(top, near + np.float32(6.0), far)
    z = torch.from_numpy(np.where(kind == 0, near, np.where(kind == 1, far, mixed)).astype(np.float32)) * temp
    with torch.no_grad():
        h = O.encoder(weights, O.init_embed(weights, task_cpu))
        ref = O.decode(weights, task_cpu, h, mode='greedy', temp=temp, inject_logits=z, record=(rec := {}))
    T = ref['pi'].size(1)
    mine = cuda_forward(cuda_model, to_task(task_cpu), mode=_lib.AGH_GREEDY, inject_z=z.cuda(), temp=temp)
    assert torch.equal(mine['pi'][:, :T].cpu().long(), ref['pi'])
    assert torch.equal(mine['serve'].cpu(), ref['serve'])
    # the test has teeth: in many steps the reference's choice is NOT the largest logit (it is the first index of a
    # collapsed tie), so a kernel selecting on z itself would produce different tours
    zt = z[:, :T].masked_fill(torch.stack(rec['mask'], 1), -np.inf)
    not_largest = (zt.gather(2, ref['pi'][:, :, None]).squeeze(-1) < zt.max(2)[0]).sum().item()
    assert not_largest > 100, not_largest


def test_decode_step_budget_and_status(cuda_model, tables):
    """An instance whose unvisited nodes are all infeasible (here: demand above the capacity) cannot finish: the
    reference would loop forever, the kernel stops at the step budget and flags AGH_DECODE_STUCK; N = 1 takes two
    steps (the node, then the depot) and writes inside its own pi row."""
    from agh_b200 import _lib, ops
    inst = {k: v[:4].clone() for k, v in make_instances(20).items()}
    inst['demand'][1, :, 5] = 1.5
    task_cpu = O.fleet_task(tables, inst, 1, O.new_tw_left_levels(inst))
    out = cuda_forward(cuda_model, to_task(task_cpu), mode=_lib.AGH_GREEDY)
    st = out['status'].cpu().tolist()
    assert st == [0, _lib.AGH_DECODE_STUCK, 0, 0]
    assert int(out['steps'][1]) == 39
    cuda_model.set_decode_type('greedy')
    with pytest.raises(AssertionError, match='did not terminate'):
        cuda_model(to_task(task_cpu))
    # N = 1
    one = {k: (v[:5, :, :1] if k == 'demand' else v[:5, :1]).contiguous() for k, v in make_instances(20).items()}
    t1 = O.fleet_task(tables, one, 1, O.new_tw_left_levels(one))
    o1 = cuda_forward(cuda_model, to_task(t1), mode=_lib.AGH_GREEDY)
    assert o1['pi'].shape == (5, 2) and o1['pi'].cpu().tolist() == [[1, 0]] * 5
    assert o1['steps'].cpu().tolist() == [2] * 5 and int(o1['status'].max()) == 0


def _rollout(model, arrays, bs=1000):
    from agh_b200.train import rollout
    from agh_b200 import AGHDataset
    ds = AGHDataset.__new__(AGHDataset)
    torch.utils.data.Dataset.__init__(ds)
    ds.arrays, ds.size = arrays, arrays['loc'].size(0)
    opts = types.SimpleNamespace(eval_batch_size=bs, device=torch.device('cuda'), no_progress_bar=True)
    return rollout(model, ds, opts)


# Every greedy divergence from the reference must happen at a step whose top-2 probability gap (in the reference's own
# log-probs) is below the fp32 noise bound.  Two fp32 evaluations of the network differ: measured on B200 at N = 100
# (tools/diag_logp.py, profiles/r2_logp_error.md) max |log_p - float64| is 1.5e-4 for the kernel and 0.8e-4 for the
# reference's own CPU path, kernel vs reference 1.85e-4.  A top-2 pair has p <= 0.5, so a decision can flip while the
# probability gap is below 0.5 * 1.85e-4 = 9.2e-5; typical gaps at the observed divergences are 4e-6 .. 2e-5.
GAP_TOL = 1e-4
# floors for the fraction of instances whose 10 fleet costs ALL equal the reference's, and the measured rates
# (B200, seed-4321 validation sets, random-init seed-1234 weights) are recorded in DESIGN.md section 5
MATCH_FLOOR = {20: 0.995, 50: 0.99, 100: 0.99}
MEAN_RTOL = 2e-5


def trace_divergences(model, tables, weights, sub):
    """Roll `sub` out on the GPU and through the oracle fleet by fleet; for every instance return the top-2 gap
    (reference probabilities) at the first step where the tours differ, or None when all 10 tours are equal."""
    from agh_b200.train import fleet_rollout
    pis = []
    model.set_decode_type('greedy')
    orig = model.forward

    def fwd(task, return_pi=False):
        c, l, s, p = orig(task, return_pi=True)
        pis.append(p.cpu())
        return c, l, s
    model.forward = fwd
    try:
        with torch.no_grad():
            fleet_rollout(model, sub, torch.device('cuda'), level_batch=False)
    finally:
        del model.forward
    B = sub['loc'].size(0)
    lv = O.new_tw_left_levels(sub)
    gaps = [None] * B
    alive = np.ones(B, bool)
    with torch.no_grad():
        for k, f in enumerate(tables.order):
            task = O.fleet_task(tables, sub, f, lv)
            rec = {}
            out = O.forward(weights, task, record=rec)
            mine, ref = pis[k].numpy(), out['pi'].numpy()
            T = max(mine.shape[1], ref.shape[1])
            mine = np.pad(mine, ((0, 0), (0, T - mine.shape[1])))
            ref = np.pad(ref, ((0, 0), (0, T - ref.shape[1])))
            for r in np.nonzero(alive)[0]:
                d = np.nonzero(mine[r] != ref[r])[0]
                if len(d):
                    p = rec['log_p'][d[0]][r].exp()
                    gaps[r] = (p[ref[r, d[0]]] - p[mine[r, d[0]]]).item()
                    alive[r] = False           # later fleets of this instance see different time windows
            O.chain_tw_left(tables, lv, f, out['serve'])
    return gaps


@pytest.mark.parametrize('n', [20, 50, 100])
def test_greedy_end_to_end(cuda_model, tables, weights, n):
    """Configs C1 (AGH-20, batch 1000), C2 and C3 instances against the reference's golden costs.  Greedy tours are
    chaotic in the logits, so: (i) the fraction of instances that reproduce ALL 10 reference fleet
    costs is above a floor, (ii) the mean cost agrees to 2e-5, (iii) every divergence is traced to its first differing
    step, and the reference's top-2 probability gap there is below the fp32 noise bound GAP_TOL."""
    g = load_golden('greedy_agh%d.npz' % n)
    inst = make_instances(n)
    costs = _rollout(cuda_model, inst).numpy()
    same = np.isclose(costs, g['costs'], rtol=COST_RTOL, atol=0).all(1)
    mean, ref_mean = costs.sum(1).mean(), float(g['avg'])
    print('AGH-%d: exact_tour_match_rate (all 10 fleet costs equal) %.4f; mean %.3f vs %.3f (rel %.2e)'
          % (n, same.mean(), mean, ref_mean, abs(mean - ref_mean) / ref_mean))
    assert same.mean() >= MATCH_FLOOR[n]
    assert abs(mean - ref_mean) / ref_mean < MEAN_RTOL
    bad = np.nonzero(~same)[0][:32 if n <= 50 else 12]
    if len(bad):
        gaps = trace_divergences(cuda_model, tables, weights, {k: v[bad] for k, v in inst.items()})
        print('AGH-%d: top-2 gaps at the first divergence: %s' % (n, ' '.join('%.1e' % x for x in gaps if x is not None)))
        for r, gap in zip(bad, gaps):
            assert gap is not None, 'instance %d: costs differ but all tours are equal' % r
            assert 0 <= gap < GAP_TOL, 'instance %d diverged at a decisive step (gap %.3e)' % (r, gap)


def test_shard_invariance(cuda_model):
    """Per-instance results do not depend on batch composition (basis of the multi-GPU sharding)."""
    inst = {k: v[:96] for k, v in make_instances(50).items()}
    full = _rollout(cuda_model, inst, bs=96)
    a = _rollout(cuda_model, {k: v[:32] for k, v in inst.items()}, bs=32)
    b = _rollout(cuda_model, {k: v[32:] for k, v in inst.items()}, bs=17)
    assert torch.equal(full, torch.cat((a, b), 0))


@pytest.mark.parametrize('n', [50, 100, 200])
def test_results_do_not_depend_on_timing(cuda_model, n):
    """Race detector: every warp of a decode CTA keeps its own copy of the vehicle state, so an unsynchronised
    shared-memory hand-off shows up as results that change with the timing.  A saturated GPU (two instances per SM,
    thousands of CTAs), a lightly loaded one (a few CTAs) and a repeat must give bit-identical costs."""
    base = {k: v[:1000] for k, v in make_instances(n).items()}
    inst = {k: v.repeat(*([5] + [1] * (v.dim() - 1))) for k, v in base.items()}      # 5 x 1000 instances
    B = inst['loc'].size(0)
    first = _rollout(cuda_model, inst, bs=B)
    again = _rollout(cuda_model, inst, bs=B)
    assert torch.equal(first, again)
    assert torch.equal(first[:1000], first[1000:2000])                  # identical instances, different CTAs / SMs
    light = _rollout(cuda_model, {k: v[:48] for k, v in inst.items()}, bs=16)
    assert torch.equal(first[:48], light)


@pytest.mark.parametrize('n', [50, 100, 200])
def test_sampling_and_replay_do_not_depend_on_timing(cuda_model, tables, n):
    """The same race detector for the other decode modes: Philox sampling (per-instance CTAs and the shared-key kernel)
    and action replay with the per-step dumps give bit-identical outputs on a saturated and on a lightly loaded GPU."""
    from agh_b200 import _lib
    base = {k: v[:500] for k, v in make_instances(n).items()}
    task_cpu = {k: v for k, v in O.fleet_task(tables, base, 2, O.new_tw_left_levels(base)).items() if k != 'distance'}
    big = to_task({k: v.repeat(6, *([1] * (v.dim() - 1))) for k, v in task_cpu.items()})       # 3000 CTAs: every SM loaded
    small = to_task({k: v[:24] for k, v in task_cpu.items()})                                   # 24 CTAs
    keys = ('pi', 'steps', 'll', 'cost', 'serve')
    # sampling: global Philox ids make row r of the big batch (= instance r % 500) draw the noise of id r
    heavy = cuda_forward(cuda_model, big, mode=_lib.AGH_SAMPLING, seed=5)
    again = cuda_forward(cuda_model, big, mode=_lib.AGH_SAMPLING, seed=5)
    light = cuda_forward(cuda_model, small, mode=_lib.AGH_SAMPLING, seed=5)
    for k in keys:
        assert torch.equal(heavy[k], again[k]), k
        assert torch.equal(heavy[k][:24], light[k]), k
    if n <= 100:      # shared-key kernel: 40 replicas of 75 / of 3 instances, same global ids
        sh_heavy = cuda_forward(cuda_model, to_task({k: v[:75] for k, v in task_cpu.items()}), mode=_lib.AGH_SAMPLING, seed=9, reps=40)
        sh_again = cuda_forward(cuda_model, to_task({k: v[:75] for k, v in task_cpu.items()}), mode=_lib.AGH_SAMPLING, seed=9, reps=40)
        for k in keys:
            assert torch.equal(sh_heavy[k], sh_again[k]), k
    # replay of the sampled tours with every dump enabled
    T = int(heavy['steps'].max())
    rp_heavy = cuda_forward(cuda_model, big, mode=_lib.AGH_REPLAY, actions=heavy['pi'], replay_steps=T, want_dumps=True, want_logp_sel=True)
    rp_light = cuda_forward(cuda_model, small, mode=_lib.AGH_REPLAY, actions=heavy['pi'][:24].contiguous(), replay_steps=T,
                            want_dumps=True, want_logp_sel=True)
    for k in ('ll', 'cost', 'serve', 'mask_dump', 'used_dump', 'cft_dump', 'logp_sel'):
        assert torch.equal(rp_heavy[k][:24], rp_light[k]), k
    assert torch.allclose(rp_heavy['ll'], heavy['ll'], rtol=1e-5, atol=1e-4)


def test_get_costs_kernel(cuda_model, tables):
    from agh_b200 import AGH
    inst = {k: v[:64] for k, v in make_instances(20).items()}
    task = O.fleet_task(tables, inst, 3, O.new_tw_left_levels(inst))
    ref_cost, _, pi = O.nearest_neighbor(task)
    dev_task = {k: v.cuda() for k, v in task.items()}
    cost, _ = AGH.get_costs(dev_task, pi.cuda())
    assert np.allclose(cost.cpu().numpy(), ref_cost.numpy(), rtol=1e-6, atol=0)
    bad = pi.clone(); bad[5, 0] = bad[5, 1]
    with pytest.raises(AssertionError, match='Invalid tour'):
        AGH.get_costs(dev_task, bad.cuda())
    one = torch.arange(1, 21).repeat(64, 1)
    with pytest.raises(AssertionError, match='capacity'):
        AGH.get_costs(dev_task, one.cuda())
    # time-window violation: serve flights in decreasing departure order, one per route
    order = torch.argsort(task['tw_right'][:, 1:], dim=1, descending=True) + 1
    late = torch.stack((order, torch.zeros_like(order)), 2).reshape(64, -1)
    tight = dict(dev_task); tight['tw_right'] = dev_task['tw_right'] - 1000
    with pytest.raises(AssertionError, match='time window'):
        AGH.get_costs(tight, late.cuda())


def test_state_api_nearest_neighbor(tables):
    """make_state / get_mask / update / all_finished through the C ABI, driven by the reference's
    nearest-neighbour heuristic (agh_baseline.py:80-101): tours equal the CPU oracle's exactly and the
    shipped-dataset KAT 147099.953125 is reproduced."""
    from agh_b200 import AGH
    g = load_golden('kat_agh20.npz')
    inst = make_instances(20)
    lv = O.new_tw_left_levels(inst)
    tot = []
    for f in tables.order:
        task = O.fleet_task(tables, inst, f, lv)
        dev_task = {k: v.cuda() for k, v in task.items()}
        state = AGH.make_state(dev_task)
        seq = []
        while not state.all_finished():
            mask = state.get_mask()[:, 0, :]
            prev_c = state.coords.gather(1, state.prev_a)
            d = dev_task['distance'][0][state.NODE_SIZE * prev_c + state.coords].clone()
            d[mask] = 10000
            sel = d.min(1)[1]
            state = state.update(sel)
            seq.append(sel)
        pi = torch.stack(seq, 1)
        ref_cost, ref_serve, ref_pi = O.nearest_neighbor(task)
        assert torch.equal(pi.cpu(), ref_pi)
        assert torch.equal(state.serve_time.cpu(), ref_serve)
        cost, _ = AGH.get_costs(dev_task, pi)
        assert np.allclose(cost.cpu().numpy(), ref_cost.numpy(), rtol=1e-6, atol=0)
        tot.append(cost.cpu())
        O.chain_tw_left(tables, lv, f, ref_serve)
    avg = torch.stack(tot, 1).sum(1).mean().item()
    assert abs(avg - 147099.953125) / 147099.953125 < 1e-6


def test_sampling_distribution_and_replay(cuda_model, tables, weights):
    """Philox/Gumbel sampling: first-step frequencies follow exp(log_p); a sampled rollout replayed through
    the CPU oracle gives the same cost / serve_time and ll within tolerance."""
    from agh_b200 import _lib
    inst = {k: v[:2] for k, v in make_instances(20).items()}
    task_cpu = O.fleet_task(tables, inst, 1, O.new_tw_left_levels(inst))
    task = to_task(task_cpu)
    R = 20000
    out = cuda_forward(cuda_model, task, mode=_lib.AGH_SAMPLING, seed=123, reps=R, want_logp_sel=True)
    first = out['pi'][:, 0].view(R, 2).cpu().numpy()
    with torch.no_grad():
        ref = O.forward(weights, task_cpu, record=(rec := {}))
    p0 = rec['log_p'][0].exp().numpy()
    for b in range(2):
        freq = np.bincount(first[:, b], minlength=21) / R
        assert np.abs(freq - p0[b]).max() < 4 * np.sqrt(0.25 / R) + 1e-3
    # replay 64 sampled rollouts through the oracle
    pi = out['pi'].view(R, 2, -1)[:32].reshape(64, -1)
    steps = out['steps'].view(R, 2)[:32].reshape(64)
    T = int(steps.max())
    task64 = {k: v.repeat(32, *([1] * (v.dim() - 1))) for k, v in task_cpu.items()}
    with torch.no_grad():
        h = O.encoder(weights, O.init_embed(weights, task64))
        rep = O.decode(weights, task64, h, mode='replay', actions=pi[:, :T].cpu().long())
    assert np.array_equal(out['serve'].view(R, 2, -1)[:32].reshape(64, -1).cpu().numpy(), rep['serve'].numpy())
    assert np.allclose(out['cost'].view(R, 2)[:32].reshape(64).cpu().numpy(), O.get_costs(task64, rep['pi']).numpy(), rtol=COST_RTOL)
    assert np.allclose(out['ll'].view(R, 2)[:32].reshape(64).cpu().numpy(), rep['ll'].numpy(), rtol=LOGP_RTOL, atol=1e-4)
    # different seeds give different tours, same seed the same
    again = cuda_forward(cuda_model, task, mode=_lib.AGH_SAMPLING, seed=123, reps=64)
    assert torch.equal(again['pi'], out['pi'][:128])
    other = cuda_forward(cuda_model, task, mode=_lib.AGH_SAMPLING, seed=124, reps=64)
    assert not torch.equal(other['pi'], out['pi'][:128])


def test_sample_many_best_of(cuda_model, tables):
    inst = {k: v[:8] for k, v in make_instances(20).items()}
    task = to_task(O.fleet_task(tables, inst, 2, O.new_tw_left_levels(inst)))
    cuda_model.set_decode_type('sampling')
    cuda_model.sampling_seed = 7
    try:
        pi, cost, serve = cuda_model.sample_many(task, batch_rep=64, iter_rep=1)
        from agh_b200 import _lib
        out = cuda_forward(cuda_model, task, mode=_lib.AGH_SAMPLING, seed=7, reps=64)
        allc = out['cost'].view(64, 8)
        assert torch.equal(cost, allc.min(0)[0])
        arg = allc.argmin(0)
        assert torch.equal(serve, out['serve'].view(64, 8, -1)[arg, torch.arange(8, device='cuda')])
        cuda_model.set_decode_type('greedy')
        gpi, gcost, _ = cuda_model.sample_many(task, 1, 1)
        assert gpi.dtype == torch.int64 and gcost.shape == (8,)
    finally:
        cuda_model.sampling_seed = None
        cuda_model.set_decode_type('greedy')


@pytest.mark.parametrize('n,B,R', [(20, 3, 200), (50, 2, 1280), (100, 2, 37)])
def test_shared_key_sampling_kernel(cuda_model, tables, weights, n, B, R):
    """f1: the shared-key kernel (one staged key set per CTA, one warp per replica; csrc/decode_shared.cu) draws the SAME
    Philox noise as the per-instance kernel: with the same seed nearly every replica samples the identical tour (the two
    kernels sum their dot products in different orders, so a Gumbel arg-max can flip at a near-tie); every tour is valid
    and its cost equals the validator's; replayed through the CPU oracle the serve times are exact and ll agrees to 1e-4."""
    from agh_b200 import ops, _lib
    inst = {k: v[:B] for k, v in make_instances(n).items()}
    task_cpu = O.fleet_task(tables, inst, 3, O.new_tw_left_levels(inst))
    task = to_task(task_cpu)
    shared = cuda_forward(cuda_model, task, mode=_lib.AGH_SAMPLING, seed=77, reps=R)
    single = cuda_forward(cuda_model, task, mode=_lib.AGH_SAMPLING, seed=77, reps=R, shared_keys=False)
    same = (shared['pi'] == single['pi']).all(1)
    assert same.float().mean().item() > 0.97, same.float().mean().item()
    assert torch.allclose(shared['ll'][same], single['ll'][same], rtol=LOGP_RTOL, atol=1e-4)
    assert torch.equal(shared['cost'][same], single['cost'][same]) and torch.equal(shared['serve'][same], single['serve'][same])
    assert torch.equal(shared['steps'][same], single['steps'][same]) and int(shared['status'].max()) == 0
    # validity + cost through the validator (replica-major rows: row r*B + b uses instance b)
    from agh_b200.ops import FleetTask
    rep = lambda v: v.repeat(R, *([1] * (v.dim() - 1))).contiguous()
    big = FleetTask(rep(task.coord), rep(task.demand), rep(task.tw_left), rep(task.tw_right), rep(task.duration), task.twr_f32, task.fleet,
                    None if task.fleet_ids is None else rep(task.fleet_ids))
    cost, status = ops.get_costs(shared['pi'], big, cuda_model.distance_table())
    assert int(status.max()) == 0 and torch.equal(cost, shared['cost'])
    # replay 48 sampled rollouts through the oracle
    G = min(48, R * B)
    rows = torch.arange(G)
    T = int(shared['steps'][:G].max())
    task_r = {k: torch.stack([v[int(r) % B] for r in rows]) for k, v in task_cpu.items()}
    with torch.no_grad():
        h = O.encoder(weights, O.init_embed(weights, task_r))
        rp = O.decode(weights, task_r, h, mode='replay', actions=shared['pi'][:G, :T].cpu().long())
    assert np.array_equal(shared['serve'][:G].cpu().numpy(), rp['serve'].numpy())
    assert np.allclose(shared['ll'][:G].cpu().numpy(), rp['ll'].numpy(), rtol=LOGP_RTOL, atol=2e-4)
    assert np.allclose(shared['cost'][:G].cpu().numpy(), O.get_costs(task_r, rp['pi']).numpy(), rtol=COST_RTOL)
    # different seeds, different tours; ragged replica count (R not a multiple of the warps per CTA) covered by R = 37 / 200
    other = cuda_forward(cuda_model, task, mode=_lib.AGH_SAMPLING, seed=78, reps=R)
    assert not torch.equal(other['pi'], shared['pi'])


@pytest.mark.parametrize('n,B', [(300, 4), (100, 2048)])
def test_large_sizes_properties(cuda_model, tables, weights, n, B):
    """Sizes beyond the golden fixtures: tours are valid under the reference's own asserts, steps are in
    [N+1, 2N-1], and (N=300: keys streamed through L2) a replay through the CPU oracle agrees."""
    from agh_b200 import ops, _lib
    np.random.seed(99)
    from agh_b200.problems.agh.problem_agh import generate_arrays
    inst = generate_arrays(B, n)
    task_cpu = O.fleet_task(tables, inst, 4, O.new_tw_left_levels(inst))
    task = to_task(task_cpu)
    out = cuda_forward(cuda_model, task, mode=_lib.AGH_GREEDY)
    steps = out['steps'].cpu().numpy()
    assert (steps >= n + 1).all() and (steps <= 2 * n - 1).all()
    cost, status = ops.get_costs(out['pi'], task, cuda_model.distance_table())
    assert int(status.max()) == 0
    assert torch.equal(cost, out['cost'])
    if B <= 8:
        T = int(steps.max())
        with torch.no_grad():
            h = O.encoder(weights, O.init_embed(weights, task_cpu))
            rep = O.decode(weights, task_cpu, h, mode='replay', actions=out['pi'][:, :T].cpu().long())
        assert np.array_equal(out['serve'].cpu().numpy(), rep['serve'].numpy())
        assert np.allclose(out['ll'].cpu().numpy(), rep['ll'].numpy(), rtol=LOGP_RTOL, atol=1e-4)
        assert np.abs(out['h'].cpu().numpy() - h.numpy()).max() < 1e-4


@pytest.mark.parametrize('n,B', [(20, 16), (50, 5)])
def test_training_step_matches_oracle(tables, weights, n, B):
    """a18: one REINFORCE batch in train mode (batch-stat BatchNorm) through the hand-written CUDA forward / backward
    (nets/train_fn.py, csrc/train.cu): loss and all 62 parameter gradients equal the CPU oracle's (torch autograd over the
    reference's formulation) for the same sampled actions."""
    from agh_b200 import AttentionModel, AGH
    from agh_b200.train import fleet_rollout
    torch.manual_seed(1234)
    model = AttentionModel(128, 128, AGH).cuda()
    model.train()
    model.set_decode_type('sampling')
    inst = {k: v[:B] for k, v in make_instances(n).items()}
    bl = torch.from_numpy(np.random.RandomState(5).uniform(20000, 30000, size=(B, 10)).astype(np.float32))
    pis = []
    orig = model.forward
    def fwd(task, return_pi=False):
        c, l, s, p = orig(task, return_pi=True)
        pis.append(p.cpu())
        return c, l, s
    model.forward = fwd
    costs, lls = fleet_rollout(model, inst, torch.device('cuda'))
    del model.forward
    loss = sum(((costs[:, i] - bl.cuda()[:, i]) * lls[i]).mean() for i in range(10)) / 10
    loss.backward()
    W = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v.clone())
         for k, v in weights.items()}
    ref_loss, outs = O.reinforce_loss(W, tables, inst, bl, pis)
    ref_loss.backward()
    assert np.allclose(costs.detach().cpu().numpy(), np.stack([o['cost'].detach().numpy() for o in outs], 1), rtol=COST_RTOL)
    assert np.allclose(torch.stack(lls, 1).detach().cpu().numpy(), np.stack([o['ll'].detach().numpy() for o in outs], 1),
                       rtol=LOGP_RTOL, atol=1e-3)
    assert loss.item() == pytest.approx(ref_loss.item(), rel=1e-3)
    # biases feeding a train-mode BatchNorm have mathematically zero gradient (pure rounding noise), so the
    # error is measured against the larger of the tensor's own norm and 1e-5 of the global gradient norm
    gnorm = np.sqrt(sum(float((W[k].grad ** 2).sum()) for k, _ in model.named_parameters()))
    for k, p in model.named_parameters():
        gm, gr = p.grad.cpu().numpy(), W[k].grad.numpy()
        denom = max(np.linalg.norm(gr), 1e-5 * gnorm)
        assert np.linalg.norm(gm - gr) / denom < 5e-3, 'gradient of %s differs: %.3e' % (k, np.linalg.norm(gm - gr) / denom)
    # running statistics were updated like the reference's (momentum 0.1, 6 BN layers x 10 fleets)
    assert int(model.state_dict()['embedder.layers.0.1.normalizer.num_batches_tracked']) == 10


def test_training_gradients_large_graph(tables, weights):
    """The training path on the N > 127 variant of the rollout kernel (compact value / logit phases, V through L2 or shared
    memory) and the 150-row attention backward: log-likelihood and parameter gradients of two fleet sub-problems against
    the oracle's autograd on the same sampled actions."""
    from agh_b200 import AttentionModel, AGH
    from agh_b200.problems.agh.problem_agh import generate_arrays
    torch.manual_seed(1234)
    model = AttentionModel(128, 128, AGH).cuda()
    model.train()
    model.set_decode_type('sampling')
    n, B = 200, 3
    np.random.seed(17)
    inst = generate_arrays(B, n)
    lv = O.new_tw_left_levels(inst)
    W = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v.clone()) for k, v in weights.items()}
    coef = torch.tensor([1.0, -0.5, 2.0])
    loss, ref_loss = 0.0, 0.0
    for f in (1, 2):
        task_cpu = O.fleet_task(tables, inst, f, lv)
        cost, ll, serve, pi = model(to_task(task_cpu), return_pi=True)
        out = O.forward(W, task_cpu, mode='replay', training=True, actions=pi.cpu())
        assert np.allclose(ll.detach().cpu().numpy(), out['ll'].detach().numpy(), rtol=LOGP_RTOL, atol=2e-3)
        assert np.array_equal(serve.cpu().numpy(), out['serve'].detach().numpy())
        loss = loss + (coef.cuda() * ll).sum()
        ref_loss = ref_loss + (coef * out['ll']).sum()
        O.chain_tw_left(tables, lv, f, out['serve'].detach())
    loss.backward()
    ref_loss.backward()
    gnorm = np.sqrt(sum(float((W[k].grad ** 2).sum()) for k, _ in model.named_parameters() if W[k].grad is not None))
    for k, p in model.named_parameters():
        gm, gr = p.grad.cpu().numpy(), W[k].grad.numpy()
        denom = max(np.linalg.norm(gr), 1e-5 * gnorm)
        assert np.linalg.norm(gm - gr) / denom < 5e-3, 'gradient of %s differs: %.3e' % (k, np.linalg.norm(gm - gr) / denom)


def test_training_gradients_tensor_core_rows(tables, weights):
    """The training path with more than 16,384 rows per fleet (B = 170 at N = 100), where the activation-sized GEMMs of the
    encoder's forward and backward run on the tcgen05 kernel (EPI_TRAIN epilogue: bias / ReLU / mask / accumulate) and the weight
    gradients on the split-K TN kernel: log-likelihood and parameter gradients of two fleet sub-problems against the oracle's
    autograd on the same sampled actions."""
    from agh_b200 import AttentionModel, AGH
    from agh_b200.problems.agh.problem_agh import generate_arrays
    torch.manual_seed(1234)
    model = AttentionModel(128, 128, AGH).cuda()
    model.train()
    model.set_decode_type('sampling')
    n, B = 100, 170
    np.random.seed(17)
    inst = generate_arrays(B, n)
    lv = O.new_tw_left_levels(inst)
    W = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v.clone()) for k, v in weights.items()}
    coef = torch.linspace(-1.0, 2.0, B)
    loss, ref_loss = 0.0, 0.0
    for f in (1, 2):
        task_cpu = O.fleet_task(tables, inst, f, lv)
        cost, ll, serve, pi = model(to_task(task_cpu), return_pi=True)
        out = O.forward(W, task_cpu, mode='replay', training=True, actions=pi.cpu())
        assert np.allclose(ll.detach().cpu().numpy(), out['ll'].detach().numpy(), rtol=LOGP_RTOL, atol=2e-3)
        assert np.array_equal(serve.cpu().numpy(), out['serve'].detach().numpy())
        loss = loss + (coef.cuda() * ll).sum()
        ref_loss = ref_loss + (coef * out['ll']).sum()
        O.chain_tw_left(tables, lv, f, out['serve'].detach())
    loss.backward()
    ref_loss.backward()
    gnorm = np.sqrt(sum(float((W[k].grad ** 2).sum()) for k, _ in model.named_parameters() if W[k].grad is not None))
    for k, p in model.named_parameters():
        gm, gr = p.grad.cpu().numpy(), W[k].grad.numpy()
        # the second feed-forward bias feeds a train-mode BatchNorm: its gradient is mathematically zero and both sides hold
        # rounding noise that grows with the 17,170 rows summed here -- judged against the global gradient norm
        floor = 1e-3 if k.endswith('2.module.2.bias') else 1e-5
        denom = max(np.linalg.norm(gr), floor * gnorm)
        assert np.linalg.norm(gm - gr) / denom < 5e-3, 'gradient of %s differs: %.3e' % (k, np.linalg.norm(gm - gr) / denom)



@pytest.mark.parametrize('K,N', [(128, 384), (128, 512), (512, 128), (384, 128)])
def test_training_gemm_on_tensor_cores(K, N):
    """agh_gemm with op(B) = B^T on a K-major weight and M >= 16,384: plain and bias / ReLU run on the tcgen05 kernel, the mask
    and accumulate forms on the CUDA-core kernels -- all against a float64 evaluation and the CUDA-core route on the [K][N] copy."""
    from agh_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(K + N)
    M = 20011
    A = torch.randn(M, K, generator=g, device='cuda')
    W = torch.randn(K, N, generator=g, device='cuda') / K ** 0.5             # [K][N]
    Wt = W.t().contiguous()                                                   # [N][K]
    bias = torch.randn(N, generator=g, device='cuda')
    mask = torch.randn(M, N, generator=g, device='cuda')
    C0 = torch.randn(M, N, generator=g, device='cuda')
    ref0 = A.double() @ W.double()
    for kw, ref in (({}, ref0),
                    ({'bias': bias, 'relu': True}, torch.relu(ref0 + bias.double())),
                    ({'bias': bias, 'accumulate': True}, ref0 + bias.double() + C0.double()),
                    ({'mask': mask, 'ldm': N}, torch.where(mask > 0, ref0, torch.zeros_like(ref0))),
                    ({'accumulate': True}, ref0 + C0.double())):
        c_tc = ops.gemm(A, Wt, C0.clone(), M, N, K, K, K, N, transB=True, **kw)
        c_cc = ops.gemm(A, W, C0.clone(), M, N, K, K, N, N, **kw)
        torch.cuda.synchronize()
        for name, c in (('tcgen05', c_tc), ('CUDA cores', c_cc)):
            e = (c.double() - ref).abs().max().item()
            assert e < 2e-5, '%s K=%d N=%d %s: max abs err %.3e' % (name, K, N, sorted(kw), e)


@pytest.mark.parametrize('n,B', [(20, 7), (100, 300), (50, 1000), (200, 40), (300, 12)])
def test_tensor_core_gemm_matches_cuda_core(cuda_model, tables, n, B):
    """tcgen05 split-FP16 GEMMs and mma.sync split-FP16 attention (encoder + precompute) against the fp32 CUDA-core path
    and a float64 evaluation: fp32-level agreement,
    including ragged last row tiles (M = B*(n+1) is not a multiple of 128)."""
    from agh_b200 import ops, _lib
    inst = {k: v[:B] for k, v in make_instances(n).items()}
    task = to_task(O.fleet_task(tables, inst, 5, O.new_tw_left_levels(inst)))
    blob = cuda_model.weight_blob()
    lib = _lib.lib()
    try:
        lib.agh_set_gemm_backend(0)
        h_s = ops.encode(blob, task)
        keys_s, psc_s, qb_s = ops.precompute(blob, h_s, task)
        lib.agh_set_gemm_backend(1)
        lib.agh_set_ffn_fused(2)                 # the fused feed-forward and (up to 112 nodes) the tcgen05 attention kernel,
        lib.agh_set_mha_tc(2)                    # whatever the size thresholds of the default configuration would pick
        h_t = ops.encode(blob, task)
        keys_t, psc_t, qb_t = ops.precompute(blob, h_s, task)
        torch.cuda.synchronize()
    finally:
        lib.agh_set_gemm_backend(1)
        lib.agh_set_ffn_fused(1)
        lib.agh_set_mha_tc(1)
    # 12 chained GEMMs + 3 attentions with split (hi + lo') operands; |h| reaches ~13.  Both back ends are judged against a float64
    # evaluation of the same layers: the split products carry ~2^-21 relative error each, a few times the fp32 FMA
    # chain's, which stays fp32-level after three layers.
    W64 = {k: v.double() for k, v in cuda_model.state_dict().items()}
    ref = O.encoder(W64, ops.init_embed(blob, task).double())
    for name, hh, lim_max, lim_mean in (('cuda-core', h_s, 1e-4, 5e-6), ('tensor-core', h_t, 2e-4, 1e-5)):
        e = (hh.double() - ref).abs()
        assert e.max().item() < lim_max, '%s encoder vs float64: max abs err %.3e' % (name, e.max().item())
        assert e.mean().item() < lim_mean, '%s encoder vs float64: mean abs err %.3e' % (name, e.mean().item())
    assert (h_t - h_s).abs().max().item() < 2e-4
    assert (keys_t - keys_s).abs().max().item() < 3e-5
    assert (psc_t - psc_s).abs().max().item() < 3e-5
    # the same against the two-GEMM feed-forward + mma.sync attention form, and against what the default configuration picks
    try:
        lib.agh_set_ffn_fused(0)
        lib.agh_set_mha_tc(0)
        h_u = ops.encode(blob, task)
        torch.cuda.synchronize()
    finally:
        lib.agh_set_ffn_fused(1)
        lib.agh_set_mha_tc(1)
    h_d = ops.encode(blob, task)
    for hh in (h_u, h_d):
        e = (hh.double() - ref).abs()
        assert e.max().item() < 2e-4 and e.mean().item() < 1e-5
        assert (h_t - hh).abs().max().item() < 5e-5


@pytest.mark.parametrize('M', [1, 127, 128, 129, 4200, 148 * 128 + 5, 300007])
def test_fused_feed_forward_kernel(cuda_model, M):
    """ffn_tc.cu alone through agh_encoder_ffn: y = BN2(x + W2 relu(W1 x + b1) + b2) (graph_encoder.py:159-172, eval-mode
    BatchNorm with non-trivial running statistics) against a float64 evaluation, for every layer, on row counts around
    the 128-row tile, around one tile per SM and many tiles per SM; the two-GEMM form on the same input as a second witness.
    Measured (relative to max(1, |y|)): fused max 3.2e-6 / mean 1.9e-7, two GEMMs max 1.7e-6 / mean 1.0e-7 -- the fused kernel
    keeps ONE hi.hi accumulator for the K = 512 contraction (TMEM is full), the two-GEMM form spreads it over three."""
    from agh_b200 import ops, _lib
    import torch.nn.functional as F
    lib = _lib.lib()
    sd = {k: v.clone() for k, v in cuda_model.state_dict().items()}
    g = torch.Generator(device='cuda').manual_seed(M)
    try:
        with torch.no_grad():
            for k, v in cuda_model.state_dict().items():
                if 'running_mean' in k: v.copy_(torch.randn(v.shape, generator=g, device='cuda') * 0.3)
                if 'running_var' in k: v.copy_(torch.rand(v.shape, generator=g, device='cuda') * 1.5 + 0.5)
                if k.endswith('normalizer.bias'): v.copy_(torch.randn(v.shape, generator=g, device='cuda') * 0.1)
        cuda_model.invalidate_blob()
        blob = cuda_model.weight_blob()
        W = {k: v.double() for k, v in cuda_model.state_dict().items()}
        x = torch.randn(M, 128, generator=g, device='cuda') * 2.0
        for layer in range(3):
            pre = 'embedder.layers.%d.' % layer
            ff = F.linear(F.relu(F.linear(x.double(), W[pre + '2.module.0.weight'], W[pre + '2.module.0.bias'])),
                          W[pre + '2.module.2.weight'], W[pre + '2.module.2.bias'])
            nrm = pre + '3.normalizer.'
            ref = (x.double() + ff - W[nrm + 'running_mean']) / torch.sqrt(W[nrm + 'running_var'] + 1e-5) * W[nrm + 'weight'] + W[nrm + 'bias']
            lib.agh_set_ffn_fused(2)
            y_f = ops.encoder_ffn(blob, layer, x)
            lib.agh_set_ffn_fused(0)
            y_u = ops.encoder_ffn(blob, layer, x)
            lib.agh_set_ffn_fused(1)
            torch.cuda.synchronize()
            scale = ref.abs().clamp_min(1.0)
            for name, y in (('fused', y_f), ('two GEMMs', y_u)):
                e = (y.double() - ref).abs() / scale
                print('feed-forward %s layer %d M=%d: max rel err %.3e mean %.3e' % (name, layer, M, e.max().item(), e.mean().item()))
                assert e.max().item() < 6e-6, '%s feed-forward, layer %d, M=%d: max rel err %.3e' % (name, layer, M, e.max().item())
                assert e.mean().item() < 4e-7, '%s feed-forward, layer %d, M=%d: mean rel err %.3e' % (name, layer, M, e.mean().item())
    finally:
        lib.agh_set_ffn_fused(1)
        cuda_model.load_state_dict(sd)
        cuda_model.invalidate_blob()


@pytest.mark.parametrize('N,B', [(1, 5), (20, 7), (50, 300), (100, 149), (100, 1000), (111, 33), (112, 9)])
def test_tensor_core_attention_kernel(N, B):
    """mha_tc.cu alone through agh_encoder_mha (graph_encoder.py:83-104: softmax(Q K^T / sqrt(16)) V per head, no projection)
    against a float64 evaluation and against the mma.sync kernel on the same input: every node count class of the tcgen05
    kernel (one key group, several, the full 112 keys with and without padding keys), instance counts below and above one per
    SM, and N = 112 (n = 113 nodes), which is served by the mma.sync kernel."""
    from agh_b200 import ops, _lib
    lib = _lib.lib()
    n = N + 1
    g = torch.Generator(device='cuda').manual_seed(100 * N + B)
    qkv = torch.randn(3, B * n, 128, generator=g, device='cuda')
    qkv[0] *= 1.5                                           # scores up to ~ +-25 in base 2: sharp rows
    q, k, v = (qkv[i].double().view(B, n, 8, 16).permute(0, 2, 1, 3) for i in range(3))
    ref = torch.softmax(q @ k.transpose(2, 3) / 4.0, dim=-1) @ v              # [B, 8, n, 16]
    ref = ref.permute(0, 2, 1, 3).reshape(B * n, 128)
    try:
        lib.agh_set_mha_tc(2)
        out_tc = ops.encoder_mha(qkv, B, N)
        lib.agh_set_mha_tc(0)
        out_mma = ops.encoder_mha(qkv, B, N)
        torch.cuda.synchronize()
    finally:
        lib.agh_set_mha_tc(1)
    for name, out in (('tcgen05', out_tc), ('mma.sync', out_mma)):
        e = (out.double() - ref).abs()
        print('attention %s N=%d B=%d: max err %.3e mean %.3e' % (name, N, B, e.max().item(), e.mean().item()))
        assert e.max().item() < 5e-6, '%s attention N=%d B=%d: max abs err %.3e' % (name, N, B, e.max().item())
        assert e.mean().item() < 3e-7, '%s attention N=%d B=%d: mean abs err %.3e' % (name, N, B, e.mean().item())
