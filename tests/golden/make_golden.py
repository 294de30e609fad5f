#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Only runs in the build container (needs /root/reference); the fixtures it
writes are committed, the reference is not.  Usage:

    python tests/golden/make_golden.py [--sizes 20 50 100] [--ref /root/reference/Construction_based]

What it pins:
  * dataset generator: sha256 of generate_agh_data(1000, N) after np.random.seed(4321)
    (generate_data.py:8-18), plus the shipped agh20 pkl equality;
  * weights: sha256 of the seed-1234 random-init state_dict (torch.manual_seed(1234);
    AttentionModel(128, 128, AGH));
  * greedy rollout (train.py:31-70 `rollout`, eval mode): per-fleet costs [B,10] for the
    whole 1000-instance set, tours / ll / serve_time for the first G instances, and per-step
    (mask, log_p, used_capacity, cur_free_time, prev_a, selected) dumps for the first S
    instances, all 10 fleets;
  * encoder: init-embed input and encoder output for the first S instances of 2 fleets;
  * sampling rollout (eval mode) with recorded actions for action-replay parity;
  * one REINFORCE training batch (train.py:159-219, train mode, batch-stat BatchNorm):
    sampled actions, per-fleet cost/ll, loss and per-parameter gradient summaries;
  * env-only KATs: nearest-neighbour and CWS heuristics (agh_baseline.py:17-101) avg cost.

The reference is imported as-is; recording is done by wrapping bound methods on the
model INSTANCE (no reference file is modified or copied).
"""
import argparse, hashlib, json, math, os, pickle, sys, time, types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode()); h.update(str(a.shape).encode()); h.update(a.tobytes())
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ref', default='/root/reference/Construction_based')
    ap.add_argument('--sizes', type=int, nargs='+', default=[20, 50, 100])
    ap.add_argument('--only', nargs='*', default=None, help='subset of: greedy sampling train kat meta')
    args = ap.parse_args()
    os.chdir(args.ref)                      # reference opens problems/agh/*.pkl by relative path
    sys.path.insert(0, args.ref)
    sys.dont_write_bytecode = True
    for name in ('matplotlib', 'matplotlib.pyplot', 'tensorboard_logger'):
        m = types.ModuleType(name); m.Logger = object; sys.modules[name] = m
    import torch
    from utils import load_problem, move_to
    from nets.attention_model import AttentionModel
    import generate_data as gd
    import train as ref_train

    want = set(args.only) if args.only else {'greedy', 'sampling', 'train', 'kat', 'meta'}
    problem = load_problem('agh')

    def make_model():
        torch.manual_seed(1234)
        m = AttentionModel(128, 128, problem)
        m.eval()
        return m

    def make_data(n, size=1000):
        np.random.seed(4321)
        return gd.generate_agh_data(size, n)

    def dataset_arrays(data):
        loc = np.asarray([d[0] for d in data], np.int64)
        arr = np.asarray([d[1] for d in data], np.int64)
        dep = np.asarray([d[2] for d in data], np.int64)
        typ = np.asarray([d[3] for d in data], np.int64)
        dem = torch.tensor([d[4] for d in data], dtype=torch.float).numpy()   # make_instance dtype
        return loc, arr, dep, typ, dem

    class DS(torch.utils.data.Dataset):
        def __init__(self, data):
            from problems.agh.problem_agh import make_instance
            self.data = [make_instance(a) for a in data]
        def __len__(self): return len(self.data)
        def __getitem__(self, i): return self.data[i]

    opts = types.SimpleNamespace(eval_batch_size=1000, device=torch.device('cpu'), no_progress_bar=True,
                                 max_grad_norm=0.0, log_step=1e9, no_tensorboard=True, baseline='rollout')
    meta = {}

    # ------------------------------------------------------------------ meta
    if 'meta' in want:
        m = make_model()
        sd = m.state_dict()
        meta['state_dict_keys'] = {k: list(v.shape) for k, v in sd.items()}
        meta['state_dict_sha256'] = sha(*[v.numpy() for v in sd.values()])
        meta['n_trainable'] = int(sum(p.numel() for p in m.parameters()))
        for n in (20, 50, 100, 200, 300):
            data = make_data(n)
            meta['dataset_sha256_%d' % n] = sha(*dataset_arrays(data))
        shipped = pickle.load(open('data/agh/agh20_validation_seed4321.pkl', 'rb'))
        meta['agh20_pkl_equals_regenerated'] = bool(
            sha(*dataset_arrays(shipped)) == meta['dataset_sha256_20'])
        meta['torch'] = torch.__version__
        meta['numpy'] = np.__version__
        json.dump(meta, open(os.path.join(HERE, 'meta.json'), 'w'), indent=1)
        print('meta', meta['state_dict_sha256'][:16], meta['agh20_pkl_equals_regenerated'])

    # ------------------------------------------------------------------ recording helpers
    def attach_recorders(model, G, S, rec, step_fleets=tuple(range(10))):
        """rec: dict of lists, one entry per fleet call."""
        orig_inner, orig_glp = model._inner, model._get_log_p
        steps = {}

        def glp(fixed, state, normalize=True):
            log_p, mask = orig_glp(fixed, state, normalize)
            steps.setdefault('log_p', []).append(log_p[:S, 0].detach().numpy().copy())
            steps.setdefault('mask', []).append(mask[:S, 0].numpy().copy())
            steps.setdefault('used', []).append(state.used_capacity[:S, 0].numpy().copy())
            steps.setdefault('cft', []).append(state.cur_free_time[:S, 0].double().numpy().copy())
            steps.setdefault('prev', []).append(state.prev_a[:S, 0].numpy().copy())
            return log_p, mask

        def inner(input, embeddings):
            steps.clear()
            log_p, pi, serve = orig_inner(input, embeddings)
            rec.setdefault('pi', []).append(pi[:G].numpy().copy())
            rec.setdefault('serve', []).append(serve[:G].numpy().copy())
            rec.setdefault('T', []).append(pi.size(1))
            rec.setdefault('h', []).append(embeddings[:S].detach().numpy().copy())
            if len(rec['pi']) - 1 in step_fleets:
                for k, v in steps.items():
                    rec.setdefault('step_' + k, []).append(np.stack(v, 1))      # [S, T, ...]
            return log_p, pi, serve

        model._inner, model._get_log_p = inner, glp
        def fwd_hook(mod, inp, out):          # hooks must return None or they replace the output
            rec.setdefault('cost', []).append(out[0].detach().numpy().copy())
            rec.setdefault('ll', []).append(out[1][:G].detach().numpy().copy())

        def emb_hook(mod, inp, out):
            rec.setdefault('h0', []).append(inp[0][:S].detach().numpy().copy())

        return model.register_forward_hook(fwd_hook), model.embedder.register_forward_hook(emb_hook)

    def pad_T(lst, fill=0):
        T = max(a.shape[1] for a in lst)
        out = []
        for a in lst:
            pad = [(0, 0)] * a.ndim; pad[1] = (0, T - a.shape[1])
            out.append(np.pad(a, pad, constant_values=fill))
        return np.stack(out, 1)      # [G, 10, T, ...]

    # ------------------------------------------------------------------ greedy
    if 'greedy' in want:
        for n in args.sizes:
            B = 1000 if n <= 100 else 128
            G = {20: 1000, 50: 256, 100: 256}.get(n, 64)
            S = {20: 8, 50: 4, 100: 2}.get(n, 1)
            GS = {20: 128, 50: 64, 100: 64}.get(n, 32)            # instances whose serve_time is kept
            step_fleets = tuple(range(10)) if n == 20 else ((0, 4, 9) if n <= 100 else (0, 9))
            data = make_data(n)[:B]
            model = make_model()
            rec = {}
            attach_recorders(model, G, S, rec, step_fleets)
            t0 = time.time()
            opts.eval_batch_size = B
            cost = ref_train.rollout(model, DS(data), opts).numpy()       # [B,10]
            dt = time.time() - t0
            tot = torch.from_numpy(cost).sum(1)
            pi = pad_T(rec['pi']).astype(np.uint8 if n < 255 else np.uint16)
            out = dict(
                costs=cost, T=np.asarray(rec['T'], np.int32), pi=pi,
                ll=np.stack(rec['ll'], 1), serve=np.stack(rec['serve'], 1)[:GS],
                step_fleets=np.asarray(step_fleets, np.int32),
                h=np.stack(rec['h'], 1)[:, :2], h0=np.stack(rec['h0'], 1)[:, :2 if n <= 50 else 0],
                step_log_p=pad_T(rec['step_log_p']), step_mask=np.packbits(pad_T(rec['step_mask'], 1), axis=-1),
                step_used=pad_T(rec['step_used']), step_cft=pad_T(rec['step_cft']),
                step_prev=pad_T(rec['step_prev']).astype(np.int16),
                avg=np.float32(tot.mean().item()),
                stderr=np.float32((torch.std(tot) / math.sqrt(len(tot))).item()),
                ref_seconds=np.float32(dt), ref_threads=np.int32(torch.get_num_threads()))
            np.savez_compressed(os.path.join(HERE, 'greedy_agh%d.npz' % n), **out)
            print('greedy N=%d B=%d avg=%r stderr=%r T=%s %.1fs' % (n, B, out['avg'], out['stderr'], rec['T'], dt))

    # ------------------------------------------------------------------ sampling (eval mode, action replay)
    if 'sampling' in want:
        n, B = 20, 64
        data = make_data(n)[:B]
        model = make_model()
        rec = {}
        attach_recorders(model, B, 4, rec)
        from nets.attention_model import set_decode_type
        # train.rollout forces greedy; drive the same fleet loop through a sampling decode type by
        # wrapping set_decode_type on the instance for the duration of the call
        model.set_decode_type('sampling')
        orig_sdt = model.set_decode_type
        model.set_decode_type = lambda *a, **k: None
        torch.manual_seed(77)
        opts.eval_batch_size = B
        cost = ref_train.rollout(model, DS(data), opts).numpy()
        model.set_decode_type = orig_sdt
        np.savez_compressed(os.path.join(HERE, 'sampling_agh20.npz'),
                            costs=cost, pi=pad_T(rec['pi']).astype(np.uint8), ll=np.stack(rec['ll'], 1),
                            serve=np.stack(rec['serve'], 1), T=np.asarray(rec['T'], np.int32),
                            step_log_p=pad_T(rec['step_log_p']))
        print('sampling N=20 B=64 mean cost', cost.sum(1).mean(), 'T', rec['T'])

    # ------------------------------------------------------------------ one REINFORCE batch (train mode)
    if 'train' in want:
        n, B = 20, 16
        data = make_data(n)[:B]
        model = make_model()
        model.train()
        rec = {}
        attach_recorders(model, B, 1, rec)
        torch.manual_seed(99)
        bl_val = torch.from_numpy(np.random.RandomState(5).uniform(20000, 30000, size=(B, 10)).astype(np.float32))
        ds = DS(data)
        batch = {'data': {k: torch.stack([ds[i][k] for i in range(B)]) for k in ds[0]}, 'baseline': bl_val}
        baseline = types.SimpleNamespace(unwrap_batch=lambda b: (b['data'], b['baseline']))
        opt = torch.optim.SGD(model.parameters(), lr=0.0)
        # capture loss: log_values receives it
        captured = {}
        ref_train.log_values = lambda cost, gn, ep, bid, step, ll, loss, bl, tb, o: captured.update(
            loss=float(loss.item()), grad_norm=float(gn[0][0]))
        opts.log_step = 1
        ref_train.train_batch_agh(model, opt, baseline, 0, 0, 0, batch, None, opts)
        grads = {k: p.grad.detach().numpy() for k, p in model.named_parameters()}
        rs = np.random.RandomState(11)
        summ = {}
        for k, g in grads.items():
            proj = rs.standard_normal(g.size).astype(np.float32)
            summ[k] = [float(np.linalg.norm(g.astype(np.float64))), float(g.astype(np.float64).sum()),
                       float((g.ravel().astype(np.float64) * proj).sum())]
        bn_stats = {k: v.numpy().copy() for k, v in model.state_dict().items() if 'running' in k or 'num_batches' in k}
        np.savez_compressed(os.path.join(HERE, 'train_agh20.npz'),
                            pi=pad_T(rec['pi']).astype(np.uint8), T=np.asarray(rec['T'], np.int32),
                            cost=np.stack(rec['cost'], 1), ll=np.stack(rec['ll'], 1), serve=np.stack(rec['serve'], 1),
                            bl_val=bl_val.numpy(), loss=np.float64(captured['loss']),
                            grad_norm=np.float64(captured['grad_norm']),
                            grad_keys=np.asarray(list(summ.keys())), grad_summ=np.asarray(list(summ.values())),
                            grad_init_embed_w=grads['init_embed.weight'], grad_fleets=grads['fleets_embedding.weight'],
                            grad_step_ctx=grads['project_step_context.weight'],
                            **{'bn_' + k: v for k, v in bn_stats.items()})
        print('train loss', captured, 'T', rec['T'])

    # ------------------------------------------------------------------ env-only KATs
    if 'kat' in want:
        import agh_baseline as ab
        n, B = 20, 1000
        data = pickle.load(open('data/agh/agh20_validation_seed4321.pkl', 'rb'))
        ds = DS(data)
        bat = {k: torch.stack([ds[i][k] for i in range(B)]) for k in ds[0]}
        model = make_model()
        fi = model.fleet_info
        res = {}
        for method in ('nearest_neighbor', 'cws'):
            tw_left_all = bat['arrival'].repeat(len(fi['next_duration']) + 1, 1, 1)
            costs, serves, pis = [], [], []
            for f in fi['order']:
                p = fi['precedence'][f]
                nd = torch.tensor(fi['next_duration'][p]).repeat(B, 1)
                tw_right = bat['departure'] - torch.gather(nd, 1, bat['type'])
                tw_right = torch.cat((torch.full_like(tw_right[:, :1], 1441), tw_right), 1)
                tw_left = torch.cat((torch.zeros_like(tw_left_all[p][:, :1]), tw_left_all[p]), 1)
                dur = torch.tensor(fi['duration'][f]).repeat(B, 1)
                fb = {'loc': bat['loc'], 'demand': bat['demand'][:, f - 1, :],
                      'distance': model.distance.expand(B, len(model.distance)),
                      'duration': torch.gather(dur, 1, bat['type']), 'tw_right': tw_right, 'tw_left': tw_left,
                      'fleet': torch.full((B, 1), f - 1)}
                if method == 'nearest_neighbor':
                    c, s = ab.nearest_neighbor(fb, problem)
                else:
                    c, s = ab.cws(fb, problem, types.SimpleNamespace(graph_size=n))
                costs.append(c.numpy()); serves.append(s[:16].numpy())
                tw_left_all[p + 1] = torch.max(tw_left_all[p + 1], s[:, 1:])
            c = np.stack(costs, 1)
            tot = torch.from_numpy(c).sum(1)
            res[method + '_costs'] = c
            res[method + '_serve'] = np.stack(serves, 1)
            res[method + '_avg'] = np.float32(tot.mean().item())
            print(method, repr(res[method + '_avg']))
        np.savez_compressed(os.path.join(HERE, 'kat_agh20.npz'), **res)


if __name__ == '__main__':
    main()
