#!/usr/bin/env python
"""Where does the log-prob difference between the CUDA path and the fp32 CPU reference come from?

For one fleet of a batch the reference's greedy tours are replayed through (a) the kernel with the tensor-core encoder,
(b) the kernel with the fp32 CUDA-core encoder, (c) the fp32 oracle and (d) a float64 evaluation of the same network;
prints the max / mean absolute log-prob differences of (a), (b), (c) against (d) and of (a) against (c).

    python tools/diag_logp.py --flights 300 --batch 64 --fleet-idx 4
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--flights', type=int, default=300)
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--fleet-idx', type=int, default=4)
    args = ap.parse_args()
    from agh_b200 import AttentionModel, AGH, ops, _lib
    from oracle import agh_oracle as O
    from conftest import make_instances
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    n, B = args.flights, args.batch
    torch.manual_seed(1234)
    model = AttentionModel(128, 128, AGH).cuda().eval()
    W = O.random_init_weights(1234)
    W64 = {k: (v.double() if v.is_floating_point() else v) for k, v in W.items()}
    tb = O.Tables.load()
    inst = {k: v[:B] for k, v in make_instances(n).items()}
    lv = O.new_tw_left_levels(inst)
    with torch.no_grad():
        for k, f in enumerate(tb.order):
            task = O.fleet_task(tb, inst, f, lv)
            rec = {}
            out = O.forward(W, task, record=rec)
            if k == args.fleet_idx:
                break
            O.chain_tw_left(tb, lv, f, out['serve'])
        T = out['pi'].size(1)
        h64 = O.encoder(W64, O.init_embed(W, task).double())
        rec64 = {}
        O.decode(W64, task, h64, mode='replay', actions=out['pi'], record=rec64)
    ref32 = torch.stack(rec['log_p'], 1).double()
    ref64 = torch.stack(rec64['log_p'], 1)
    fin = torch.isfinite(ref64)
    acts = torch.zeros(B, 2 * n - 1, dtype=torch.int16)
    acts[:, :T] = out['pi'].to(torch.int16)
    dtask = ops.FleetTask.from_input({k: v.cuda() for k, v in task.items() if k != 'distance'})
    blob = model.weight_blob()
    res = {}
    for name, backend in (('tensor-core encoder', 1), ('cuda-core encoder', 0)):
        _lib.lib().agh_set_gemm_backend(backend)
        h = ops.encode(blob, dtask)
        keys, psc, qbase = ops.precompute(blob, h, dtask)
        o = ops.decode(blob, model.distance_table(), dtask, keys, psc, qbase, mode=_lib.AGH_REPLAY, actions=acts.cuda(),
                       replay_steps=T, want_logp_full=True)
        res[name] = o['logp_full'][:, :T].cpu().double()
        print('%-20s |h - h64| max %.3e mean %.3e' % (name, (h.cpu().double() - h64).abs().max(), (h.cpu().double() - h64).abs().mean()))
    _lib.lib().agh_set_gemm_backend(1)
    print('oracle fp32          |h - h64| max %.3e' % (out['h'].double() - h64).abs().max())

    def stats(a, b, what):
        d = (a - b).abs()[fin]
        rel = (d / ref64.abs()[fin].clamp(min=1.0))
        i = int(d.argmax())
        print('%-44s max abs %.3e  mean abs %.3e  max rel(>=1) %.3e  (log_p at max: %.4f)' % (what, d.max(), d.mean(), rel.max(), ref64[fin][i]))
    stats(res['tensor-core encoder'], ref64, 'kernel (tensor-core encoder) vs float64')
    stats(res['cuda-core encoder'], ref64, 'kernel (cuda-core encoder) vs float64')
    stats(ref32, ref64, 'fp32 CPU reference (oracle) vs float64')
    stats(res['tensor-core encoder'], ref32, 'kernel (tensor-core encoder) vs fp32 reference')
    stats(res['cuda-core encoder'], ref32, 'kernel (cuda-core encoder) vs fp32 reference')


if __name__ == '__main__':
    main()
