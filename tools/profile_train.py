#!/usr/bin/env python
"""Where one REINFORCE training batch spends its time: wall clock per phase (with device syncs), device time from CUDA
events, library kernel launches, and the top device kernels from torch.profiler.
    python tools/profile_train.py --flights 100 --batch-size 64"""
import argparse
import os
import sys
import time
import types

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--flights', type=int, default=100)
    ap.add_argument('--batch-size', type=int, default=64)
    ap.add_argument('--profile', action='store_true')
    ap.add_argument('--cprofile', action='store_true', help='host-side profile of one step (cProfile, cumulative time)')
    args = ap.parse_args()
    from agh_b200 import AttentionModel, AGH, _lib, ops
    from agh_b200.train import fleet_rollout
    dev = torch.device('cuda')
    torch.manual_seed(1234); np.random.seed(1234)
    model = AttentionModel(128, 128, AGH).to(dev)
    model.train(); model.set_decode_type('sampling')
    model.defer_checks = True          # one validity sync per batch (train_batch_agh does the same)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    ds = AGH.make_dataset(size=args.flights, num_samples=args.batch_size)
    bat = ds.arrays
    bl = torch.zeros(args.batch_size, 10, device=dev)
    lib = _lib.lib()

    def step():
        t = {}
        torch.cuda.synchronize(); t0 = time.perf_counter(); l0 = lib.agh_launch_count()
        model.weight_blob()
        torch.cuda.synchronize(); t['pack_blob'] = time.perf_counter() - t0; t1 = time.perf_counter()
        costs, lls = fleet_rollout(model, bat, dev)
        model.raise_pending()
        torch.cuda.synchronize(); t['forward_10_fleets'] = time.perf_counter() - t1; t2 = time.perf_counter()
        loss = sum(((costs[:, i] - bl[:, i]) * lls[i]).mean() for i in range(10)) / 10
        opt.zero_grad()
        loss.backward()
        torch.cuda.synchronize(); t['backward'] = time.perf_counter() - t2; t3 = time.perf_counter()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        torch.cuda.synchronize(); t['clip+adam'] = time.perf_counter() - t3
        t['total'] = time.perf_counter() - t0
        t['launches'] = lib.agh_launch_count() - l0
        return t

    for _ in range(3):
        step()
    ts = [step() for _ in range(5)]
    print({k: round(1e3 * sum(x[k] for x in ts) / len(ts), 3) if k != 'launches' else ts[0][k] for k in ts[0]}, '(ms)')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); w0 = time.perf_counter()
    for _ in range(5):
        costs, lls = fleet_rollout(model, bat, dev)
        loss = sum(((costs[:, i] - bl[:, i]) * lls[i]).mean() for i in range(10)) / 10
        opt.zero_grad(); loss.backward(); opt.step()
    e1.record(); torch.cuda.synchronize()
    print('5 steps: wall %.1f ms/step, device span %.1f ms/step' % (1e3 * (time.perf_counter() - w0) / 5, e0.elapsed_time(e1) / 5))
    if args.cprofile:
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(3):
            step()
        pr.disable()
        st = pstats.Stats(pr)
        st.sort_stats('tottime').print_stats(35)
    if args.profile:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            step()
        print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=30, max_name_column_width=60))
        print(prof.key_averages().table(sort_by='self_cpu_time_total', row_limit=20, max_name_column_width=60))


if __name__ == '__main__':
    main()
