#!/usr/bin/env python
"""Greedy parity with TRAINED weights: train for a few epochs on the CUDA path, then roll the validation set out on the
GPU and through the CPU oracle (= the reference's arithmetic) with the same weights, and report the exact-tour match rate.
    python tools/check_trained_parity.py --flights 20 --epochs 2"""
import argparse
import os
import sys
import types

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--flights', type=int, default=20)
    ap.add_argument('--epochs', type=int, default=2)
    ap.add_argument('--epoch-size', type=int, default=12800)
    ap.add_argument('--val', type=int, default=1000)
    args = ap.parse_args()
    from agh_b200 import AttentionModel, AGH
    from agh_b200.reinforce_baselines import RolloutBaseline
    from agh_b200.train import train_epoch, rollout
    from oracle import agh_oracle as O
    dev = torch.device('cuda')
    torch.manual_seed(1234); np.random.seed(1234)
    model = AttentionModel(128, 128, AGH).to(dev)
    opts = types.SimpleNamespace(eval_batch_size=1000, device=dev, no_progress_bar=True, max_grad_norm=1.0, log_step=10 ** 9,
                                 no_tensorboard=True, baseline='rollout', save_dir='/tmp', checkpoint_epochs=0, n_epochs=args.epochs,
                                 epoch_size=args.epoch_size, batch_size=64, graph_size=args.flights, data_distribution=None,
                                 val_size=args.val, bl_alpha=0.05, run_name='parity')
    baseline = RolloutBaseline(model, AGH, opts)
    optimizer = torch.optim.Adam([{'params': model.parameters(), 'lr': 1e-4}])
    sched = torch.optim.lr_scheduler.LambdaLR(optimizer, lambda e: 1.0)
    np.random.seed(4321)
    val = AGH.make_dataset(size=args.flights, num_samples=args.val)
    for epoch in range(args.epochs):
        train_epoch(model, optimizer, baseline, sched, epoch, val, AGH, None, opts)
    model.eval()
    gpu = rollout(model, val, opts)
    W = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    ref = torch.cat([O.rollout(W, O.Tables.load(), {k: v[s:s + 250].clone() for k, v in val.arrays.items()}) for s in range(0, args.val, 250)], 0)
    same = torch.isclose(gpu, ref, rtol=1e-5, atol=0).all(1).float().mean().item()
    print('trained %d epoch(s) at AGH-%d: GPU avg cost %.3f, oracle (reference arithmetic) avg cost %.3f, rel diff %.2e, exact-tour match rate %.4f'
          % (args.epochs, args.flights, gpu.sum(1).mean().item(), ref.sum(1).mean().item(),
             abs(gpu.sum(1).mean().item() - ref.sum(1).mean().item()) / ref.sum(1).mean().item(), same))


if __name__ == '__main__':
    main()
