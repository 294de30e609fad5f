#!/usr/bin/env python
"""The activation-sized GEMMs of a training step at batch 512 (51,712 rows): agh_gemm's tcgen05 route (op(B) = B^T on the K-major
weight; plain and bias / ReLU run on the tensor cores, mask / accumulate stay on the CUDA cores) against the CUDA-core SGEMM.
    python tools/bench_train_gemm.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from agh_b200 import ops


def timeit(fn, iters=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
M = int(sys.argv[1]) if len(sys.argv) > 1 else 51712
for K, N in ((128, 384), (128, 128), (128, 512), (512, 128), (384, 128)):
    A = torch.randn(M, K, device='cuda'); W = torch.randn(K, N, device='cuda'); Wt = W.t().contiguous()
    C = torch.zeros(M, N, device='cuda'); bias = torch.randn(N, device='cuda'); mask = torch.randn(M, N, device='cuda')
    for name, kw in (('plain', {}), ('bias+relu', {'bias': bias, 'relu': True}), ('accumulate', {'accumulate': True}), ('mask', {'mask': mask, 'ldm': N})):
        t_tc = timeit(lambda: ops.gemm(A, Wt, C, M, N, K, K, K, N, transB=True, **kw))
        t_cc = timeit(lambda: ops.gemm(A, W, C, M, N, K, K, N, N, **kw))
        print('M=%d K=%d N=%d %-10s tcgen05 %7.1f us   CUDA cores %7.1f us' % (M, K, N, name, t_tc, t_cc), flush=True)
