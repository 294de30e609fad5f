import sys, os; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from agh_b200 import AttentionModel, AGH, ops
from agh_b200.problems.agh.problem_agh import generate_arrays
from oracle import agh_oracle as O
torch.manual_seed(1234)
model = AttentionModel(128, 128, AGH).cuda().eval()
blob, dist = model.weight_blob(), model.distance_table()
tb = O.Tables.load()
np.random.seed(4321)
for N in (20, 50, 100, 200):
    inst = generate_arrays(512, N)
    task = ops.FleetTask.from_input({k: v.cuda() for k, v in O.fleet_task(tb, inst, 2, O.new_tw_left_levels(inst)).items() if k != 'distance'})
    h = ops.encode(blob, task); keys, psc, qb = ops.precompute(blob, h, task)
    out = ops.decode(blob, dist, task, keys, psc, qb, want_dumps=True)
    md = out['mask_dump'].cpu().numpy().astype(np.uint32)      # [B,T,words]
    steps = out['steps'].cpu().numpy()
    n = N + 1
    bits = np.unpackbits(md.view(np.uint8).reshape(md.shape[0], md.shape[1], -1), axis=2, bitorder='little')[:, :, :n]
    feas = (1 - bits).sum(2)
    valid = np.arange(md.shape[1])[None, :] < steps[:, None]
    f = feas[valid]
    print('N=%d: mean feasible %.1f, median %d, p90 %d, mean steps %.1f' % (N, f.mean(), np.median(f), np.percentile(f, 90), steps.mean()))
