#!/usr/bin/env python
"""Convert the AGH problem constants shipped with the reference into one npz.

Reads (in THIS container only) the reference fixtures
  Construction_based/problems/agh/fleet_info.pkl   (attention_model.py:78-84)
  Construction_based/problems/agh/distance.pkl     (attention_model.py:85-88)
  Construction_based/problems/agh/arrival_prob.npy (problem_agh.py:105)
and writes agh_b200/problems/agh/agh_tables.npz, the only problem-data file the
package reads at run time.  Dtypes are the ones the reference ends up with:
distance -> f32 (torch.tensor(list of python floats)), duration/next_duration
-> f64, except precedence level 4 whose next_duration is a list of python
floats and therefore becomes an f32 tensor in train.py:44-46.
"""
import pickle, sys, os
import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference/Construction_based'
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'agh_b200', 'problems', 'agh', 'agh_tables.npz')
fi = pickle.load(open(os.path.join(ref, 'problems/agh/fleet_info.pkl'), 'rb'))
dist = pickle.load(open(os.path.join(ref, 'problems/agh/distance.pkl'), 'rb'))
prob = np.load(os.path.join(ref, 'problems/agh/arrival_prob.npy'))
keys = list(dist.keys())
assert keys == [(i, j) for i in range(92) for j in range(92)], 'row-major (i,j) order expected'
duration = np.zeros((11, 3), np.float64)            # row f = fleet id 1..10
for f, v in fi['duration'].items():
    duration[f] = [float(x) for x in v]
next_duration = np.zeros((5, 3), np.float64)        # row p = precedence level 0..4
nd_is_f32 = np.zeros(5, np.uint8)
for p, v in fi['next_duration'].items():
    next_duration[p] = [float(x) for x in v]
    nd_is_f32[p] = 0 if all(isinstance(x, np.floating) for x in v) else 1
precedence = np.zeros(11, np.int64)
for f, p in fi['precedence'].items():
    precedence[f] = p
np.savez(out,
         order=np.asarray(fi['order'], np.int64), precedence=precedence,
         duration=duration, next_duration=next_duration, next_duration_is_f32=nd_is_f32,
         distance=np.asarray(list(dist.values()), np.float64).astype(np.float32),
         distance_f64=np.asarray(list(dist.values()), np.float64),
         arrival_prob=prob)
print('wrote', out, os.path.getsize(out))
