"""Writes tests/golden/ref_checkpoint_epoch-3.pt: an `epoch-*.pt` exactly as the UNMODIFIED reference's train_epoch
saves it (train.py:130-141) -- model state_dict, Adam state, RNG states and the rollout baseline with its PICKLED
MODULE and dataset (reinforce_baselines.py:228-233).  Run in the build container (needs /root/reference or the vendored
copy under baseline/_ref):   python tests/golden/make_ref_checkpoint.py

To keep the fixture small the baseline's module is the training model itself (torch.save stores shared storages once;
the reference deep-copies it, which only doubles the bytes) and the baseline dataset holds 4 AGH-20 instances."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_arm as RA          # noqa: E402


def main():
    R = RA.load()
    model = RA.make_model(1234)
    with RA._cwd(RA.REF):
        np.random.seed(4321)
        dataset = R.AGH.make_dataset(size=20, num_samples=4)
    optimizer = torch.optim.Adam([{'params': model.parameters(), 'lr': 1e-4}])
    torch.manual_seed(99)
    ckpt = {'model': model.state_dict(), 'optimizer': optimizer.state_dict(), 'rng_state': torch.get_rng_state(),
            'cuda_rng_state': [], 'baseline': {'model': model, 'dataset': dataset, 'epoch': 3}}
    out = os.path.join(ROOT, 'tests', 'golden', 'ref_checkpoint_epoch-3.pt')
    torch.save(ckpt, out)
    print(out, os.path.getsize(out), 'bytes; baseline model class:', type(model).__module__, type(model).__qualname__)


if __name__ == '__main__':
    main()
