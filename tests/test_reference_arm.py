"""The unmodified reference, vendored under baseline/_ref (oracle/vendor_reference.py), against the oracle and the
package's checkpoint loader.  CPU only; each check runs in its own interpreter because the reference's top-level
module names (nets, problems, utils, train) must not leak into the other tests."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.isfile(os.path.join(ROOT, 'baseline', '_ref', 'train.py'))


def _run(code):
    res = subprocess.run([sys.executable, '-c', code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    return res.stdout


@pytest.mark.skipif(not HAVE_REF, reason='baseline/_ref not vendored (python oracle/vendor_reference.py)')
def test_oracle_equals_the_vendored_reference_rollout():
    """train.rollout of the unmodified reference and oracle.rollout give bit-equal [B,10] cost tables (AGH-20 and
    AGH-50, seed-4321 instances, seed-1234 weights): the "port" is a faithful stand-in for the reference arm."""
    out = _run('''
import sys; sys.path.insert(0, "tests")
import torch
from conftest import make_instances
from oracle import reference_arm as RA, agh_oracle as O
model = RA.make_model(1234)
W, tb = O.random_init_weights(1234), O.Tables.load()
for k, v in model.state_dict().items():
    assert torch.equal(v, W[k]), k
for n, B in ((20, 48), (50, 16)):
    inst = {k: v[:B] for k, v in make_instances(n).items()}
    ref = RA.rollout(model, inst)
    mine = O.rollout(W, tb, inst)
    assert ref.shape == (B, 10) and torch.equal(ref, mine), n
print("EQUAL")
''')
    assert 'EQUAL' in out


def test_reference_checkpoint_with_pickled_baseline_module_loads():
    """A reference-written epoch-*.pt (tests/golden/make_ref_checkpoint.py: pickled nets.attention_model.AttentionModel
    and problems.agh.problem_agh.AGHDataset inside `baseline`) loads through utils.torch_load_cpu without the
    reference tree: the module paths are aliased to this package's state_dict-compatible classes."""
    out = _run('''
import sys, torch
from agh_b200.utils import torch_load_cpu
from oracle import agh_oracle as O
ck = torch_load_cpu("tests/golden/ref_checkpoint_epoch-3.pt")
assert set(ck) == {"model", "optimizer", "rng_state", "cuda_rng_state", "baseline"}
bl = ck["baseline"]
m = bl["model"]
assert type(m).__module__ == "agh_b200.nets.attention_model" and bl["epoch"] == 3
W = O.random_init_weights(1234)
sd = m.state_dict()
assert list(sd) == list(ck["model"]) == list(W)
for k in W:
    assert torch.equal(sd[k], W[k]) and torch.equal(ck["model"][k], W[k]), k
ds = bl["dataset"]
assert type(ds).__module__ == "agh_b200.problems.agh.problem_agh" and len(ds) == 4
assert ds[2]["demand"].shape == (10, 20) and ds.arrays["loc"].shape == (4, 20)
assert "nets" not in sys.modules and "problems" not in sys.modules       # the aliases are gone after the load
print("LOADED")
''')
    assert 'LOADED' in out
