"""Drive the UNMODIFIED reference (vendored copy under baseline/_ref, see oracle/vendor_reference.py) on the CPU --
TEST / BENCHMARK INFRASTRUCTURE ONLY: bench.py's `--impl reference` arm and cpu_baseline leg, and a cross-check of
the oracle in tests/.  Nothing under agh_b200/ imports this.

The reference is a script tree with relative pickle paths (attention_model.py:78,85; problem_agh.py:105) and two
imports this image lacks (matplotlib in utils/log_utils.py:3, tensorboard_logger in run.py:11): `load()` puts the
copy first on sys.path, installs EMPTY stand-in modules for those two, and changes into the copy while the model
is constructed.  The reference code itself is untouched.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')


def available() -> bool:
    return os.path.isfile(os.path.join(REF, 'train.py')) and os.path.isfile(os.path.join(REF, 'nets', 'attention_model.py'))


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


_mods = None


def load():
    """Import the vendored reference; returns a namespace with train, AttentionModel, AGH, AGHDataset."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError('baseline/_ref is missing: run `python oracle/vendor_reference.py` where /root/reference exists')
    for name in ('matplotlib', 'matplotlib.pyplot', 'tensorboard_logger'):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                if name == 'tensorboard_logger':
                    m.Logger = object
                sys.modules[name] = m
    if 'matplotlib' in sys.modules and not hasattr(sys.modules['matplotlib'], 'pyplot') and 'matplotlib.pyplot' in sys.modules:
        sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    # the reference's top-level package names (nets, problems, utils, train) must resolve to the vendored copy
    clash = [k for k in sys.modules if k.split('.')[0] in ('nets', 'problems', 'utils', 'train', 'reinforce_baselines')
             and not getattr(sys.modules[k], '__file__', REF).startswith(REF)]
    assert not clash, 'modules shadowing the reference tree are already imported: %s' % clash
    sys.path.insert(0, REF)
    try:
        with _cwd(REF):
            import train as ref_train
            from nets.attention_model import AttentionModel
            from problems.agh.problem_agh import AGH, AGHDataset
    finally:
        sys.path.remove(REF)
    _mods = types.SimpleNamespace(train=ref_train, AttentionModel=AttentionModel, AGH=AGH, AGHDataset=AGHDataset)
    return _mods


def make_model(seed: int = 1234):
    """torch.manual_seed(seed); AttentionModel(128, 128, AGH) -- the reference's random init."""
    R = load()
    with _cwd(REF):
        torch.manual_seed(seed)
        model = R.AttentionModel(128, 128, R.AGH)
    model.eval()
    return model


def make_dataset(arrays):
    """A reference AGHDataset holding the given instance arrays (same item dicts as make_instance, problem_agh.py:82-90)."""
    R = load()
    ds = R.AGHDataset.__new__(R.AGHDataset)
    torch.utils.data.Dataset.__init__(ds)
    B = arrays['loc'].size(0)
    ds.data = [{k: arrays[k][i] for k in ('loc', 'arrival', 'departure', 'type', 'demand')} for i in range(B)]
    ds.size = B
    return ds


def rollout(model, arrays, batch_size=None):
    """The reference's own train.rollout (train.py:31-70) on the CPU -> per-fleet costs [B,10]."""
    R = load()
    B = arrays['loc'].size(0)
    opts = types.SimpleNamespace(eval_batch_size=batch_size or B, device=torch.device('cpu'), no_progress_bar=True)
    return R.train.rollout(model, make_dataset(arrays), opts)
