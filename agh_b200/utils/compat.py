"""Loading checkpoints written by the REFERENCE (train.py:130-141).

`epoch-*.pt` holds `baseline = {'model': <whole nn.Module, pickled>, 'dataset': <AGHDataset>, 'epoch'}`
(reinforce_baselines.py:228-233).  Pickle stores classes by module path -- `nets.attention_model.AttentionModel`,
`nets.graph_encoder.*`, `problems.agh.problem_agh.{AGH,AGHDataset}` -- which only resolve inside the reference's
script tree.  `reference_module_aliases()` maps those paths onto this package's state_dict-compatible classes for
the duration of a load, so a reference checkpoint resumes unchanged."""
from __future__ import annotations

import contextlib
import importlib
import sys

_ALIASES = {
    'nets': 'agh_b200.nets',
    'nets.attention_model': 'agh_b200.nets.attention_model',
    'nets.graph_encoder': 'agh_b200.nets.graph_encoder',
    'problems': 'agh_b200.problems',
    'problems.agh': 'agh_b200.problems.agh',
    'problems.agh.problem_agh': 'agh_b200.problems.agh.problem_agh',
    'problems.agh.state_agh': 'agh_b200.problems.agh.state_agh',
}


@contextlib.contextmanager
def reference_module_aliases():
    """Temporarily register the reference's module paths as aliases of this package's modules (paths that are
    already importable, e.g. when running inside the reference tree, are left alone)."""
    added = []
    try:
        for name, target in _ALIASES.items():
            if name not in sys.modules:
                sys.modules[name] = importlib.import_module(target)
                added.append(name)
        yield
    finally:
        for name in added:
            sys.modules.pop(name, None)
