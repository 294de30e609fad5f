"""agh_b200: B200-native rollout path of RoyalSkye/AGH's Construction_based attention model.

Public surface mirrors the reference:
    from agh_b200 import AttentionModel, AGH, StateAGH, set_decode_type
    from agh_b200.train import rollout, validate, train_batch_agh, train_epoch
    from agh_b200.reinforce_baselines import RolloutBaseline
    from agh_b200.eval import _eval_dataset
All compute goes through libagh_b200.so (include/agh_b200.h); there is no CPU fallback.
"""
from .nets.attention_model import AttentionModel, set_decode_type
from .problems.agh.problem_agh import AGH, AGHDataset
from .problems.agh.state_agh import StateAGH
from ._lib import build_library, AghError

__all__ = ['AttentionModel', 'set_decode_type', 'AGH', 'AGHDataset', 'StateAGH', 'build_library', 'AghError']
