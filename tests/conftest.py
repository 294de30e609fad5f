import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def meta():
    return json.load(open(os.path.join(GOLDEN, 'meta.json')))


@pytest.fixture(scope='session')
def tables():
    from oracle import agh_oracle as O
    return O.Tables.load()


@pytest.fixture(scope='session')
def weights():
    from oracle import agh_oracle as O
    return O.random_init_weights(1234)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def make_instances(n, size=1000, seed=4321):
    """Seed-4321 validation set of the reference (generate_data.py:8-18), via the package generator."""
    from agh_b200.problems.agh.problem_agh import generate_arrays
    np.random.seed(seed)
    return generate_arrays(size, n)


@pytest.fixture(scope='session')
def cuda_model():
    from agh_b200 import AttentionModel, AGH
    torch.manual_seed(1234)
    m = AttentionModel(128, 128, AGH).cuda()
    m.eval()
    return m
