import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
for p in (ROOT, GOLDEN):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_sessionstart(session):
    """The in-tree library is a build artefact (git-ignored): (re)build it when it is missing or older than its sources
    (incremental, a no-op otherwise; nvcc cross-compiles sm_100a without a GPU)."""
    try:
        from shapemol_b200 import build as _build
        _build.build()
    except Exception as e:     # no nvcc: the tests that need the library report it themselves
        print('shapemol_b200: library build skipped (%s)' % e)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location='cpu', weights_only=False)


def manifest_shapes():
    with open(os.path.join(GOLDEN, 'manifest.json')) as f:
        return {k: tuple(v[0]) for k, v in json.load(f).items()}


def golden_weights(fx):
    """Synthetic weights for a forward/trajectory fixture (see tests/golden/synth.py)."""
    import synth
    shapes = {k: tuple(v) for k, v in fx['shapes'].items()} if 'shapes' in fx else manifest_shapes()
    return synth.synth_state_dict(shapes, fx['seed'])


def oracle_cfg(fx):
    from oracle import shapemol_oracle as orc
    cfg = dict(orc.DEFAULT_CFG)
    cfg.update(knn=fx['k'], hidden_dim=fx.get('hidden', 128), n_heads=fx.get('heads', 16))
    return cfg


@pytest.fixture(scope='session')
def cuda_lib():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from shapemol_b200 import _lib
    return _lib.load()
