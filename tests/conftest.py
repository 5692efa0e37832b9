import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'carla-driving-rl-agent_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with `-m gpu`)')


@pytest.fixture(scope='session')
def built_libs():
    """Build (if stale) the sm_100a library and the CPU logic-check build once per session."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('cdra_build', os.path.join(PKG, 'build.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if os.path.exists('/usr/local/cuda/bin/nvcc'):
        mod.build_cuda()
    mod.build_emu()
    return mod
