import os
import sys

import pytest

# The library reads its IQL_* environment switches (path selection for A/B tests) only under this master debug flag.
os.environ.setdefault("IQL_B200_DEBUG", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")
    config.addinivalue_line("markers", "slow: longer CPU test")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The engine library must exist for every test (CPU tests load it and check symbols)."""
    from jsrl_corl_b200 import build as _b

    _b.build()
    so = os.path.join(ROOT, "oracle", "_build", "libphilox_ref.so")
    if not os.path.exists(so):
        import subprocess

        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "oracle", "philox_ref.c")])

