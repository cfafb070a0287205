import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "a1-qp-mpc-controller_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """the shared libraries must exist; build them if this is a fresh checkout"""
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "a1-qp-mpc-controller_b200", "liba1mpc.so")) or \
            not os.path.exists(os.path.join(ROOT, "oracle", "liba1mpc_oracle.so")):
        subprocess.check_call(["make", "-C", ROOT, "-j8", "all"])
    return True


@pytest.fixture(scope="session")
def gpu_engine(built):
    import a1mpc
    eng = a1mpc.Engine(a1mpc.default_config(), device=0)
    yield eng
    eng.close()
