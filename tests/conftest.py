import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _make(target_dir, *args):
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, target_dir), *args], check=True)


@pytest.fixture(scope="session", autouse=True)
def built_libraries():
    """Build the product library and the oracle if they are not there yet.
    (The oracle is the checker; it is never on the product path.)"""
    if not os.path.exists(os.path.join(ROOT, "motionestimation_b200", "libme_b200.so")):
        _make("motionestimation_b200", "-j8")
    if not os.path.exists(os.path.join(ROOT, "oracle", "_build", "libme_oracle.so")):
        _make("oracle")
    if os.path.isdir("/root/reference/src/cpu") and not os.path.exists(
            os.path.join(ROOT, "oracle", "_ref", "libme_ref.so")):
        _make("oracle", "ref")
    yield
