import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "symmetry-ode-discovery_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("WANDB_MODE", "disabled")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the C-ABI library is needed by the loader tests (CPU) and by every GPU test
    so = os.path.join(PKG, "libsindy_b200.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-j8", "-C", os.path.join(PKG, "csrc")], check=True, stdout=subprocess.DEVNULL)


def pytest_report_header(config):
    """GPU identity in the test header, so that a numerical outlier can be tied to a physical device."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=name,serial,uuid", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip()
        return [f"gpu: {line}" for line in out.splitlines()] if out else None
    except (OSError, subprocess.SubprocessError):
        return None


class Golden:
    def __init__(self, name):
        self._z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))

    def __getitem__(self, k):
        return self._z[k]

    def keys(self):
        return self._z.files


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = Golden(name)
        return cache[name]

    return get
