import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scalable-video-codec_b200"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()  # builds liboracle.so on first use if it is missing
    return O


@pytest.fixture(scope="session")
def svc():
    import subprocess

    import svc_b200
    if not os.path.exists(svc_b200.lib_path()):
        # fresh checkout (built artefacts are git-ignored): build the product first
        subprocess.run(["make", "-C", os.path.join(ROOT, "scalable-video-codec_b200"), "-j4"], check=True,
                       stdout=subprocess.DEVNULL)
    svc_b200.lib()  # hard failure if the CUDA library cannot be loaded
    return svc_b200


@pytest.fixture(scope="session")
def gpu(svc):
    if svc.device_count() < 1:
        pytest.fail("gpu-marked test selected but no CUDA device is visible")
    return svc


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))
