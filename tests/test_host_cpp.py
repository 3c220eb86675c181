"""The C++ host layer (scalable-video-codec_b200/host): reference-compatible
motion.hpp signatures and the Encoder functor, checked by tests/cpp/test_host.cpp
against the C oracle."""
import os
import subprocess

import pytest

from conftest import ROOT

PKG = os.path.join(ROOT, "scalable-video-codec_b200")
BIN = os.path.join(PKG, "build", "test_host")


def _build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True,
                   stdout=subprocess.DEVNULL)
    subprocess.run(["make", "-C", PKG, "build/test_host"], check=True, stdout=subprocess.DEVNULL)


def test_host_layer_builds_and_validates_without_gpu():
    _build()
    r = subprocess.run([BIN, "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_host_layer_encoder_matches_oracle(gpu):
    _build()
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "PASS frames=6" in r.stdout, r.stdout + r.stderr
