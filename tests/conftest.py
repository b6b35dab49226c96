import importlib.util
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_binding():
    spec = importlib.util.spec_from_file_location("h264lab_binding", os.path.join(ROOT, "h264-lab_b200", "binding.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def binding():
    return load_binding()


@pytest.fixture(scope="session")
def emu_lib(binding):
    """TEST-ONLY host emulation of the device code + the product's host C layer."""
    path = os.path.join(ROOT, "tests", "_emu", "libh264lab_emu.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emu")])
    return binding.Library(path)


@pytest.fixture(scope="session")
def cuda_lib(binding):
    """The product: host C + sm_100a kernels. No fallback: missing library == failure."""
    return binding.Library(os.path.join(ROOT, "h264-lab_b200", "libh264lab_b200.so"))


@pytest.fixture(scope="session")
def ref():
    import refenc
    if not refenc.have_ref():
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    if not refenc.have_ref():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return refenc
