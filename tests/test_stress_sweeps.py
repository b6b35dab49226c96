"""Short runs of the randomised parity sweeps (tools/stress_*.py; DESIGN.md 2, profiles/r02_stress_parity.txt): random sizes,
contents, GOPs, quantisers / rate control, per-frame run parameters, heterogeneous batches, input layouts, the extension
entry points -- every result compared byte for byte with the compiled reference.  CPU: the host emulation of the device
code behind the product's host layer; GPU (-m gpu): the product library."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "_emu", "libh264lab_emu.so")
CUDA = os.path.join(ROOT, "h264-lab_b200", "libh264lab_b200.so")
TOOLS = ["stress_parity.py", "stress_frames.py", "stress_batch.py", "stress_strides.py", "stress_api.py", "stress_rc.py"]


def _run(tool, lib, seconds, seed, extra=()):
    env = dict(os.environ, H264B200_LIB=lib)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), str(seconds), str(seed), *extra],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, (tool, out.stdout[-3000:], out.stderr[-2000:])
    assert " 0 mismatches" in out.stdout, out.stdout[-500:]


@pytest.mark.parametrize("tool", TOOLS)
def test_sweep_emulation(tool, emu_lib, ref):
    _run(tool, EMU, 6, 1000 + TOOLS.index(tool))


@pytest.mark.gpu
@pytest.mark.parametrize("tool", TOOLS)
def test_sweep_gpu(tool, cuda_lib, ref):
    _run(tool, CUDA, 8, 2000 + TOOLS.index(tool))


@pytest.mark.gpu
def test_sweep_gpu_threads(cuda_lib, ref):
    """several host threads, each with its own random sessions on its own lane"""
    env = dict(os.environ, H264B200_LIB=CUDA)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress_threads.py"), "12", "4", "7"],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, (out.stdout[-3000:], out.stderr[-2000:])
