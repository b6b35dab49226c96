"""CPU suite: the product's C-ABI library loads, exports every symbol that include/*.h
declares, and its host-only entry points behave like the reference's (sizes, status
codes).  No compute calls: the container has no GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    syms = set()
    for hdr in ("h264-lab.h", "h264b200_shim.h"):
        txt = open(os.path.join(ROOT, "include", hdr)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        syms |= set(re.findall(r"\b((?:H264E|h264b200)_[A-Za-z0-9_]+)\s*\(", txt))
    return sorted(syms)


def test_exports(cuda_lib):
    syms = declared_symbols()
    assert "H264E_encode" in syms and "h264b200_encode_frames" in syms
    for s in syms:
        assert hasattr(cuda_lib.lib, s), "missing export %s" % s


SIZES = [(352, 288), (1280, 720), (1920, 1080), (3840, 2160), (16, 16), (200, 120)]


@pytest.mark.parametrize("wh", SIZES)
@pytest.mark.parametrize("const_input", [0, 1])
def test_sizeof_matches_reference(wh, const_input, binding, cuda_lib, ref):
    w, h = wh
    cp = binding.CreateParam(width=w, height=h, gop=10, const_input_flag=const_input, num_layers=1)
    a, b = C.c_int(), C.c_int()
    ra, rb = C.c_int(), C.c_int()
    rcp = ref.CreateParam(width=w, height=h, gop=10, const_input_flag=const_input, num_layers=1)
    e1 = cuda_lib.lib.H264E_sizeof(C.byref(cp), C.byref(a), C.byref(b))
    e2 = ref.lib().ref_sizeof(C.byref(rcp), C.byref(ra), C.byref(rb))
    assert e1 == e2
    if not e1:
        assert (a.value, b.value) == (ra.value, rb.value)


def test_status_codes(binding, cuda_lib, ref):
    bad = [dict(width=0, height=288), dict(width=353, height=288), dict(width=352, height=288, gop=-1),
           dict(width=352, height=288, const_input_flag=2), dict(width=352, height=288, max_long_term_reference_frames=9)]
    for kw in bad:
        a, b = C.c_int(), C.c_int()
        cp = binding.CreateParam(**kw)
        rcp = ref.CreateParam(**kw)
        assert cuda_lib.lib.H264E_sizeof(C.byref(cp), C.byref(a), C.byref(b)) == \
            ref.lib().ref_sizeof(C.byref(rcp), C.byref(a), C.byref(b)) != 0
    assert cuda_lib.lib.H264E_sizeof(None, None, None) == 1


def test_no_cpu_fallback(binding, cuda_lib):
    """Without a CUDA device H264E_init must fail (status 100), never encode on the CPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="100"):
        binding.Encoder(cuda_lib, 352, 288, 10)
