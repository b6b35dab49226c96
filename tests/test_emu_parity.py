"""CPU suite: the macroblock code that nvcc compiles for sm_100a, compiled for the host
(tests/emu, TEST ONLY) behind the product's host C layer, must reproduce the compiled
reference (oracle/_ref) byte for byte: bit stream, per-frame sizes and reconstruction.
This pins the algorithm; the GPU suite (test_gpu_parity.py) pins the CUDA execution."""
import numpy as np
import pytest

import cases


@pytest.mark.parametrize("case", cases.SMALL + cases.CIF_FOREMAN_SUBSTITUTE[:2], ids=lambda c: c[0])
def test_emu_matches_reference(case, binding, emu_lib, ref):
    name, kind, w, h, n, gop, kw = case
    frames = cases.make(kind, w, h, n)
    rbs, rsizes, rrec, _ = ref.encode_sequence(frames, w, h, gop, **kw)
    bs, sizes, rec = binding.encode_sequence(emu_lib, frames, w, h, gop, **kw)
    assert list(sizes) == list(rsizes)
    assert bs == rbs
    assert np.array_equal(rec, rrec)
