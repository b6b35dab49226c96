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


def test_emu_in_place_transparent_frames(binding, emu_lib, ref):
    """Host logic of the VBV-overflow transparent frames (H:6497-6508) incl. in-place reconstruction and the batch entry."""
    import numpy as np
    w, h, n = 352, 288, 6
    frames = cases.make("noise", w, h, n)
    rbs, rsizes, rrec, _ = ref.encode_sequence(frames, w, h, n, kbps=2500, empty_frames=1)
    assert list(rsizes[1:4]) == [10, 10, 10]
    enc = binding.Encoder(emu_lib, w, h, n, const_input=0, vbv_overflow_empty_frame_flag=1)
    rp = enc.run_param(kbps=2500)
    out = b""
    for i in range(n):
        f = frames[i].copy()
        out += enc.encode(f, rp)
        assert np.array_equal(f, rrec[i]), i
    enc.close()
    assert out == rbs
    encs = [binding.Encoder(emu_lib, w, h, n, vbv_overflow_empty_frame_flag=1) for _ in range(2)]
    rps = [e.run_param(kbps=2500) for e in encs]
    outs = [b"", b""]
    for i in range(n):
        res = binding.encode_batch(emu_lib, encs, [frames[i].copy(), frames[i].copy()], rps)
        outs = [o + r for o, r in zip(outs, res)]
    assert outs[0] == rbs and outs[1] == rbs
    for e in encs:
        e.close()


def _droppable_sequence(binding, lib, ref):
    """KEY, P, DROPPABLE, P, DROPPABLE, P in in-place mode (const_input_flag = 0): the caller's planes receive the
    reconstruction of EVERY frame, also of the droppable ones that never become the reference (H:6598-6623), and
    H264E_get_recon returns the reconstruction of the last encoded frame, droppable or not."""
    w, h, n = 352, 288, 6
    types = [0, 0, 1, 0, 1, 0]                 # H264E_FRAME_TYPE_DROPPABLE = 1
    frames = cases.make("panning", w, h, n)
    rs = ref.RefSession(w, h, 0, const_input=0)
    enc = binding.Encoder(lib, w, h, 0, const_input=0)
    for i in range(n):
        fr, fo = frames[i].copy(), frames[i].copy()
        rbs = rs.encode(fr, qp=28, frame_type=types[i])
        rp = enc.run_param(qp=28, frame_type=types[i])
        bs = enc.encode(fo, rp)
        assert bs == rbs, "frame %d" % i
        assert np.array_equal(fo, fr), "in-place reconstruction of frame %d" % i
        assert np.array_equal(enc.recon(), fr), "H264E_get_recon after frame %d" % i
    enc.close()


def test_emu_droppable_frames(binding, emu_lib, ref):
    _droppable_sequence(binding, emu_lib, ref)


def test_filler_nal_larger_than_the_scratch_buffer_is_an_error(binding, emu_lib):
    """vbv_underflow_stuffing_flag on a tiny picture at a high bit rate asks for a filler-data NAL that does not fit the
    caller's scratch buffer.  The reference writes past the buffer (heap corruption, found by tools/stress_parity.py);
    this layer returns H264E_STATUS_OUTPUT_OVERFLOW (103) and touches nothing beyond the buffer."""
    frames = cases.make("noise", 32, 32, 3)
    with pytest.raises(RuntimeError, match="error 103"):
        binding.encode_sequence(emu_lib, frames, 32, 32, 3, kbps=3000, empty_frames=1, stuffing=1)
