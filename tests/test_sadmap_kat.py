"""Known-answer test of the SAD-map pre-pass (csrc/h264_sadmap.h, k_sadmap) on its own: every tabulated number must be
the SAD of the input macroblock's 8x8 quadrant against the prediction the REFERENCE's own interpolation
(h264e_qpel_interpolate_luma H:2079, exported by oracle/ref_harness.c) produces at that position of the reference
picture -- full-sample offsets and all sixteen quarter-sample phases.  The maps are what the motion search looks up
instead of touching pixels, so this pins the kernel (and the claim that every probe of the reference's search is a
standard quarter-sample position) independently of whole-encoder parity.  CPU: the host emulation's definition of the
record; GPU (-m gpu): the sm_100a kernel."""
import ctypes as C

import numpy as np
import pytest

import cases

SM_R, SM_QR = 7, 2
SM_N, SM_QN = 2 * SM_R + 1, 2 * SM_QR + 1
SM_INT_OFF = 4
SM_Q_OFF = SM_INT_OFF + 2 * SM_N * SM_N
INVALID = 0xFFFFFFFF


def _s16(v):
    v &= 0xFFFF
    return v - 0x10000 if v & 0x8000 else v


def _check(binding, lib, ref, w, h, every):
    lib.lib.H264E_b200_ctx.restype = C.c_void_p
    frames = cases.make("panning", w, h, 3)
    enc = binding.Encoder(lib, w, h, 60)
    rp = enc.run_param(qp=28)
    enc.encode(frames[0].copy(), rp)
    enc.encode(frames[1].copy(), rp)
    refpic = enc.recon().copy()                      # reconstruction of frame 1 = reference picture of frame 2
    enc.encode(frames[2].copy(), rp)
    ctx = C.c_void_p(lib.lib.H264E_b200_ctx(C.c_void_p(enc.persist)))
    nmbx, nmby = (w + 15) // 16, (h + 15) // 16
    w16, h16 = nmbx * 16, nmby * 16
    words = np.zeros(nmbx * nmby * 1024, np.uint32)
    smw = lib.lib.h264b200_debug_get_sadmap(ctx, words.ctypes.data_as(C.c_void_p), words.size)
    assert smw > SM_Q_OFF + 2 * SM_QN * SM_QN
    recs = words[:nmbx * nmby * smw].reshape(nmbx * nmby, smw)
    enc.close()
    # padded reference luma (16 guard samples, edge replication = h264e_copy_borders H:2232), with slack for the 6-tap filters
    G = 24
    ry = np.pad(refpic[:w16 * h16].reshape(h16, w16), 16, mode="edge")
    ry = np.ascontiguousarray(np.pad(ry, G - 16, mode="edge"))
    stride = ry.shape[1]
    # input picture with the reference's edge replication for cropped sizes (pix_copy_cropped_mb H:3536)
    iy = np.pad(frames[2][:w * h].reshape(h, w), ((0, h16 - h), (0, w16 - w)), mode="edge").astype(np.int32)
    rl = ref.lib()
    rl.ref_qpel_luma.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    rl.ref_qpel_luma.restype = None
    pred = np.zeros((16, 16), np.uint8)

    def quads(mbx, mby, qx, qy):
        """quadrant SADs of MB (mbx, mby) against the prediction at macroblock-relative quarter-sample vector (qx, qy)"""
        ax, ay = mbx * 64 + qx, mby * 64 + qy
        fx, fy = ax >> 2, ay >> 2
        src = ry.ctypes.data + (fy + G) * stride + fx + G
        rl.ref_qpel_luma(C.c_void_p(src), stride, pred.ctypes.data_as(C.c_void_p), 16, 16, ax & 3, ay & 3)
        d = np.abs(iy[mby * 16:mby * 16 + 16, mbx * 16:mbx * 16 + 16] - pred.astype(np.int32))
        return [int(d[:8, :8].sum()), int(d[:8, 8:].sum()), int(d[8:, :8].sum()), int(d[8:, 8:].sum())]

    checked = 0
    for n in range(0, nmbx * nmby, every):
        mbx, mby = n % nmbx, n // nmbx
        rec = recs[n]
        assert rec[2] == 1, "record of macroblock %d not marked valid" % n
        cx, cy = _s16(int(rec[0])), _s16(int(rec[0]) >> 16)
        qcx, qcy = _s16(int(rec[1])), _s16(int(rec[1]) >> 16)
        centre_ok = False
        for k in range(SM_N * SM_N):
            lo, hi = int(rec[SM_INT_OFF + 2 * k]), int(rec[SM_INT_OFF + 2 * k + 1])
            if lo == INVALID:
                continue
            dx, dy = k % SM_N - SM_R, k // SM_N - SM_R
            want = quads(mbx, mby, 4 * (cx + dx), 4 * (cy + dy))
            assert [lo & 0xFFFF, lo >> 16, hi & 0xFFFF, hi >> 16] == want, ("integer map", n, dx, dy)
            centre_ok |= dx == 0 and dy == 0
            checked += 1
        assert centre_ok, "centre of the integer map of macroblock %d not tabulated" % n
        for k in range(SM_QN * SM_QN):
            lo, hi = int(rec[SM_Q_OFF + 2 * k]), int(rec[SM_Q_OFF + 2 * k + 1])
            if lo == INVALID:
                continue
            qx, qy = k % SM_QN - SM_QR, k // SM_QN - SM_QR
            want = quads(mbx, mby, 4 * qcx + qx, 4 * qcy + qy)
            assert [lo & 0xFFFF, lo >> 16, hi & 0xFFFF, hi >> 16] == want, ("quarter map", n, qx, qy)
            checked += 1
    assert checked > 500
    return checked


def test_sadmap_definition_against_reference_interpolation(binding, emu_lib, ref):
    """the record as the host emulation defines it (sadmap_build_mb)"""
    _check(binding, emu_lib, ref, 176, 144, every=9)
    _check(binding, emu_lib, ref, 200, 120, every=13)          # cropped: bottom / right macroblocks replicate the last row / column


@pytest.mark.gpu
def test_k_sadmap_against_reference_interpolation(binding, cuda_lib, ref):
    """the sm_100a kernel, incl. a cropped size and picture-border macroblocks"""
    _check(binding, cuda_lib, ref, 352, 288, every=11)
    _check(binding, cuda_lib, ref, 366, 250, every=17)
