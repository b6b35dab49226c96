"""ctypes binding of the B200 encoder's C API (include/h264-lab.h).

Python is only a convenience for the tests and bench.py; the product is the C-ABI
shared library ``h264-lab_b200/libh264lab_b200.so`` (host C + sm_100a CUDA).  There is
no CPU fallback: loading fails loudly when the library has not been built, and
``H264E_init`` fails with status 100 when no CUDA device is usable.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.environ.get("H264B200_LIB") or os.path.join(HERE, "libh264lab_b200.so")   # H264B200_LIB: developer A/B builds


class CreateParam(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "width", "height", "gop", "vbv_size_bytes", "vbv_overflow_empty_frame_flag",
        "vbv_underflow_stuffing_flag", "fine_rate_control_flag", "const_input_flag",
        "max_long_term_reference_frames", "enableNEON", "temporal_denoise_flag", "sps_id",
        "num_layers", "inter_layer_pred_flag")]


NALU_CB = C.CFUNCTYPE(None, C.POINTER(C.c_ubyte), C.c_int, C.c_void_p)


class RunParam(C.Structure):
    _fields_ = [("encode_speed", C.c_int), ("frame_type", C.c_int), ("long_term_idx_use", C.c_int),
                ("long_term_idx_update", C.c_int), ("desired_frame_bytes", C.c_int), ("qp_min", C.c_int),
                ("qp_max", C.c_int), ("desired_nalu_bytes", C.c_int), ("nalu_callback", NALU_CB),
                ("nalu_callback_token", C.c_void_p)]


class IoYuv(C.Structure):
    _fields_ = [("yuv", C.c_void_p * 3), ("stride", C.c_int * 3)]


class Library:
    def __init__(self, path=None):
        path = path or DEFAULT_LIB
        if not os.path.exists(path):
            raise RuntimeError("h264-lab_b200: %s not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                               "there is no CPU fallback" % path)
        self.path = path
        self.lib = l = C.CDLL(path)
        l.H264E_sizeof.argtypes = [C.POINTER(CreateParam), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        l.H264E_init.argtypes = [C.c_void_p, C.POINTER(CreateParam)]
        l.H264E_encode.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(RunParam), C.POINTER(IoYuv),
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
        l.H264E_close.argtypes = [C.c_void_p]
        l.H264E_close.restype = None
        l.H264E_get_recon.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.H264E_encode_batch.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.h264b200_backend_name.restype = C.c_char_p
        l.h264b200_launch_count.restype = C.c_long
        l.h264b200_last_timing.argtypes = [C.POINTER(C.c_float)]
        l.h264b200_last_timing.restype = None

    def backend(self):
        return self.lib.h264b200_backend_name().decode()

    def launch_count(self):
        return int(self.lib.h264b200_launch_count())


class Encoder:
    """One encoder session == one closed-GOP segment or one stream (fresh H264E_init)."""

    def __init__(self, library, width, height, gop, const_input=1, vbv_size_bytes=100000 // 8, **extra):
        self.L = library
        l = library.lib
        self.width, self.height = width, height
        self.w16, self.h16 = (width + 15) & ~15, (height + 15) & ~15
        self.cp = CreateParam(width=width, height=height, gop=gop, const_input_flag=const_input,
                              vbv_size_bytes=vbv_size_bytes, enableNEON=1, num_layers=1, **extra)
        sp, ss = C.c_int(0), C.c_int(0)
        err = l.H264E_sizeof(C.byref(self.cp), C.byref(sp), C.byref(ss))
        if err:
            raise RuntimeError("H264E_sizeof error %d" % err)
        self.sizeof_persist, self.sizeof_scratch = sp.value, ss.value
        self._persist = np.zeros(sp.value + 64, dtype=np.uint8)
        self._scratch = np.zeros(ss.value + 64, dtype=np.uint8)
        self.persist = (self._persist.ctypes.data + 63) & ~63
        self.scratch = (self._scratch.ctypes.data + 63) & ~63
        err = l.H264E_init(self.persist, C.byref(self.cp))
        if err:
            raise RuntimeError("H264E_init error %d (100 = no CUDA device: there is no CPU fallback)" % err)
        self.closed = False

    def run_param(self, qp=28, kbps=0, speed=0, frame_type=0):
        rp = RunParam()
        rp.frame_type = frame_type
        rp.encode_speed = speed
        if kbps:
            rp.desired_frame_bytes = kbps * 1000 // 8 // 30
            rp.qp_min, rp.qp_max = 10, 50
        else:
            rp.qp_min = rp.qp_max = qp
        return rp

    def io_yuv(self, frame):
        """frame: contiguous uint8 array of width*height*3/2 bytes (I420)."""
        w, h = self.width, self.height
        base = frame.ctypes.data
        yuv = IoYuv()
        yuv.yuv[0], yuv.yuv[1], yuv.yuv[2] = base, base + w * h, base + w * h * 5 // 4
        yuv.stride[0], yuv.stride[1], yuv.stride[2] = w, w // 2, w // 2
        return yuv

    def encode(self, frame, rp):
        yuv = self.io_yuv(frame)
        data, n = C.c_void_p(0), C.c_int(0)
        err = self.L.lib.H264E_encode(self.persist, self.scratch, C.byref(rp), C.byref(yuv), C.byref(data), C.byref(n))
        if err:
            raise RuntimeError("H264E_encode error %d" % err)
        return C.string_at(data.value, n.value)

    def recon(self):
        ysz = self.w16 * self.h16
        out = np.zeros(ysz * 3 // 2, dtype=np.uint8)
        err = self.L.lib.H264E_get_recon(self.persist, out.ctypes.data, out.ctypes.data + ysz, out.ctypes.data + ysz + ysz // 4)
        if err:
            raise RuntimeError("H264E_get_recon error %d" % err)
        return out

    def close(self):
        if not self.closed:
            self.L.lib.H264E_close(self.persist)
            self.closed = True

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def encode_sequence(library, frames, width, height, gop, qp=28, kbps=0, speed=0, want_recon=True, denoise=0,
                    empty_frames=0, stuffing=0):
    """Mirror of tests/refenc.encode_sequence on the B200 library: one session, frame by frame."""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    extra = {}
    if denoise:
        extra["temporal_denoise_flag"] = 1
    if empty_frames:
        extra["vbv_overflow_empty_frame_flag"] = 1
    if stuffing:
        extra["vbv_underflow_stuffing_flag"] = 1
    enc = Encoder(library, width, height, gop, **extra)
    rp = enc.run_param(qp=qp, kbps=kbps, speed=speed)
    out, sizes, recon = [], [], []
    for i in range(frames.shape[0]):
        f = frames[i].copy()
        bs = enc.encode(f, rp)
        out.append(bs)
        sizes.append(len(bs))
        if want_recon:
            recon.append(enc.recon())
    enc.close()
    return b"".join(out), np.array(sizes, dtype=np.int32), (np.stack(recon) if want_recon else None)


def encode_batch(library, encoders, frames, rps):
    """One frame for each of n sessions in a single device submission (H264E_encode_batch)."""
    n = len(encoders)
    yuvs = [e.io_yuv(f) for e, f in zip(encoders, frames)]
    P = (C.c_void_p * n)(*[e.persist for e in encoders])
    S = (C.c_void_p * n)(*[e.scratch for e in encoders])
    R = (C.c_void_p * n)(*[C.addressof(r) for r in rps])
    Y = (C.c_void_p * n)(*[C.addressof(y) for y in yuvs])
    D = (C.c_void_p * n)()
    N = (C.c_int * n)()
    err = library.lib.H264E_encode_batch(n, P, S, R, Y, D, N)
    if err:
        raise RuntimeError("H264E_encode_batch error %d" % err)
    return [C.string_at(D[i], N[i]) for i in range(n)]
