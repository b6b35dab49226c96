"""ctypes access to oracle/_ref/libh264ref*.so (the compiled, unmodified reference).

TEST INFRASTRUCTURE: imported by tests/, bench.py's cpu_baseline / --impl reference
leg and __graft_entry__.smoke() only.
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


class CreateParam(C.Structure):
    # H:83-172 (default build: H264E_SVC_API=1, H264E_MAX_THREADS=0) -> 14 ints
    _fields_ = [(n, C.c_int) for n in (
        "width", "height", "gop", "vbv_size_bytes", "vbv_overflow_empty_frame_flag",
        "vbv_underflow_stuffing_flag", "fine_rate_control_flag", "const_input_flag",
        "max_long_term_reference_frames", "enableNEON", "temporal_denoise_flag", "sps_id",
        "num_layers", "inter_layer_pred_flag")]


NALU_CB = C.CFUNCTYPE(None, C.POINTER(C.c_ubyte), C.c_int, C.c_void_p)


class RunParam(C.Structure):
    # H:177-226
    _fields_ = [("encode_speed", C.c_int), ("frame_type", C.c_int), ("long_term_idx_use", C.c_int),
                ("long_term_idx_update", C.c_int), ("desired_frame_bytes", C.c_int), ("qp_min", C.c_int),
                ("qp_max", C.c_int), ("desired_nalu_bytes", C.c_int), ("nalu_callback", NALU_CB),
                ("nalu_callback_token", C.c_void_p)]


class IoYuv(C.Structure):
    # H:231-237
    _fields_ = [("yuv", C.c_void_p * 3), ("stride", C.c_int * 3)]


assert C.sizeof(CreateParam) == 56 and C.sizeof(RunParam) == 48 and C.sizeof(IoYuv) == 40

_libs = {}


def have_ref(variant=""):
    return os.path.exists(os.path.join(REF_DIR, "libh264ref%s.so" % variant))


def lib(variant=""):
    if variant not in _libs:
        path = os.path.join(REF_DIR, "libh264ref%s.so" % variant)
        l = C.CDLL(path)
        l.ref_encode_sequence.restype = C.c_long
        l.ref_encode_sequence.argtypes = [C.c_int] * 7 + [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p,
                                                         C.c_void_p, C.POINTER(C.c_double)]
        if hasattr(l, "ref_encode_sequence_ex"):
            l.ref_encode_sequence_ex.restype = C.c_long
            l.ref_encode_sequence_ex.argtypes = [C.c_int] * 8 + [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p,
                                                                C.c_void_p, C.POINTER(C.c_double)]
        _libs[variant] = l
    return _libs[variant]


def encode_sequence(frames, width, height, gop, qp=28, kbps=0, speed=0, want_recon=True, variant="", denoise=0,
                    empty_frames=0, stuffing=0):
    """frames: uint8 [n, w*h*3/2].  Returns (bitstream bytes, sizes[n], recon [n, W16*H16*3/2] or None, seconds)."""
    l = lib(variant)
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    n = frames.shape[0]
    w16, h16 = (width + 15) & ~15, (height + 15) & ~15
    cap = n * (w16 * h16 // 256) * 396 * 3 // 2 + 4096
    out = np.zeros(cap, dtype=np.uint8)
    sizes = np.zeros(n, dtype=np.int32)
    recon = np.zeros((n, w16 * h16 * 3 // 2), dtype=np.uint8) if want_recon else None
    secs = C.c_double(0)
    flags = (1 if denoise else 0) | (2 if empty_frames else 0) | (4 if stuffing else 0)
    if flags:
        tot = l.ref_encode_sequence_ex(width, height, gop, qp, kbps, speed, flags, n, frames.ctypes.data, out.ctypes.data, cap,
                                       sizes.ctypes.data, recon.ctypes.data if want_recon else None, C.byref(secs))
    else:
        tot = l.ref_encode_sequence(width, height, gop, qp, kbps, speed, n, frames.ctypes.data, out.ctypes.data, cap,
                                    sizes.ctypes.data, recon.ctypes.data if want_recon else None, C.byref(secs))
    if tot < 0:
        raise RuntimeError("reference encoder error %d" % -tot)
    return out[:tot].tobytes(), sizes, recon, secs.value


class RefSession:
    """The reference's public API driven frame by frame (ref_sizeof / ref_init / ref_encode), for tests that need
    per-frame run parameters (frame types) or in-place reconstruction (const_input_flag = 0)."""

    def __init__(self, width, height, gop, const_input=1, variant="", **extra):
        self.l = lib(variant)
        self.width, self.height = width, height
        self.cp = CreateParam(width=width, height=height, gop=gop, const_input_flag=const_input,
                              vbv_size_bytes=100000 // 8, enableNEON=1, num_layers=1, **extra)
        sp, ss = C.c_int(), C.c_int()
        err = self.l.ref_sizeof(C.byref(self.cp), C.byref(sp), C.byref(ss))
        if err:
            raise RuntimeError("ref_sizeof error %d" % err)
        self._persist = np.zeros(sp.value + 64, np.uint8)
        self._scratch = np.zeros(ss.value + 64, np.uint8)
        self.persist = (self._persist.ctypes.data + 63) & ~63
        self.scratch = (self._scratch.ctypes.data + 63) & ~63
        self.l.ref_init(C.c_void_p(self.persist), C.byref(self.cp))

    def encode(self, frame, qp=28, kbps=0, speed=0, frame_type=0):
        """frame: contiguous uint8 I420 array (overwritten with the reconstruction when const_input_flag = 0)."""
        w, h = self.width, self.height
        rp = RunParam()
        rp.frame_type, rp.encode_speed = frame_type, speed
        if kbps:
            rp.desired_frame_bytes, rp.qp_min, rp.qp_max = kbps * 1000 // 8 // 30, 10, 50
        else:
            rp.qp_min = rp.qp_max = qp
        yuv = IoYuv()
        base = frame.ctypes.data
        yuv.yuv[0], yuv.yuv[1], yuv.yuv[2] = base, base + w * h, base + w * h * 5 // 4
        yuv.stride[0], yuv.stride[1], yuv.stride[2] = w, w // 2, w // 2
        data, n = C.c_void_p(0), C.c_int(0)
        err = self.l.ref_encode(C.c_void_p(self.persist), C.c_void_p(self.scratch), C.byref(rp), C.byref(yuv),
                                C.byref(data), C.byref(n))
        if err:
            raise RuntimeError("ref_encode error %d" % err)
        return C.string_at(data.value, n.value)
