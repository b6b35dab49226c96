"""Developer tool (GPU box or emulation): randomised parity sweep of INPUT LAYOUTS -- separate plane buffers with row
strides larger than the picture (padding bytes filled with noise), frame by frame through H264E_encode, against the
compiled reference given the very same buffers.  usage: stress_strides.py <seconds> [seed]"""
import os, sys, time, random, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import conftest, cases, refenc
B = conftest.load_binding()
L = B.Library(os.environ.get("H264B200_LIB") or os.path.join(ROOT, "h264-lab_b200", "libh264lab_b200.so"))
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
nrng = np.random.default_rng(rng.randint(0, 1 << 30))
t0 = time.time(); n = 0; fails = 0
def planes(frame, w, h, sy, sc):
    """three separately allocated planes with the given strides; padding = noise"""
    out = []
    off = 0
    for k, (pw, ph, st) in enumerate([(w, h, sy), (w // 2, h // 2, sc), (w // 2, h // 2, sc)]):
        buf = nrng.integers(0, 256, size=(ph, st), dtype=np.uint8)
        buf[:, :pw] = frame[off:off + pw * ph].reshape(ph, pw)
        off += pw * ph
        out.append(np.ascontiguousarray(buf))
    return out
while time.time() - t0 < budget:
    kind = rng.choice(["panning", "multi", "noise", "chess"])
    w = rng.choice([16, 48, 100, 176, 352, 366, 640]); h = rng.choice([16, 50, 144, 250, 288])
    nf = rng.randint(2, 5); gop = rng.choice([1, 3, 60]); qp = rng.choice([20, 28, 40])
    sy = w + rng.choice([0, 0, 16, 32, 64, 48]); sc = w // 2 + rng.choice([0, 0, 8, 16, 32, 24])
    try:
        frames = cases.make(kind, w, h, nf)
        rs = refenc.RefSession(w, h, gop)
    except Exception: continue
    enc = B.Encoder(L, w, h, gop)
    ok = True
    for t in range(nf):
        pl = planes(frames[t], w, h, sy, sc)
        outs = []
        for which in (0, 1):
            yuv = (refenc.IoYuv if which == 0 else B.IoYuv)()
            for k in range(3): yuv.yuv[k] = pl[k].ctypes.data; yuv.stride[k] = sy if k == 0 else sc
            data, nb = C.c_void_p(0), C.c_int(0)
            if which == 0:
                rp = refenc.RunParam(); rp.qp_min = rp.qp_max = qp
                err = rs.l.ref_encode(C.c_void_p(rs.persist), C.c_void_p(rs.scratch), C.byref(rp), C.byref(yuv), C.byref(data), C.byref(nb))
            else:
                rp = enc.run_param(qp=qp)
                err = L.lib.H264E_encode(enc.persist, enc.scratch, C.byref(rp), C.byref(yuv), C.byref(data), C.byref(nb))
            outs.append((err, C.string_at(data.value, nb.value) if not err else b""))
        if outs[0] != outs[1]:
            ok = False
            print("MISMATCH", kind, w, h, nf, gop, qp, "strides", sy, sc, "frame", t, "errors", outs[0][0], outs[1][0], flush=True)
            break
        if outs[0][0]: break
    enc.close(); n += 1; fails += not ok
print("%d sessions in %.0f s, %d mismatches" % (n, time.time() - t0, fails))
sys.exit(1 if fails else 0)
