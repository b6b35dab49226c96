"""Developer tool (GPU box or emulation): randomised parity sweep of the EXTENSION entry points -- H264E_prefetch with
right and wrong predictions of the next frame, H264E_preload + frames taken from the resident clip (mixed with frames
from host memory), H264E_get_recon after every frame -- against the compiled reference driven frame by frame.
usage: stress_api.py <seconds> [seed]"""
import os, sys, time, random, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import conftest, cases, refenc
B = conftest.load_binding()
L = B.Library(os.environ.get("H264B200_LIB") or os.path.join(ROOT, "h264-lab_b200", "libh264lab_b200.so"))
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t0 = time.time(); n = 0; fails = 0
while time.time() - t0 < budget:
    kind = rng.choice(["panning", "multi", "noise", "chess"])
    w = rng.choice([48, 176, 352, 366, 640]); h = rng.choice([50, 144, 250, 288, 360])
    nf = rng.randint(3, 7); gop = rng.choice([1, 3, 60]); qp = rng.choice([20, 28, 40])
    try:
        frames = np.ascontiguousarray(cases.make(kind, w, h, nf))
        rbs, rsz, rrec, _ = refenc.encode_sequence(frames, w, h, gop, qp=qp)
    except Exception:
        continue
    enc = B.Encoder(L, w, h, gop)
    rp = enc.run_param(qp=qp)
    resident = rng.random() < 0.4
    if resident: assert L.lib.H264E_preload(C.c_void_p(enc.persist), nf, C.c_void_p(frames.ctypes.data)) == 0
    out = b""; ok = True
    keep = []                                  # buffers named by H264E_prefetch must stay untouched until they are used
    host = [frames[t].copy() for t in range(nf)]
    for t in range(nf):
        yuv = enc.io_yuv(host[t])
        if resident and rng.random() < 0.6:
            yuv.yuv[0] = None; yuv.stride[0] = t
        if t + 1 < nf and rng.random() < 0.6:
            nxt = t + 1 if rng.random() < 0.7 else rng.randrange(nf)       # a wrong prediction now and then
            ny = enc.io_yuv(host[nxt]); keep.append(ny)
            assert L.lib.H264E_prefetch(C.c_void_p(enc.persist), C.byref(ny)) == 0
        data, nb = C.c_void_p(0), C.c_int(0)
        err = L.lib.H264E_encode(enc.persist, enc.scratch, C.byref(rp), C.byref(yuv), C.byref(data), C.byref(nb))
        if err: ok = False; print("ERROR", err, flush=True); break
        out += C.string_at(data.value, nb.value)
        if rng.random() < 0.5 and not np.array_equal(enc.recon(), rrec[t]): ok = False; print("recon differs at frame", t, flush=True); break
    if ok and out != rbs: ok = False
    enc.close(); n += 1
    if not ok:
        fails += 1
        print("MISMATCH", kind, w, h, nf, gop, qp, "resident", resident, flush=True)
print("%d sessions in %.0f s, %d mismatches" % (n, time.time() - t0, fails))
sys.exit(1 if fails else 0)
