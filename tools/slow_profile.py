"""Developer tool (GPU box, libh264lab_b200_prof.so): phase cycles of the macroblocks that took the COMPLETE path
(encode_mb) in the last P frame -- the fast path does not touch the profile rows, so a row that changed since the
previous frame belongs to a macroblock encode_mb has seen (sweep 0 or a repair)."""
import ctypes as C, importlib.util, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import content
spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "h264-lab_b200", "binding.py")); B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
L = B.Library(os.path.join(ROOT, "h264-lab_b200", "libh264lab_b200_prof.so"))
L.lib.H264E_b200_ctx.restype = C.c_void_p
w, h, n = 1920, 1080, 5
nsess = int(sys.argv[1]) if len(sys.argv) > 1 else 10
fr = content.panning(w, h, n, seed=1000)
encs = [B.Encoder(L, w, h, 60) for _ in range(nsess)]
rps = [e.run_param(qp=28) for e in encs]
nmbx, nmby = 120, 68
nmb = nmbx * nmby
names = ["load", "win", "a_publish", "a_skiptest", "s16_qpel7", "wait_tasks", "decide", "tq_w0", "a_candlist", "a_cands", "tq_join", "record"]
prev = None
for i in range(n):
    B.encode_batch(L, encs, [fr[i].copy() for _ in encs], rps)
    prof = np.zeros((nmb, 20), np.int32)
    L.lib.h264b200_get_profile(C.c_void_p(L.lib.H264E_b200_ctx(C.c_void_p(encs[0].persist))), prof.ctypes.data_as(C.c_void_p))
    if prev is not None and i == n - 1:
        ch = (prof != prev).any(1)
        sub = prof[:, 19].astype(np.int64)
        s16 = np.stack([sub & 0xFFFF, (sub >> 16) & 0xFFFF], 1)
        tot = prof[:, :12].sum(1) + s16.sum(1)
        print("frame %d: %d macroblocks through encode_mb; cycles mean %.0f p50 %.0f p90 %.0f max %.0f" % (i, ch.sum(), tot[ch].mean(), np.percentile(tot[ch], 50), np.percentile(tot[ch], 90), tot[ch].max()))
        rows = np.bincount(np.nonzero(ch)[0] // nmbx, minlength=nmby)
        print("per row:", " ".join("%d:%d" % (y, rows[y]) for y in range(nmby) if rows[y]))
        for t in sorted(set(prof[ch][:, 16].tolist())):
            m = ch & (prof[:, 16] == t)
            ph = prof[m][:, :12].mean(0)
            print("   type %2d: %5d MBs, mean %7.0f cyc | " % (t, m.sum(), tot[m].mean()) + " ".join("%s %.0f" % (names[k], ph[k]) for k in range(12))
                  + " s16_walk %.0f s16_probes %.0f" % tuple(s16[m].mean(0)) + " | arrive w0..3 %s" % (prof[m][:, 12:16].mean(0).astype(int).tolist()))
        # the rectangle rows only
        yy = np.arange(nmb) // nmbx
        m = ch & (yy >= 17) & (yy <= 30)
        if m.any():
            ph = prof[m][:, :12].mean(0)
            print("   rows 17..30: %d MBs, mean %.0f cyc | " % (m.sum(), tot[m].mean()) + " ".join("%s %.0f" % (names[k], ph[k]) for k in range(12)) + " | types %s" % np.bincount(prof[m][:, 16] + 4).tolist())
    prev = prof
