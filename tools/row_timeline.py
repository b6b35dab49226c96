"""developer helper (GPU box, -DH264_FASTPROF variant): per-row timeline of sweep 0 of the last frame (ns)"""
import ctypes as C, importlib.util, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import content
spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "h264-lab_b200", "binding.py")); B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
L = B.Library(os.environ.get("H264B200_LIB"))
L.lib.H264E_b200_ctx.restype = C.c_void_p
w, h, n = 1920, 1080, 4
fr = content.panning(w, h, n, seed=1000)
nsess = int(sys.argv[1]) if len(sys.argv) > 1 else 1
encs = [B.Encoder(L, w, h, 60) for _ in range(nsess)]
rps = [e.run_param(qp=28) for e in encs]
enc = encs[0]
nmb = 120 * 68
for i in range(n):
    B.encode_batch(L, encs, [fr[i].copy() for _ in encs], rps)
prof = np.zeros((nmb * 20,), np.int32)
L.lib.h264b200_get_profile(C.c_void_p(L.lib.H264E_b200_ctx(C.c_void_p(enc.persist))), prof.ctypes.data_as(C.c_void_p))
r = prof[:68 * 8].reshape(68, 8).astype(np.int64)
t0 = r[:, 0].min()
print("row  start_us  dur_us  poll_us  ring_us  slow_us  fast  decide_us  publish_us  stage_us")
for y in range(68):
    print("%3d %9.1f %7.1f %8.1f %8.1f %8.1f %5d %9.1f %9.1f %9.1f" % (y, ((r[y, 0] - t0) & 0x7fffffff) / 1e3, r[y, 1] / 1e3, r[y, 2] / 1e3, r[y, 3] / 1e3, r[y, 4] / 1e3, r[y, 5], r[y, 6] / 1e3, r[y, 7] / 1e3, prof[68 * 8 + 300 + y] / 1e3))
f = prof[68 * 8:68 * 8 + 210].astype(np.int64)
print("follower: end %.1f us, waiting %.1f us; reaches row r at (us):" % (f[200] / 1e3, f[201] / 1e3), " ".join("%d:%.0f" % (y, f[2 * y] / 1e3) for y in range(0, 68, 6)))
