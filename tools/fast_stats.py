"""developer helper (GPU box): share of P macroblocks taken by the decide / work fast path, per frame, 1 session"""
import ctypes as C, importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import content
spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "h264-lab_b200", "binding.py")); B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
L = B.Library(os.environ.get("H264B200_LIB"))
L.lib.H264E_b200_ctx.restype = C.c_void_p
w, h, n = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 6
fr = content.panning(w, h, n, seed=1000)
enc = B.Encoder(L, w, h, 60); rp = enc.run_param(qp=28)
prev = [0] * 8
tm = (C.c_float * 8)()
for i in range(n):
    enc.encode(fr[i].copy(), rp)
    st = (C.c_int * 13)(); L.lib.h264b200_ctx_stats_ex(C.c_void_p(L.lib.H264E_b200_ctx(C.c_void_p(enc.persist))), st, 13)
    L.lib.h264b200_last_timing_ex(tm, 8)
    cur = list(st)[:8]
    d = [a - b for a, b in zip(cur, prev)]; prev = cur
    dbg = list(st)[8:13]
    print("frame %d: passes %d reenc %d (in the waves, one after the other: %d) checks %d fast %d slow %d | ms total %.2f sweep %.2f deblock %.2f sadmap %.2f me %.2f intra_check %.2f" % (i, d[0], d[1], dbg[0], d[2], d[4], d[5], tm[0], tm[1], tm[2], tm[4], tm[5], tm[6]))
