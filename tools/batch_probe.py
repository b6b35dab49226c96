"""Developer tool: per-step timing / speculation statistics for a batch of sessions."""
import ctypes as C, importlib.util, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import content, numpy as np
spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "h264-lab_b200", "binding.py"))
B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
L = B.Library()
L.lib.H264E_b200_ctx.restype = C.c_void_p
w, h, nseg, nfr = 1920, 1080, int(sys.argv[1]), int(sys.argv[2])
denoise = len(sys.argv) > 3 and sys.argv[3] == "denoise"      # temporal noise suppressor in front of the encoder
clips = [content.panning(w, h, nfr, seed=1000 + s) for s in range(nseg)]
encs = [B.Encoder(L, w, h, 60, temporal_denoise_flag=1) if denoise else B.Encoder(L, w, h, 60) for _ in range(nseg)]
rps = [e.run_param(qp=28) for e in encs]
for t in range(nfr):
    fr = [clips[s][t].copy() for s in range(nseg)]
    t0 = time.perf_counter()
    B.encode_batch(L, encs, fr, rps)
    dt = (time.perf_counter() - t0) * 1e3
    tm = (C.c_float * 4)(); L.lib.h264b200_last_timing(tm)
    st = []
    for e in encs:
        a = (C.c_int * 4)(); L.lib.h264b200_ctx_stats(C.c_void_p(L.lib.H264E_b200_ctx(C.c_void_p(e.persist))), a); st.append(list(a)[:3])
    print("step %d wall %.1f ms dev %.1f enc %.1f df %.1f cavlc %.2f  stats(passes,reenc,checks) %s" % (t, dt, tm[0], tm[1], tm[2], tm[3], st))
