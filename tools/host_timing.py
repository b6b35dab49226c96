"""Developer tool (GPU box): where the host time of H264E_encode_batch goes (10 x 1080p segments per step)."""
import ctypes as C, importlib.util, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import content
spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "h264-lab_b200", "binding.py"))
B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
L = B.Library()
w, h, nseg, nfr = 1920, 1080, 10, 14
clips = [content.panning(w, h, nfr, seed=1000 + s) for s in range(nseg)]
encs = [B.Encoder(L, w, h, 60) for _ in range(nseg)]
rps = [e.run_param(qp=28) for e in encs]
t = (C.c_double * 3)()
tm = (C.c_float * 4)()
dev = 0.0
for i in range(nfr):
    if i == 2:
        L.lib.H264E_b200_host_timing(t); base = list(t); t0 = time.perf_counter(); dev = 0.0
    B.encode_batch(L, encs, [clips[s][i] for s in range(nseg)], rps)
    L.lib.h264b200_last_timing(tm); dev += tm[0]
L.lib.H264E_b200_host_timing(t)
k = nfr - 2
print("per step: wall %.2f ms | plan %.3f | submit + device + wait %.3f (device events %.3f) | NAL assembly + RC %.3f ms"
      % ((time.perf_counter() - t0) * 1e3 / k, (t[0] - base[0]) / k, (t[1] - base[1]) / k, dev / k, (t[2] - base[2]) / k))
