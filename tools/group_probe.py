"""Developer tool: the bench workload (N concurrent 1080p sessions, input resident) driven by G host threads, each
submitting its own group of sessions on its own lane (stream pair).  Prints frames/s over the steady P frames.
usage: group_probe.py <sessions> <frames> <groups>"""
import ctypes as C, importlib.util, os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import content, numpy as np
spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "h264-lab_b200", "binding.py"))
B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
L = B.Library()
w, h, nseg, nfr, G = 1920, 1080, int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
clips = [np.ascontiguousarray(content.panning(w, h, nfr, seed=1000 + s)) for s in range(nseg)]
encs = [B.Encoder(L, w, h, 60) for _ in range(nseg)]
rps = [e.run_param(qp=28) for e in encs]
for e, c in zip(encs, clips):
    assert L.lib.H264E_preload(C.c_void_p(e.persist), nfr, C.c_void_p(c.ctypes.data)) == 0
groups = [list(range(g, nseg, G)) for g in range(G)]
bar = threading.Barrier(G + 1)
def work(g):
    ids = groups[g]; n = len(ids)
    yuvs = [B.IoYuv() for _ in ids]
    P = (C.c_void_p * n)(*[encs[s].persist for s in ids]); S = (C.c_void_p * n)(*[encs[s].scratch for s in ids])
    R = (C.c_void_p * n)(*[C.addressof(rps[s]) for s in ids]); Y = (C.c_void_p * n)(*[C.addressof(y) for y in yuvs])
    D = (C.c_void_p * n)(); N = (C.c_int * n)()
    for t in range(nfr):
        if t == 3: bar.wait()
        for y in yuvs: y.yuv[0] = None; y.stride[0] = t
        assert L.lib.H264E_encode_batch(n, P, S, R, Y, D, N) == 0
    bar.wait()
th = [threading.Thread(target=work, args=(g,)) for g in range(G)]
for t in th: t.start()
bar.wait(); t0 = time.perf_counter()
bar.wait(); dt = time.perf_counter() - t0
for t in th: t.join()
print("sessions %d groups %d: %.1f frames/s (%.2f ms per step of %d frames)" % (nseg, G, nseg * (nfr - 3) / dt, dt / (nfr - 3) * 1e3, nseg))
