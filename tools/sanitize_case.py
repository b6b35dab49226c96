"""Developer tool (GPU box): small encodes that touch every code path once (for a sanitizer run where one is available, or as a quick smoke): CIF I+P with repair passes, cropped size,
temporal noise suppressor, prefetch path, a batch of two sessions.  No reference needed."""
import ctypes as C, importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import content
spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "h264-lab_b200", "binding.py"))
B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
L = B.Library()
for (kind, w, h, n, kw) in (("multi", 352, 288, 3, dict(qp=28)), ("panning", 366, 250, 3, dict(qp=24, denoise=1)), ("panning", 200, 120, 3, dict(kbps=300))):
    fr = getattr(content, "multi_motion" if kind == "multi" else "panning")(w, h, n)
    bs, sz, rec = B.encode_sequence(L, fr, w, h, n, **kw)
    print(kind, w, h, kw, len(bs))
w, h, n = 320, 240, 3
clips = [content.panning(w, h, n, seed=7 + s) for s in range(2)]
encs = [B.Encoder(L, w, h, 60) for _ in range(2)]
rps = [e.run_param(qp=30) for e in encs]
for t in range(n):
    if t + 1 < n:
        for e, c in zip(encs, clips):
            nxt = e.io_yuv(c[t + 1])
            L.lib.H264E_prefetch(C.c_void_p(e.persist), C.byref(nxt))
    B.encode_batch(L, encs, [c[t] for c in clips], rps)
print("batch + prefetch ok, hits", L.lib.h264b200_prefetch_hits())
