"""Checker-side developer tool (GPU box; lives under tests/ because it runs the oracle): byte-for-byte parity over a whole GOP and across an IDR at the bench's size:
1080p, GOP 60, 62 frames, one session; and the same with the temporal noise suppressor."""
import hashlib, importlib.util, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import content, refenc, numpy as np
spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "h264-lab_b200", "binding.py"))
B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
L = B.Library()
w, h, n, gop = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 62, 60
frames = content.panning(w, h, n, seed=4242)
for kw in (dict(qp=28), dict(kbps=8000), dict(qp=28, denoise=1)):
    t0 = time.time()
    rbs, rsz, _, _ = refenc.encode_sequence(frames, w, h, gop, want_recon=False, variant="_fast", **kw)
    t1 = time.time()
    bs, sz, _ = B.encode_sequence(L, frames, w, h, gop, want_recon=False, **kw)
    t2 = time.time()
    print(kw, "frames", n, "bytes", len(bs), "identical" if bs == rbs else "DIFFERENT", "first differing frame",
          next((i for i in range(n) if sz[i] != rsz[i]), None), "ref %.1fs ours %.1fs" % (t1 - t0, t2 - t1), hashlib.md5(bs).hexdigest())
    assert bs == rbs
print("long parity ok")
