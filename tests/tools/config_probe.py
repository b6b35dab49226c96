"""Checker-side developer tool (GPU box; lives under tests/ because it runs the oracle): parity + throughput of the other BASELINE configs at reduced frame counts.
  config 3: 2160p rate-controlled closed-GOP segments in one batch
  config 4: 1080p all-intra
  config 5: 64 concurrent 720p streams
Each result is compared byte for byte with the compiled reference (oracle/_ref) on the first streams."""
import importlib.util, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import content, refenc
spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "h264-lab_b200", "binding.py"))
B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
L = B.Library()

def run(name, w, h, nstreams, nframes, gop, check, **kw):
    clips = [content.panning(w, h, nframes, seed=500 + s) for s in range(nstreams)]
    encs = [B.Encoder(L, w, h, gop) for _ in range(nstreams)]
    rps = [e.run_param(**kw) for e in encs]
    outs = [b"" for _ in range(nstreams)]
    t0 = time.perf_counter()
    for t in range(nframes):
        res = B.encode_batch(L, encs, [c[t].copy() for c in clips], rps)
        for s in range(nstreams):
            outs[s] += res[s]
    dt = time.perf_counter() - t0
    ok = True
    for s in range(check):
        rbs, _, _, _ = refenc.encode_sequence(clips[s], w, h, gop, **kw)
        ok = ok and (rbs == outs[s])
    for e in encs:
        e.close()
    print("%-40s %4dx%-4d %2d streams x %2d frames: %7.1f frames/s (%6.1f MP/s), parity on %d streams: %s" % (
        name, w, h, nstreams, nframes, nstreams * nframes / dt, nstreams * nframes * w * h / dt / 1e6, check, "OK" if ok else "MISMATCH"), flush=True)

run("config 4: 1080p all-intra", 1920, 1080, 10, 4, 1, 2, qp=28)
run("config 5: 64 x 720p IPPP", 1280, 720, 64, 6, 60, 3, qp=28)
run("config 3: 2160p rate-controlled GOP shards", 3840, 2160, 4, 5, 30, 1, kbps=20000)
run("config 2 (bench): 1080p IPPP", 1920, 1080, 10, 8, 60, 1, qp=28)
