"""Developer tool (GPU box): randomised parity sweep -- random picture sizes (incl. cropped ones), contents, GOP lengths,
fixed QP or rate control, single sessions and batches, every run compared byte for byte with the compiled reference
(oracle/_ref).  usage: stress_parity.py <seconds> [seed] [big]"""
import os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import conftest, cases, refenc
B = conftest.load_binding()
L = B.Library(os.environ.get("H264B200_LIB") or os.path.join(ROOT, "h264-lab_b200", "libh264lab_b200.so"))
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
big = len(sys.argv) > 3 and sys.argv[3] == "big"
t0 = time.time(); n = 0; fails = 0
while time.time() - t0 < budget:
    kind = rng.choice(["panning", "multi", "noise", "chess", "panning", "multi", "flat", "static", "static", "slow", "fastpan", "fastpan"])
    w = rng.choice([16, 32, 48, 100, 176, 200, 320, 352, 366, 640, 854, 1280, 1920])
    h = rng.choice([16, 32, 50, 120, 144, 180, 250, 288, 360, 480, 720, 1080])
    if big and w * h < 320 * 240: continue           # "big": pictures with many macroblock rows (the wavefront at scale)
    if not big and w * h > 1280 * 720 and rng.random() < 0.7: continue
    nf = rng.randint(2, 7) if w * h < 400 * 300 else rng.randint(2, 4)
    gop = rng.choice([1, 2, 3, nf, nf, 60])
    kw = dict(qp=rng.choice([10, 20, 28, 33, 40, 51])) if rng.random() < 0.7 else dict(kbps=rng.choice([100, 500, 3000]))
    if rng.random() < 0.15: kw["denoise"] = 1
    if rng.random() < 0.2: kw["speed"] = rng.choice([1, 2, 5, 9, 10])
    if rng.random() < 0.15: kw["empty_frames"] = 1
    if rng.random() < 0.1 and kw.get("kbps", 0) * 1000 // 240 < w * h // 4: kw["stuffing"] = 1      # (the reference writes past its scratch buffer when a filler NAL outgrows a tiny picture)
    if os.environ.get("STRESS_VERBOSE"): print("case", kind, w, h, nf, gop, kw, flush=True)
    try:
        if kind == "static":       # the same picture again and again (+ a little noise now and then): early skips everywhere
            base = cases.make("panning", w, h, 2)[0]
            frames = np.stack([base] * nf)
            if rng.random() < 0.5:
                nz = np.random.default_rng(rng.randint(0, 1 << 30)).integers(-1, 2, size=frames.shape)
                frames = np.clip(frames.astype(np.int16) + nz * (np.arange(nf)[:, None] % 2), 0, 255).astype(np.uint8)
        elif kind == "fastpan":    # large global motion (up to +-24 samples per frame): search range, vector limits, untabulated positions
            base = cases.make("panning", w, h, 2)[0]
            y0 = base[:w * h].reshape(h, w); u0 = base[w * h:w * h + (w // 2) * (h // 2)].reshape(h // 2, w // 2); v0 = base[w * h + (w // 2) * (h // 2):].reshape(h // 2, w // 2)
            dx, dy = rng.randint(-24, 24) & ~1, rng.randint(-24, 24) & ~1
            fl = []
            for t in range(nf):
                fl.append(np.concatenate([np.roll(y0, (dy * t, dx * t), (0, 1)).ravel(), np.roll(u0, (dy * t // 2, dx * t // 2), (0, 1)).ravel(), np.roll(v0, (dy * t // 2, dx * t // 2), (0, 1)).ravel()]))
            frames = np.stack(fl).astype(np.uint8)
        elif kind == "slow":       # every picture twice: skips and motion alternate
            base = cases.make("multi", w, h, (nf + 1) // 2)
            frames = np.repeat(base, 2, axis=0)[:nf]
        else:
            frames = cases.make(kind, w, h, nf)
        rbs, rsz, rrec, _ = refenc.encode_sequence(frames, w, h, gop, **kw)
    except Exception as e:          # a combination the reference itself refuses
        continue
    bs, sz, rec = B.encode_sequence(L, frames, w, h, gop, **kw)
    ok = bs == rbs and np.array_equal(rec, rrec)
    if ok and rng.random() < 0.3 and not (set(kw) & {"denoise", "empty_frames", "stuffing"}):
        # the same clip as a batch of 3 sessions in one submission
        encs = [B.Encoder(L, w, h, gop) for _ in range(3)]
        rps = [e.run_param(**{k: v for k, v in kw.items() if k in ("qp", "kbps", "speed")}) for e in encs]
        outs = [b"", b"", b""]
        for t in range(nf):
            res = B.encode_batch(L, encs, [frames[t].copy() for _ in encs], rps)
            for k in range(3): outs[k] += res[k]
        ok = all(o == rbs for o in outs)
        for e in encs: e.close()
    n += 1
    if not ok:
        fails += 1
        print("MISMATCH", kind, w, h, nf, gop, kw, flush=True)
print("%d random cases in %.0f s, %d mismatches" % (n, time.time() - t0, fails))
sys.exit(1 if fails else 0)
