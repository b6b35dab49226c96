"""Developer tool (GPU box or emulation): randomised parity sweep of HETEROGENEOUS batches -- 2..6 sessions of different
picture sizes, contents, GOP lengths (so that I and P frames meet in one submission), QPs / rate control and speeds
advance together through H264E_encode_batch; every session's stream is compared with the compiled reference.
usage: stress_batch.py <seconds> [seed]"""
import os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import conftest, cases, refenc
B = conftest.load_binding()
L = B.Library(os.environ.get("H264B200_LIB") or os.path.join(ROOT, "h264-lab_b200", "libh264lab_b200.so"))
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t0 = time.time(); nb = 0; ns = 0; fails = 0
while time.time() - t0 < budget:
    k = rng.randint(2, 6)
    sess = []
    for _ in range(k):
        kind = rng.choice(["panning", "multi", "noise", "chess", "flat"])
        w = rng.choice([32, 100, 176, 352, 366, 640, 1280]); h = rng.choice([32, 50, 144, 250, 288, 360, 720])
        nf = rng.randint(2, 6); gop = rng.choice([1, 2, 3, 60])
        kw = dict(qp=rng.choice([12, 22, 28, 33, 44])) if rng.random() < 0.7 else dict(kbps=rng.choice([200, 2000]))
        if rng.random() < 0.25: kw["speed"] = rng.choice([1, 2, 9, 10])
        frames = cases.make(kind, w, h, nf)
        rbs, _, _, _ = refenc.encode_sequence(frames, w, h, gop, want_recon=False, **kw)
        enc = B.Encoder(L, w, h, gop)
        sess.append(dict(desc=(kind, w, h, nf, gop, kw), frames=frames, nf=nf, ref=rbs, enc=enc, rp=enc.run_param(**kw), out=b""))
    for t in range(max(s["nf"] for s in sess)):
        act = [s for s in sess if t < s["nf"]]
        res = B.encode_batch(L, [s["enc"] for s in act], [s["frames"][t].copy() for s in act], [s["rp"] for s in act])
        for s, r in zip(act, res): s["out"] += r
    for s in sess:
        ns += 1
        if s["out"] != s["ref"]:
            fails += 1
            print("MISMATCH", s["desc"], "in a batch of", [x["desc"] for x in sess], flush=True)
        s["enc"].close()
    nb += 1
print("%d batches, %d sessions in %.0f s, %d mismatches" % (nb, ns, time.time() - t0, fails))
sys.exit(1 if fails else 0)
