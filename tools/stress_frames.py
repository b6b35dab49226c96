"""Developer tool (GPU box or emulation): randomised FRAME-BY-FRAME parity sweep -- every frame of a session gets its own
run parameters: frame type (default / key / I / P / droppable), QP (changes between frames exercise the reference's
quantiser-table update rule, H:5839), rate control on or off, speed; const_input_flag 0 (reconstruction written back into
the caller's frame) or 1; cropped and uncropped sizes.  Each coded frame (and the in-place reconstruction) is compared
with the compiled reference driven the same way; equal error codes count as agreement.
usage: stress_frames.py <seconds> [seed]"""
import os, sys, time, random, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import conftest, cases, refenc
B = conftest.load_binding()
L = B.Library(os.environ.get("H264B200_LIB") or os.path.join(ROOT, "h264-lab_b200", "libh264lab_b200.so"))
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
TYPES = [0, 0, 0, 0, 6, 5, 2, 2, 1, 1]           # DEFAULT, KEY, I, P, DROPPABLE
t0 = time.time(); n = 0; nfr = 0; fails = 0
def err_of(e):
    m = re.search(r"error (\d+)", str(e)); return int(m.group(1)) if m else -1
while time.time() - t0 < budget:
    kind = rng.choice(["panning", "multi", "noise", "chess", "flat"])
    w = rng.choice([16, 48, 100, 176, 200, 352, 366, 640]); h = rng.choice([16, 50, 120, 144, 250, 288, 360])
    nf = rng.randint(3, 9); gop = rng.choice([0, 1, 3, 4, 60]); const_input = rng.choice([1, 1, 0])
    if const_input == 0 and (w % 16 or h % 16): const_input = 1       # the reference asks for const_input_flag with cropped sizes (H:6279)
    extra = {}
    if rng.random() < 0.15: extra["temporal_denoise_flag"] = 1
    try:
        frames = cases.make(kind, w, h, nf)
        rs = refenc.RefSession(w, h, gop, const_input=const_input, **extra)
    except Exception:
        continue
    enc = B.Encoder(L, w, h, gop, const_input=const_input, **extra)
    ok = True; desc = (kind, w, h, nf, gop, const_input, extra); log = []
    for t in range(nf):
        ft = rng.choice(TYPES); qp = rng.choice([12, 20, 28, 28, 33, 44, 51]); kbps = rng.choice([0, 0, 0, 200, 2000]); speed = rng.choice([0, 0, 0, 2, 9])
        log.append((ft, qp, kbps, speed))
        fa, fb = frames[t].copy(), frames[t].copy()
        ea = eb = 0; ra = rb = b""
        try: ra = rs.encode(fa, qp=qp, kbps=kbps, speed=speed, frame_type=ft)
        except RuntimeError as e: ea = err_of(e)
        try: rb = enc.encode(fb, enc.run_param(qp=qp, kbps=kbps, speed=speed, frame_type=ft))
        except RuntimeError as e: eb = err_of(e)
        nfr += 1
        if ea != eb or ra != rb or (const_input == 0 and not ea and not np.array_equal(fa, fb)):
            ok = False
            print("MISMATCH", desc, "frame", t, "params (type, qp, kbps, speed) so far", log, "errors ref/ours", ea, eb, "sizes", len(ra), len(rb), flush=True)
            break
        if ea: break
    enc.close()
    n += 1
    fails += not ok
print("%d sessions, %d frames in %.0f s, %d mismatches" % (n, nfr, time.time() - t0, fails))
sys.exit(1 if fails else 0)
