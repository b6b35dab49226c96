"""Developer tool (emulation or GPU box): randomised parity sweep of the HOST layer's parameter handling and rate control --
create parameters (vbv size, the VBV flags, sps_id, gop incl. 0, odd and tiny sizes, unsupported combinations) and per-frame
run parameters (desired_frame_bytes, qp_min / qp_max ranges, speeds 0..10, frame types), H264E_set_vbv_state now and
then -- against the compiled reference called with the same structs.  Equal error codes count as agreement; features
this product reports as H264E_STATUS_UNSUPPORTED are not generated.  usage: stress_rc.py <seconds> [seed]"""
import os, sys, time, random, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import conftest, cases, refenc
B = conftest.load_binding()
L = B.Library(os.environ.get("H264B200_LIB") or os.path.join(ROOT, "h264-lab_b200", "libh264lab_b200.so"))
R = refenc.lib()
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t0 = time.time(); n = 0; nfr = 0; fails = 0
def aligned(nbytes):
    a = np.zeros(nbytes + 64, np.uint8); return a, (a.ctypes.data + 63) & ~63
while time.time() - t0 < budget:
    kind = rng.choice(["panning", "multi", "noise", "chess", "flat"])
    w = rng.choice([16, 18, 48, 100, 176, 200, 352, 366]); h = rng.choice([16, 18, 50, 120, 144, 250, 288])
    nf = rng.randint(3, 10)
    cpk = dict(width=w, height=h, gop=rng.choice([0, 1, 2, 5, 60]), vbv_size_bytes=rng.choice([0, 500, 12500, 100000, 1000000]),
               vbv_overflow_empty_frame_flag=rng.choice([0, 0, 1]), vbv_underflow_stuffing_flag=0, const_input_flag=1,
               sps_id=rng.choice([0, 0, 1, 3]), enableNEON=1, num_layers=1, temporal_denoise_flag=rng.choice([0, 0, 0, 1]))
    try: frames = cases.make(kind, w, h, nf)
    except Exception: continue
    sides = []
    for which, (lib, CP, fn) in enumerate([(R, refenc.CreateParam, ("ref_sizeof", "ref_init")), (L.lib, B.CreateParam, ("H264E_sizeof", "H264E_init"))]):
        cp = CP(**cpk); sp, ss = C.c_int(0), C.c_int(0)
        e1 = getattr(lib, fn[0])(C.byref(cp), C.byref(sp), C.byref(ss))
        st = dict(err_sizeof=e1, err_init=None)
        if not e1:
            st["pa"], st["p"] = aligned(sp.value); st["sa"], st["s"] = aligned(ss.value)
            st["err_init"] = getattr(lib, fn[1])(C.c_void_p(st["p"]), C.byref(cp))
        sides.append(st)
    ok = (sides[0]["err_sizeof"], sides[0]["err_init"]) == (sides[1]["err_sizeof"], sides[1]["err_init"])
    desc = dict(cpk, kind=kind, nf=nf); log = []
    if ok and not sides[0]["err_sizeof"] and not sides[0]["err_init"]:
        for t in range(nf):
            lo = rng.randint(0, 51); hi = rng.randint(lo, 51) if rng.random() < 0.8 else lo
            prm = dict(encode_speed=rng.choice([0, 0, 0, 1, 2, 3, 5, 8, 9, 10]), frame_type=rng.choice([0, 0, 0, 0, 6, 5, 2, 1]),
                       desired_frame_bytes=rng.choice([0, 0, 100, 400, 2000, 20000]), qp_min=lo, qp_max=hi)
            log.append(prm)
            vb = (rng.randint(0, 2000000), rng.randint(-200000, 200000)) if rng.random() < 0.1 else None
            res = []
            for which, (lib, RP, IY, fe, fv) in enumerate([(R, refenc.RunParam, refenc.IoYuv, "ref_encode", "ref_set_vbv_state"),
                                                           (L.lib, B.RunParam, B.IoYuv, "H264E_encode", "H264E_set_vbv_state")]):
                st = sides[which]
                if vb: getattr(lib, fv)(C.c_void_p(st["p"]), C.c_int(vb[0]), C.c_int(vb[1]))
                rp = RP(**prm); f = frames[t].copy(); yuv = IY()
                yuv.yuv[0], yuv.yuv[1], yuv.yuv[2] = f.ctypes.data, f.ctypes.data + w * h, f.ctypes.data + w * h * 5 // 4
                yuv.stride[0], yuv.stride[1], yuv.stride[2] = w, w // 2, w // 2
                data, nb = C.c_void_p(0), C.c_int(0)
                e = getattr(lib, fe)(C.c_void_p(st["p"]), C.c_void_p(st["s"]), C.byref(rp), C.byref(yuv), C.byref(data), C.byref(nb))
                res.append((e, C.string_at(data.value, nb.value) if not e else b""))
            nfr += 1
            if res[0] != res[1]:
                ok = False
                print("MISMATCH", desc, "frame", t, "run params so far", log, "errors ref/ours", res[0][0], res[1][0], "sizes", len(res[0][1]), len(res[1][1]), flush=True)
                break
            if res[0][0]: break
    elif not ok:
        print("MISMATCH (create)", desc, sides[0]["err_sizeof"], sides[0]["err_init"], sides[1]["err_sizeof"], sides[1]["err_init"], flush=True)
    if sides[1].get("p") and not sides[1]["err_init"]: L.lib.H264E_close(C.c_void_p(sides[1]["p"]))
    n += 1; fails += not ok
print("%d sessions, %d frames in %.0f s, %d mismatches" % (n, nfr, time.time() - t0, fails))
sys.exit(1 if fails else 0)
