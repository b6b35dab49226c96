"""Developer tool: per-macroblock phase timing on the GPU (needs libh264lab_b200_prof.so,
`make -C h264-lab_b200 libh264lab_b200_prof.so`).  Prints, for the last frame encoded, the
distribution of macroblock latencies by type and the share of each phase."""
import ctypes as C
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import content  # noqa: E402

spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "h264-lab_b200", "binding.py"))
B = importlib.util.module_from_spec(spec)
spec.loader.exec_module(B)
L = B.Library(os.path.join(ROOT, "h264-lab_b200", "libh264lab_b200_prof.so"))
L.lib.H264E_b200_ctx.restype = C.c_void_p
w, h, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
kind = sys.argv[4] if len(sys.argv) > 4 else "panning"
fr = getattr(content, kind)(w, h, n)
nsess = int(sys.argv[5]) if len(sys.argv) > 5 else 1      # concurrent sessions (profile shows session 0)
encs = [B.Encoder(L, w, h, 60) for _ in range(nsess)]
enc = encs[0]
rps = [e.run_param(qp=28) for e in encs]
rp = rps[0]
nmb = ((w + 15) // 16) * ((h + 15) // 16)
names = ["load", "win", "a_publish", "a_skiptest", "s16_qpel7", "wait_tasks", "decide", "tq_w0", "a_candlist", "a_cands", "tq_join", "record"]
for i in range(n):
    B.encode_batch(L, encs, [fr[i].copy() for _ in encs], rps)
    prof = np.zeros((nmb, 20), np.int32)
    L.lib.h264b200_get_profile(C.c_void_p(L.lib.H264E_b200_ctx(C.c_void_p(enc.persist))), prof.ctypes.data_as(C.c_void_p))
    sub = prof[:, 19].astype(np.int64)
    s16 = np.stack([sub & 0xFFFF, (sub >> 16) & 0xFFFF], 1)      # 16x16 search: integer part (incl. wait for its start), copy + half-sample plane fetch
    tot = prof[:, :12].sum(1) + s16.sum(1)
    tm = (C.c_float * 4)()
    L.lib.h264b200_last_timing(tm)
    print("frame %d: k_encode %.2f ms; MB cycles mean %.0f p50 %.0f p90 %.0f p99 %.0f max %.0f  (sum/1.9GHz = %.1f ms serial)" % (
        i, tm[1], tot.mean(), np.percentile(tot, 50), np.percentile(tot, 90), np.percentile(tot, 99), tot.max(), tot.sum() / 1.9e6))
    # who decides the join time of the slow macroblocks (critical path of the wavefront)?
    arr = prof[:, 12:16]
    late = arr.argmax(1)
    join = arr.max(1)
    thr = np.percentile(tot, 90)
    slow = tot >= thr
    print("   join time mean %.0f; last warp to arrive: %s ; among slowest 10%% (>= %.0f cyc): %s, their mean arrive %s" % (
        join.mean(), np.bincount(late, minlength=4).tolist(), thr, np.bincount(late[slow], minlength=4).tolist(), arr[slow].mean(0).astype(int).tolist()))
    for t in sorted(set(prof[:, 16].tolist())):
        m = prof[:, 16] == t
        ph = prof[m][:, :12].mean(0)
        print("   type %2d: %5d MBs, mean %7.0f cyc | " % (t, m.sum(), tot[m].mean()) + " ".join("%s %.0f" % (names[k], ph[k]) for k in range(12) if names[k] != "-")
              + " s16_int %.0f s16_fetch %.0f" % tuple(s16[m].mean(0))
              + " | arrive w0..3 %s chroma w2,w3 %s" % (prof[m][:, 12:16].mean(0).astype(int).tolist(), prof[m][:, 17:19].mean(0).astype(int).tolist()))
