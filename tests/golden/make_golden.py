"""Generates tests/golden/manifest.json from the compiled reference (oracle/_ref).
Run where /root/reference exists:  python tests/golden/make_golden.py
The manifest pins input, bit stream and reconstruction digests of reference output."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cases  # noqa: E402
import refenc  # noqa: E402

GOLD = [c for c in cases.SMALL if c[0] in (
    "cif_i_only", "cif_ipp_qp28", "cif_speed2", "cif_speed10", "cif_kbps300", "noise_qp12", "crop_366x250",
    "small_48x32_inf_gop", "cif_denoise", "crop_denoise_366x250", "vbv_empty_denoise", "crop_skip_480x50_qp44", "crop_skip_640x50_rc")] + cases.CIF_FOREMAN_SUBSTITUTE[:1] + [("1080p_ipp", "panning", 1920, 1080, 3, 3, dict(qp=28))]

out = {"generator": "tests/golden/make_golden.py", "reference": "l646773422/h264-lab src/h264-lab.h via oracle/ref_harness.c, gcc -O2",
       "cases": []}
for name, kind, w, h, n, gop, kw in GOLD:
    frames = cases.make(kind, w, h, n)
    bs, sizes, rec, _ = refenc.encode_sequence(frames, w, h, gop, **kw)
    out["cases"].append(dict(name=name, kind=kind, width=w, height=h, frames=n, gop=gop, kw=kw,
                             input_md5=hashlib.md5(frames.tobytes()).hexdigest(),
                             bitstream_md5=hashlib.md5(bs).hexdigest(),
                             recon_md5=hashlib.md5(rec.tobytes()).hexdigest(),
                             frame_sizes=[int(s) for s in sizes]))
    print(name, len(bs))
json.dump(out, open(os.path.join(HERE, "manifest.json"), "w"), indent=1)
