#!/usr/bin/env python
"""Algorithmic integer work of the macroblock path per macroblock (SURVEY.md 8(d) "Algorithmic integer work"): the
UNMODIFIED reference is run on each bench workload's content with one counter per leaf-function call
(oracle/opcount.sed -> oracle/_ref/libh264ref_count.so; the counting rules are in that file) and the totals are
divided by the macroblocks coded.  Data-dependent (the searches stop where they stop), so it is measured on the very
clips bench.py encodes.  Writes profiles/r02_opcount.json, which bench.py's roofline reads.
Needs /root/reference (build container); TEST/MEASUREMENT infrastructure, never shipped."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import refenc  # noqa: E402
import bench  # noqa: E402

CATS = ["sad", "sixtap_mac", "qpel_avg", "chroma_mc_mac", "fwd_transform", "inv_transform", "quant_mul", "intra_pred_sad", "deblock"]


def counts(l):
    a = (C.c_longlong * 16)()
    l.ref_get_opcounts(a)
    return [int(x) for x in a[:len(CATS)]]


def main():
    l = refenc.lib("_count")
    out = {}
    for name, cfg in bench.CONFIGS.items():
        w, h = cfg["w"], cfg["h"]
        nmb = ((w + 15) // 16) * ((h + 15) // 16)
        nfr = 4 if cfg["gop"] != 1 else 2
        clip = bench.make_clip(cfg, 0, nfr)
        kw = dict(kbps=cfg["kbps"]) if cfg["kbps"] else dict(qp=cfg["qp"])
        # I frame alone, then I + P frames: the difference is the P frames' work
        l.ref_reset_opcounts()
        refenc.encode_sequence(clip[:1], w, h, cfg["gop"], want_recon=False, variant="_count", **kw)
        ci = counts(l)
        row = {"ops_per_mb_i": sum(ci) / nmb, "by_category_i": {k: v / nmb for k, v in zip(CATS, ci)}}
        if cfg["gop"] != 1:
            l.ref_reset_opcounts()
            refenc.encode_sequence(clip, w, h, cfg["gop"], want_recon=False, variant="_count", **kw)
            ca = counts(l)
            cp = [(a - b) / (nfr - 1) for a, b in zip(ca, ci)]
            row.update({"ops_per_mb_p": sum(cp) / nmb, "by_category_p": {k: v / nmb for k, v in zip(CATS, cp)}})
        else:
            row.update({"ops_per_mb_p": row["ops_per_mb_i"], "by_category_p": row["by_category_i"]})
        row["sample"] = "%d frames of unit 0 (%dx%d), counting rules: oracle/opcount.sed" % (nfr, w, h)
        out[name] = row
        print(name, "P: %.0f ops/MB  I: %.0f ops/MB" % (row["ops_per_mb_p"], row["ops_per_mb_i"]))
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_opcount.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
