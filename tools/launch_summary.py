#!/usr/bin/env python3
"""developer helper: per-kernel time of the LAST step in an ncu launch list (tools/ncu_launches.sh):
usage: tools/launch_summary.py gpurun_out/<tag>_launches.csv [kernel that starts a step = k_frame_init]"""
import csv, sys, collections
rows = []
for r in csv.reader(open(sys.argv[1])):
    if len(r) > 10 and r[0].isdigit():
        rows.append(r)
hdr = None
for r in csv.reader(open(sys.argv[1])):
    if r and r[0] == "ID":
        hdr = r
        break
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
seq = [(r[ki].split("(")[0], float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[ui], 1e-6)) for r in rows]
starts = [i for i, (k, _) in enumerate(seq) if k == "k_frame_init"]
last = seq[starts[-1]:] if starts else seq
ends = [i for i, (k, _) in enumerate(last) if k == "k_gather_info"]
if ends: last = last[:ends[0] + 1]        # host-driven extra passes and whatever follows the step are not part of it
tot = sum(v for _, v in last)
print("last step: %d launches, %.3f ms of kernel time (serialised, cold caches)" % (len(last), tot))
for i, (k, v) in enumerate(last):
    print("  %2d %-22s %8.3f ms  %5.1f %%" % (i, k, v, 100 * v / tot))
agg = collections.OrderedDict()
for k, v in last:
    agg[k] = agg.get(k, 0) + v
print("by kernel:")
for k, v in sorted(agg.items(), key=lambda x: -x[1]):
    print("  %-22s %8.3f ms  %5.1f %%" % (k, v, 100 * v / tot))
