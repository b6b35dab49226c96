#!/usr/bin/env python3
"""developer helper: turn the ncu reports of tools/ncu_all.sh (gpurun_out/<tag>_*.ncu-rep) into the tracked
summaries under profiles/:
  <out>_step_kernels.json / .txt : one entry per kernel launch of one P-frame step (10 concurrent 1080p frames):
                                   duration, grid, registers, DRAM bytes, achieved DRAM GB/s and its share of
                                   the measured HBM peak, SM throughput, IPC, issue-slot use, instruction-cache
                                   hit rate, top stall reasons
  <out>_k_encode_rows_{1,10}stream_keymetrics.json + _details.txt : the dominant kernel
usage: tools/ncu_summarise.py <tag> <out-prefix>      e.g.  r38 profiles/r01b
"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, out = sys.argv[1], sys.argv[2]
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
hbm_peak = float(peaks.get("hbm_gbs", 6555.5))
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__icc_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "Tbyte": 1e12}
TUNIT = {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1.0, "second": 1.0, "nsecond": 1e-9}

def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]

def entry(hdr, units, r):
    H = {h: i for i, h in enumerate(hdr)}
    e = {"kernel": r[H["Kernel Name"]].split("(")[0]}
    for k in KEYS:
        if k in H:
            e[k] = [r[H[k]], units[H[k]]]
    stalls = {h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""): float(r[H[h]] or 0)
              for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
    e["stall_cycles_per_issue_top"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
    try:
        t = float(e["gpu__time_duration.sum"][0]) * TUNIT[e["gpu__time_duration.sum"][1]]
        b = sum(float(e[k][0]) * UNIT[e[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        e["duration_ms"] = t * 1e3
        e["dram_bytes"] = b
        e["dram_gbs"] = b / t / 1e9
        e["dram_frac_of_measured_hbm_peak"] = b / t / 1e9 / hbm_peak
    except Exception:
        pass
    return e

rep = os.path.join(ROOT, "gpurun_out", tag + "_step_allkernels.ncu-rep")
if os.path.exists(rep):
    hdr, units, rows = raw(rep)
    ents = [entry(hdr, units, r) for r in rows]
    json.dump({"workload": "one P-frame step of bench.py's workload: frame 3 of 10 concurrent 1080p segments (tools/batch_probe.py 10 4), "
                           "ncu --set full --clock-control none, every kernel launch of the step in launch order",
               "hbm_peak_gbs": hbm_peak, "launches": ents}, open(out + "_step_kernels.json", "w"), indent=1)
    tot = sum(e.get("duration_ms", 0) for e in ents)
    with open(out + "_step_kernels.txt", "w") as f:
        f.write("one P-frame step, 10 concurrent 1080p frames; per-launch ncu --set full (serialised, cold caches)\n")
        f.write("%-16s %9s %6s %5s %5s %9s %8s %6s %6s %6s %6s  %s\n" % ("kernel", "ms", "share", "grid", "regs", "DRAM MB", "GB/s", "%HBM", "IPC", "issue%", "I$hit", "top stalls (cycles per issue)"))
        for e in ents:
            g = lambda k: e.get(k, ["", ""])[0]
            f.write("%-16s %9.4f %5.1f%% %5s %5s %9.2f %8.1f %5.1f%% %6s %6s %6s  %s\n" % (
                e["kernel"], e.get("duration_ms", 0), 100 * e.get("duration_ms", 0) / tot if tot else 0, g("launch__grid_size"), g("launch__registers_per_thread"),
                e.get("dram_bytes", 0) / 1e6, e.get("dram_gbs", 0), 100 * e.get("dram_frac_of_measured_hbm_peak", 0),
                g("sm__inst_executed.avg.per_cycle_active")[:5], g("smsp__issue_active.avg.pct_of_peak_sustained_active")[:5], g("sm__icc_request_hit_rate.pct")[:5],
                " ".join("%s=%.1f" % kv for kv in e["stall_cycles_per_issue_top"].items())))
        f.write("total %.3f ms\n" % tot)
    print(open(out + "_step_kernels.txt").read())
for n in ("1stream", "10stream"):
    rep = os.path.join(ROOT, "gpurun_out", "%s_enc_%s.ncu-rep" % (tag, n))
    if not os.path.exists(rep):
        continue
    hdr, units, rows = raw(rep)
    H = {h: i for i, h in enumerate(hdr)}
    r = rows[0]
    km = {k: [r[H[k]], units[H[k]]] for k in KEYS if k in H}
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            km[h] = [r[H[h]], units[H[h]]]
    json.dump(km, open("%s_k_encode_rows_%s_keymetrics.json" % (out, n), "w"), indent=1)
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    open("%s_k_encode_rows_%s_ncu_full_details.txt" % (out, n), "w").write(det)
    print(n, km.get("gpu__time_duration.sum"), km.get("sm__inst_executed.avg.per_cycle_active"), km.get("sm__icc_request_hit_rate.pct"))
