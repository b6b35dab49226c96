#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 H.264 macroblock-encode path.

Default workload = BASELINE.json configs[1] ("c2"): synthetic 1080p YUV420, 600 frames, IPPP GOP=60, fixed
QP 28, on ONE B200 = 10 closed-GOP segments of 60 frames, one fresh encoder session per segment (the reference
run once per segment is the golden output).  A "step" is one pass of the hot path over one batch: frame t of
every segment (n_units frames) in one device submission.  The default `--steps 60 --warmup 3` times a WHOLE GOP
of every segment: P frames 3..59, the IDR that opens the next GOP, and two more P frames.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--config c1|c2|c3|c4|c5] [--scaling weak|strong]

--config: c1 CIF 352x288 GOP 20 QP 33 (T:10-11; synthetic foreman substitute), 15 segments;  c2 (default);
          c3 2160p GOP 30 --kbps 20000 (rate control), 10 segments;  c4 1080p all-intra (GOP 1) QP 28, 10 streams;
          c5 64 independent 720p streams GOP 60 QP 28.
N > 1: launched by torchrun, one rank per GPU, no data-path collective (torch.distributed only for the barrier and
the max-over-ranks time).  --scaling weak (default): every rank encodes its own n_units units;  strong: the config's
n_units are split over the ranks (unit k -> rank k mod N).

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = frames/s through the public
C API with pinned host buffers (H2D of every input, D2H of every payload inside the timed region; the copy of frame
t + 1 is started with H264E_prefetch before frame t is encoded).  `parity`: AFTER the timed loops every byte both runs
produced is compared with the unmodified reference (oracle/_ref) run on the same frames; a mismatch suppresses
`value`.  `roofline`: the path is integer-issue / latency bound, so the headline fraction is algorithmic integer
operations (SURVEY 8(d) counting rules, measured with oracle/_ref/libh264ref_count.so) per second against the measured
packed-integer issue peak (tools/ubench/intpeak); the HBM view sits beside it.  `cpu_baseline` = the unmodified
reference on the host cores, one process per unit.  `--impl reference` times that reference arm alone.
"""
import argparse
import ctypes as C
import hashlib
import importlib.util
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIGS = {
    # name: width, height, gop, qp, kbps, units, seed0, description
    "c1": dict(w=352, h=288, gop=20, qp=33, kbps=0, units=15, seed0=2000,
               desc="synthetic CIF 352x288 (foreman substitute), 300 frames, IPPP GOP=20, fixed QP 33 = 15 closed-GOP segments"),
    "c2": dict(w=1920, h=1080, gop=60, qp=28, kbps=0, units=10, seed0=1000,
               desc="synthetic 1080p YUV420, 600 frames, IPPP GOP=60, fixed QP 28 = 10 closed-GOP segments"),
    "c3": dict(w=3840, h=2160, gop=30, qp=0, kbps=20000, units=10, seed0=3000,
               desc="synthetic 2160p YUV420, 300 frames, IPPP GOP=30, rate-controlled 20000 kbps = 10 closed-GOP segments"),
    "c4": dict(w=1920, h=1080, gop=1, qp=28, kbps=0, units=10, seed0=1000,
               desc="synthetic 1080p YUV420, intra-only (all IDR), fixed QP 28, 10 streams"),
    "c5": dict(w=1280, h=720, gop=60, qp=28, kbps=0, units=64, seed0=0,
               desc="64 concurrent independent synthetic 720p streams, IPPP GOP=60, fixed QP 28"),
}


def load_binding():
    spec = importlib.util.spec_from_file_location("h264lab_binding", os.path.join(ROOT, "h264-lab_b200", "binding.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_clip(cfg, unit, nframes):
    """Deterministic synthetic clip of one unit (tests/content.py, seed per unit)."""
    import content
    return content.panning(cfg["w"], cfg["h"], nframes, seed=cfg["seed0"] + unit)


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------
# reference arm (CPU): one process per unit, persistent encoder, stepped by the parent
# --------------------------------------------------------------------------------------
def _pick_ref_variant(refenc):
    """The fastest build of the unmodified reference this host can run: -O3 -march=x86-64-v3 when the CPU has AVX2/BMI2/FMA
    (the .so is compiled in the build container and travels; -march=native of THAT machine might not run here)."""
    try:
        flags = open("/proc/cpuinfo").read()
    except Exception:
        flags = ""
    if refenc.have_ref("_fast_v3") and all((" " + f) in flags for f in ("avx2", "bmi2", "fma", "movbe")):
        return "_fast_v3"
    return "_fast" if refenc.have_ref("_fast") else ""


def _ref_worker(cfg, unit, nframes, conn, core):
    import numpy as np
    import refenc
    if core is not None:
        try:
            os.sched_setaffinity(0, {core})
        except Exception:
            pass
    variant = _pick_ref_variant(refenc)
    l = refenc.lib(variant)
    W, H = cfg["w"], cfg["h"]
    clip = make_clip(cfg, unit, nframes)
    cp = refenc.CreateParam(width=W, height=H, gop=cfg["gop"], const_input_flag=1, vbv_size_bytes=100000 // 8, enableNEON=1, num_layers=1)
    sp, ss = C.c_int(), C.c_int()
    l.ref_sizeof(C.byref(cp), C.byref(sp), C.byref(ss))
    persist = np.zeros(sp.value + 64, np.uint8)
    scratch = np.zeros(ss.value + 64, np.uint8)
    pp = (persist.ctypes.data + 63) & ~63
    sc = (scratch.ctypes.data + 63) & ~63
    l.ref_init(C.c_void_p(pp), C.byref(cp))
    if cfg["kbps"]:
        rp = refenc.RunParam(desired_frame_bytes=cfg["kbps"] * 1000 // 8 // 30, qp_min=10, qp_max=50)
    else:
        rp = refenc.RunParam(qp_min=cfg["qp"], qp_max=cfg["qp"])
    conn.send(("ready", variant))
    t = 0
    digests = []
    while True:
        cmd = conn.recv()
        if cmd == "quit":
            break
        if cmd == "digests":
            conn.send(digests)
            continue
        if isinstance(cmd, tuple):          # ("run", k): k frames back to back, free-running (no lock step with the others)
            t0 = time.perf_counter()
            err_any = 0
            for _ in range(cmd[1]):
                f = clip[t % nframes].copy()
                yuv = refenc.IoYuv()
                base = f.ctypes.data
                yuv.yuv[0], yuv.yuv[1], yuv.yuv[2] = base, base + W * H, base + W * H * 5 // 4
                yuv.stride[0], yuv.stride[1], yuv.stride[2] = W, W // 2, W // 2
                data, n = C.c_void_p(), C.c_int()
                err_any |= l.ref_encode(C.c_void_p(pp), C.c_void_p(sc), C.byref(rp), C.byref(yuv), C.byref(data), C.byref(n))
                digests.append(hashlib.md5(C.string_at(data.value, n.value)).hexdigest() if not err_any else None)
                t += 1
            conn.send((err_any, time.perf_counter() - t0))
            continue
        f = clip[t % nframes].copy()
        yuv = refenc.IoYuv()
        base = f.ctypes.data
        yuv.yuv[0], yuv.yuv[1], yuv.yuv[2] = base, base + W * H, base + W * H * 5 // 4
        yuv.stride[0], yuv.stride[1], yuv.stride[2] = W, W // 2, W // 2
        data, n = C.c_void_p(), C.c_int()
        err = l.ref_encode(C.c_void_p(pp), C.c_void_p(sc), C.byref(rp), C.byref(yuv), C.byref(data), C.byref(n))
        conn.send((err, n.value))
        # after the reply: the parent's clock does not see the digest of the frame just coded
        digests.append(hashlib.md5(C.string_at(data.value, n.value)).hexdigest() if not err else None)
        t += 1


def _core_order():
    """logical CPUs this process may use, one hardware thread of every physical core first, their siblings after"""
    allowed = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else []
    first, rest, seen = [], [], set()
    for c in allowed:
        try:
            sib = open("/sys/devices/system/cpu/cpu%d/topology/thread_siblings_list" % c).read().strip()
        except Exception:
            sib = str(c)
        if sib in seen:
            rest.append(c)
        else:
            seen.add(sib)
            first.append(c)
    return first + rest


class RefPool:
    def __init__(self, cfg, units, nframes):
        ctx = mp.get_context("spawn")
        self.conns, self.procs = [], []
        cores = _core_order()
        for k, u in enumerate(units):
            a, b = ctx.Pipe()
            core = cores[k] if len(units) <= len(cores) else None     # one core each when they fit, else the OS schedules
            p = ctx.Process(target=_ref_worker, args=(cfg, u, nframes, b, core), daemon=True)
            p.start()
            self.conns.append(a)
            self.procs.append(p)
        self.variant = ""
        for c in self.conns:
            msg = c.recv()
            assert msg[0] == "ready"
            self.variant = msg[1]

    def step(self):
        for c in self.conns:
            c.send("go")
        out = [c.recv() for c in self.conns]
        assert all(e == 0 for e, _ in out)
        return sum(n for _, n in out)

    def run(self, k):
        """every process encodes its next k frames at its own pace; returns the wall time until the last one is done"""
        t0 = time.perf_counter()
        for c in self.conns:
            c.send(("run", k))
        out = [c.recv() for c in self.conns]
        dt = time.perf_counter() - t0
        assert all(e == 0 for e, _ in out)
        return dt

    def digests(self):
        for c in self.conns:
            c.send("digests")
        return [c.recv() for c in self.conns]

    def close(self):
        for c in self.conns:
            c.send("quit")
        for p in self.procs:
            p.join(timeout=5)


def run_reference(cfg, units, steps, warmup, want_digests=False):
    """steps x len(units) frames of the workload on the host cores (one process per unit)."""
    nframes = min(max(cfg["gop"], 1) if cfg["gop"] > 1 else 8, steps + warmup)
    pool = RefPool(cfg, units, nframes)
    if warmup:
        pool.run(warmup)
    dt = pool.run(steps)
    dig = pool.digests() if want_digests else None
    variant = pool.variant
    pool.close()
    return len(units) * steps / dt, dt, dig, variant


def clip_frames(cfg, steps, warmup):
    """frames held per unit: a whole GOP (the timed region wraps around into the next GOP's IDR), 8 for all-intra"""
    return min(cfg["gop"] if cfg["gop"] > 1 else 8, steps + warmup)


# --------------------------------------------------------------------------------------
# roofline inputs measured elsewhere and committed under profiles/
# --------------------------------------------------------------------------------------
def l2_policy(nseg, W, H, frame_bytes):
    """What a step touches: every step reads input frames that were never read before (the clip is resident, frame t of
    every unit) and re-writes the per-step state of every unit -- two padded pictures, three half-sample planes,
    SAD-map / motion-estimation records (2336 B per macroblock), levels, bit strings, macroblock records."""
    w16, h16 = (W + 15) // 16 * 16, (H + 15) // 16 * 16
    pad = (w16 + 32) * (h16 + 32)
    nmb = (w16 // 16) * (h16 // 16)
    per_unit = frame_bytes + 2 * pad * 3 // 2 + 3 * pad + nmb * (2336 + 832 + 2048 + 144 + 64 + 64)
    ws = nseg * per_unit
    l2 = 126 * 2**20
    if ws > l2:
        return ("no flush needed: %d MB of fresh input per step (never re-read) and a per-step working set of %d MB "
                "(pictures, half-sample planes, SAD-map records, levels, bit strings) > 126 MB of L2" % (nseg * frame_bytes // 2**20, ws // 2**20))
    return ("per-step working set %d MB < 126 MB of L2 and no flush: %d MB of fresh input per step (never re-read); what stays "
            "in L2 between steps is the previous reconstruction, as in the encoder's real operation" % (ws // 2**20, nseg * frame_bytes // 2**20))


def load_json(path, default=None):
    try:
        return json.load(open(path))
    except Exception:
        return default


def int_peak(live=True):
    """Measured packed-integer issue peak of this chip (tools/ubench/intpeak): live when the binary travelled, else the
    committed measurement."""
    exe = os.path.join(ROOT, "tools", "ubench", "intpeak")
    rows = None
    source = None
    if live and os.path.exists(exe):
        try:
            out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=60).stdout
            rows = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
            source = "measured live (tools/ubench/intpeak)"
        except Exception:
            rows = None
    if not rows:
        rows = load_json(os.path.join(ROOT, "profiles", "r02_intpeak.json"), [])
        source = "profiles/r02_intpeak.json" if rows else None
    return {r["op"]: r for r in rows}, source


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--segments", type=int, default=0, help="override the number of units per GPU (developer)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the reference leg (then parity is NOT checked)")
    args = ap.parse_args()
    # stdout carries the ONE JSON line and nothing else: whatever libraries print (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = dict(CONFIGS[args.config])
    if args.segments:
        cfg["units"] = args.segments
    W, H, GOP = cfg["w"], cfg["h"], cfg["gop"]
    FRAME_BYTES = W * H * 3 // 2
    K, Wm = args.steps, max(args.warmup, 0)
    ngpu = max(world, args.gpus, 1)
    if args.scaling == "weak":
        units_of = lambda r: list(range(r * cfg["units"], (r + 1) * cfg["units"]))       # noqa: E731
        total_units = cfg["units"] * ngpu
    else:
        units_of = lambda r: [k for k in range(cfg["units"]) if k % ngpu == r]           # noqa: E731
        total_units = cfg["units"]
    my_units = units_of(rank)
    nseg = len(my_units)
    cores = os.cpu_count() or 1
    metric = "1080p encode fps, bit-exact to ref" if args.config in ("c2", "c4") else "%dx%d encode fps, bit-exact to ref" % (W, H)
    config = {"workload": "%s: %s; %d unit(s) per GPU (%s scaling), step = frame t of every unit" % (args.config, cfg["desc"], nseg, args.scaling),
              "config": args.config, "width": W, "height": H, "gop": GOP, "qp": cfg["qp"], "kbps": cfg["kbps"],
              "units_per_gpu": nseg if args.scaling == "weak" else None, "units_total": total_units, "speed": 0,
              "timed_frames_per_unit": "t = %d..%d of an endless IPPP stream with GOP %d (IDR whenever t %% GOP == 0)" % (Wm, Wm + K - 1, GOP),
              "l2_policy": l2_policy(nseg, W, H, FRAME_BYTES)}

    if args.impl == "reference":
        if rank != 0:
            return
        units = list(range(total_units))
        fps, dt, _, variant = run_reference(cfg, units, K, Wm)
        line = {"impl": "reference", "metric": metric, "value": fps, "unit": "frames/s",
                "n_gpus": args.gpus, "steps": K, "warmup": Wm, "ms_per_step": dt / K * 1e3, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                "megapixels_per_s": fps * W * H / 1e6,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": min(cores, len(units)), "kind": "reference",
                                 "sample": "%d steps x %d units, one process per unit on %d host cores, build %s"
                                           % (K, len(units), cores, "libh264ref%s.so" % variant)},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B = load_binding()
    lib = B.Library()
    nframes = clip_frames(cfg, K, Wm)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # clips in pinned host memory (one per unit of this rank)
    clips = []
    for u in my_units:
        pinned = torch.empty((nframes, FRAME_BYTES), dtype=torch.uint8, pin_memory=True)
        pinned.numpy()[:] = make_clip(cfg, u, nframes)
        clips.append(pinned)
    NT = 8
    lib.lib.h264b200_last_timing_ex.argtypes = [C.POINTER(C.c_float), C.c_int]
    lib.lib.h264b200_last_timing_ex.restype = C.c_int

    def run(resident):
        n = nseg
        encs = [B.Encoder(lib, W, H, GOP) for _ in range(n)]
        rps = [e.run_param(qp=cfg["qp"], kbps=cfg["kbps"]) for e in encs]
        if resident:
            for e, c in zip(encs, clips):
                err = lib.lib.H264E_preload(C.c_void_p(e.persist), nframes, C.c_void_p(c.data_ptr()))
                assert err == 0
        yuvs = [B.IoYuv() for _ in range(n)]
        nxt = [B.IoYuv() for _ in range(n)]
        P = (C.c_void_p * n)(*[e.persist for e in encs])
        S = (C.c_void_p * n)(*[e.scratch for e in encs])
        R = (C.c_void_p * n)(*[C.addressof(r) for r in rps])
        Y = (C.c_void_p * n)(*[C.addressof(y) for y in yuvs])
        D = (C.c_void_p * n)()
        N = (C.c_int * n)()
        tm = (C.c_float * NT)()
        kern = [0.0] * NT
        out_bytes = 0
        # every coded frame is kept (one host copy per unit and step, the consumer's read of the result) for the parity check
        keep = [[] for _ in range(n)]

        def step(t):
            nonlocal out_bytes
            for i in range(n):
                if resident:
                    yuvs[i].yuv[0] = None
                    yuvs[i].stride[0] = t % nframes
                else:
                    base = clips[i].data_ptr() + (t % nframes) * FRAME_BYTES
                    yuvs[i].yuv[0], yuvs[i].yuv[1], yuvs[i].yuv[2] = base, base + W * H, base + W * H * 5 // 4
                    yuvs[i].stride[0], yuvs[i].stride[1], yuvs[i].stride[2] = W, W // 2, W // 2
            if not resident and t + 1 < Wm + K:
                # double buffering through the public API: the copy of frame t + 1 overlaps the encoding of frame t
                for i in range(n):
                    base = clips[i].data_ptr() + ((t + 1) % nframes) * FRAME_BYTES
                    nxt[i].yuv[0], nxt[i].yuv[1], nxt[i].yuv[2] = base, base + W * H, base + W * H * 5 // 4
                    nxt[i].stride[0], nxt[i].stride[1], nxt[i].stride[2] = W, W // 2, W // 2
                    err = lib.lib.H264E_prefetch(C.c_void_p(encs[i].persist), C.byref(nxt[i]))
                    assert err == 0, "H264E_prefetch error %d" % err
            err = lib.lib.H264E_encode_batch(n, P, S, R, Y, D, N)
            assert err == 0, "H264E_encode_batch error %d" % err
            for i in range(n):
                keep[i].append(C.string_at(D[i], N[i]))
                out_bytes += N[i]

        for t in range(Wm):
            step(t)
        out_bytes = 0
        launches0 = lib.launch_count()
        barrier()
        t0 = time.perf_counter()
        for t in range(Wm, Wm + K):
            step(t)
            lib.lib.h264b200_last_timing_ex(tm, NT)
            for i in range(NT):
                kern[i] += tm[i]
        barrier()
        dt = time.perf_counter() - t0
        launches = lib.launch_count() - launches0
        for e in encs:
            e.close()
        return dt, kern, launches, out_bytes, keep

    sampler = ClockSampler(local_rank)
    sampler.start()
    dt_res, kern, launches, _, keep_res = run(resident=True)
    clocks = sampler.stop()
    lib.lib.h264b200_prefetch_hits.restype = C.c_long
    hits0 = lib.lib.h264b200_prefetch_hits()
    dt_e2e, kern_e2e, _, out_bytes, keep_e2e = run(resident=False)
    prefetch_hits = int(lib.lib.h264b200_prefetch_hits() - hits0)
    dt_res = max_over_ranks(dt_res)
    dt_e2e = max_over_ranks(dt_e2e)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    total_frames = total_units * K
    peaks = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"), {})
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # algorithmic HBM bytes of a frame: input + padded reference + recon + border + bitstream
    w16, h16 = (W + 15) & ~15, (H + 15) & ~15
    b_in, b_ref = 1.5 * W * H, 1.5 * (w16 + 32) * (h16 + 32)
    b_rec = 1.5 * w16 * h16 + (b_ref - 1.5 * w16 * h16)
    p_frames = sum(1 for t in range(Wm, Wm + K) if GOP != 1 and (GOP == 0 or t % GOP))
    i_frames = K - p_frames
    alg_bytes_step = nseg * ((b_in + b_ref + b_rec) * p_frames + (b_in + b_rec) * i_frames) / K + out_bytes / K
    nmb = (w16 // 16) * (h16 // 16)
    k_main_ms = kern[1] / K
    achieved_hbm = alg_bytes_step / (k_main_ms * 1e-3) / 1e9 if k_main_ms > 0 else 0.0
    # algorithmic integer work (SURVEY 8(d) counting rules), measured on this workload with the reference's own counters
    opc = load_json(os.path.join(ROOT, "profiles", "r02_opcount.json"), {}).get(args.config)
    peak_rows, peak_source = int_peak(live=not os.environ.get("H264B200_NO_LIVE_PEAK"))
    roof = {"bound": "issue (integer ALU) / wavefront latency -- not hbm, not tensor", "kernel": "macroblock sweep (motion search + mode decision + transform/quant/recon kernels of the step)",
            "achieved": None, "peak": None, "unit": "G int-ops/s", "frac": None, "traffic": None}
    if opc and peak_rows.get("vsadu4") and k_main_ms > 0:
        ops_p, ops_i = float(opc["ops_per_mb_p"]), float(opc["ops_per_mb_i"])
        ops_step = nseg * nmb * (ops_p * p_frames + ops_i * i_frames) / K
        ach = ops_step / (k_main_ms * 1e-3) / 1e9
        # peak in the same unit: one packed-byte SAD instruction retires 4 sample-operations per lane; the measured
        # VABSDIFF4 issue rate x 4 is the most the SM can do for the dominant operation class of this path
        peak = float(peak_rows["vsadu4"]["gsamples"])
        roof.update({"achieved": ach, "peak": peak, "frac": ach / peak,
                     "algorithmic_ops_per_mb": {"p": ops_p, "i": ops_i, "source": "profiles/r02_opcount.json (tools/opcount.py, oracle/_ref/libh264ref_count.so)"},
                     "peak_source": "%s: packed-byte SAD (VABSDIFF4) issue rate x 4 samples; 32-bit IMAD rate %.0f G/s" % (peak_source, float(peak_rows.get("imad", {}).get("gops", 0)))})
    tr = load_json(os.path.join(ROOT, "profiles", "r02_traffic.json"), {}).get(args.config)
    roof["traffic"] = tr.get("bytes_per_launch") if tr else None
    roof["traffic_source"] = tr.get("source") if tr else None
    roof["hbm"] = {"achieved": achieved_hbm, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_hbm / hbm_peak,
                   "algorithmic_bytes_per_launch": alg_bytes_step, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback"}
    steps_wave = (w16 // 16) + 2 * ((h16 // 16) - 1)
    roof["us_per_wavefront_step"] = k_main_ms * 1e3 / steps_wave if k_main_ms else None
    names = ["device_total", "macroblock_sweep", "deblock+borders+hpel", "cavlc+scan+pack (second stream)", "sad_maps", "me_prepass", "intra_check", "extra7"]
    line = {
        "metric": metric, "value": total_frames / dt_res, "unit": "frames/s",
        "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": dt_res / K * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
        "megapixels_per_s": total_frames / dt_res * W * H / 1e6,
        "e2e": {"value": total_frames / dt_e2e, "unit": "frames/s", "h2d_bytes_per_step": nseg * FRAME_BYTES,
                "d2h_bytes_per_step": int(out_bytes / K), "ms_per_step": dt_e2e / K * 1e3,
                "device_ms_per_step": kern_e2e[0] / K, "macroblock_sweep_ms_per_step": kern_e2e[1] / K},
        "gpu_launches": launches, "e2e_prefetch_hits": prefetch_hits,
        "kernel_ms_per_step": {names[i]: kern[i] / K for i in range(NT) if kern[i] or i < 4},
        "roofline": roof,
        "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        try:
            # the reference encodes exactly the frames both runs encoded (warm-up included): its time over the TIMED frames
            # is the CPU baseline, its bytes are the parity oracle
            fps, dt, dig, variant = run_reference(cfg, my_units, K, Wm, want_digests=True)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": min(cores, nseg), "kind": "reference",
                                    "sample": "the %d timed frames of each of the %d units of rank 0, one pinned process per unit on %d host cores, build libh264ref%s.so"
                                              % (K, nseg, cores, variant)}
            checked, bad = 0, []
            for i in range(nseg):
                for t in range(Wm + K):
                    for tag, keep in (("resident", keep_res), ("e2e", keep_e2e)):
                        checked += 1
                        if hashlib.md5(keep[i][t]).hexdigest() != dig[i][t]:
                            bad.append((tag, my_units[i], t))
            line["parity"] = {"checked_frames": checked, "mismatches": len(bad), "first_mismatches": bad[:5],
                              "scope": "every coded frame (warm-up + timed, resident run and e2e run) of rank 0's %d units vs the unmodified reference" % nseg}
            if bad:
                line["value"] = None
                line["e2e"]["value"] = None
                line["invalid"] = "output differs from the reference"
        except Exception as ex:      # the oracle always exists; report loudly if it does not run
            line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": "failed: %r" % (ex,)}
            line["parity"] = {"checked_frames": 0, "mismatches": None, "error": repr(ex)}
    json_out.write(json.dumps(line) + "\n")
    json_out.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
