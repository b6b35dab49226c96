#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 H.264 macroblock-encode path.

Workload (BASELINE.json configs[1]): synthetic 1080p YUV420, 600 frames, IPPP GOP=60,
fixed QP 28, on ONE B200 = 10 closed-GOP segments of 60 frames, one fresh encoder
session per segment (the reference run once per segment is the golden output).  A
"step" is one pass of the hot path over one batch: frame t of every segment
(n_segments frames) in one device submission.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1: launched by torchrun, one rank per GPU; every rank encodes its own n_segments
segments (weak scaling, no data-path collective; torch.distributed is used only for the
barrier and the max-over-ranks time).

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM;
`e2e` = frames/s through the public C API with pinned host buffers (H2D of every input,
D2H of every payload inside the timed region; the copy of frame t + 1 is started with
H264E_prefetch before frame t is encoded, so it overlaps the kernels of frame t).  `roofline` is for the dominant kernel
(k_encode_rows): algorithmic HBM bytes / measured kernel time against the measured HBM
peak -- the path is latency/ALU bound, not HBM bound, and the fraction says so.
`cpu_baseline` = the unmodified reference (oracle/_ref) on the host cores, one process per
segment.  `--impl reference` times that reference arm alone.
"""
import argparse
import ctypes as C
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

W, H, GOP, QP, NSEG = 1920, 1080, 60, 28, 10
FRAME_BYTES = W * H * 3 // 2


def load_binding():
    spec = importlib.util.spec_from_file_location("h264lab_binding", os.path.join(ROOT, "h264-lab_b200", "binding.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_clip(seg, nframes):
    """Deterministic synthetic clip of one segment (tests/content.py, seed per segment)."""
    import content
    return content.panning(W, H, nframes, seed=1000 + seg)


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
# reference arm (CPU): one process per segment, persistent encoder, stepped by the parent
# --------------------------------------------------------------------------------------
def _ref_worker(seg, nframes, conn):
    import numpy as np
    import refenc
    l = refenc.lib("_fast") if refenc.have_ref("_fast") else refenc.lib()
    clip = make_clip(seg, nframes)
    cp = refenc.CreateParam(width=W, height=H, gop=GOP, const_input_flag=1, vbv_size_bytes=100000 // 8, enableNEON=1, num_layers=1)
    sp, ss = C.c_int(), C.c_int()
    l.ref_sizeof(C.byref(cp), C.byref(sp), C.byref(ss))
    persist = np.zeros(sp.value + 64, np.uint8)
    scratch = np.zeros(ss.value + 64, np.uint8)
    pp = (persist.ctypes.data + 63) & ~63
    sc = (scratch.ctypes.data + 63) & ~63
    l.ref_init(C.c_void_p(pp), C.byref(cp))
    rp = refenc.RunParam(qp_min=QP, qp_max=QP)
    conn.send("ready")
    t = 0
    while True:
        cmd = conn.recv()
        if cmd == "quit":
            break
        f = clip[t % nframes].copy()
        yuv = refenc.IoYuv()
        base = f.ctypes.data
        yuv.yuv[0], yuv.yuv[1], yuv.yuv[2] = base, base + W * H, base + W * H * 5 // 4
        yuv.stride[0], yuv.stride[1], yuv.stride[2] = W, W // 2, W // 2
        data, n = C.c_void_p(), C.c_int()
        err = l.ref_encode(C.c_void_p(pp), C.c_void_p(sc), C.byref(rp), C.byref(yuv), C.byref(data), C.byref(n))
        conn.send((err, n.value))
        t += 1


class RefPool:
    def __init__(self, nseg, nframes):
        ctx = mp.get_context("spawn")
        self.conns, self.procs = [], []
        for s in range(nseg):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_ref_worker, args=(s, nframes, b), daemon=True)
            p.start()
            self.conns.append(a)
            self.procs.append(p)
        for c in self.conns:
            assert c.recv() == "ready"

    def step(self):
        for c in self.conns:
            c.send("go")
        out = [c.recv() for c in self.conns]
        assert all(e == 0 for e, _ in out)
        return sum(n for _, n in out)

    def close(self):
        for c in self.conns:
            c.send("quit")
        for p in self.procs:
            p.join(timeout=5)


def run_reference(nseg, steps, warmup):
    """steps x nseg frames of the workload on the host cores (all cores: one process per segment)."""
    nframes = min(GOP, steps + warmup)
    pool = RefPool(nseg, nframes)
    for _ in range(warmup):
        pool.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        pool.step()
    dt = time.perf_counter() - t0
    pool.close()
    return nseg * steps / dt, dt


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--segments", type=int, default=NSEG)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # stdout carries the ONE JSON line and nothing else: whatever libraries print (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    nseg, K, Wm = args.segments, args.steps, max(args.warmup, 0)
    cores = os.cpu_count() or 1
    config = {"workload": "synthetic 1080p YUV420, IPPP GOP=60, fixed QP 28, %d closed-GOP segments per GPU, "
                          "step = frame t of every segment (%d frames)" % (nseg, nseg),
              "width": W, "height": H, "gop": GOP, "qp": QP, "segments_per_gpu": nseg, "speed": 0,
              "l2_policy": "inputs larger than L2: %d MB of fresh input + %d MB of reference pictures per step"
                           % (nseg * FRAME_BYTES // 2**20, nseg * FRAME_BYTES // 2**20)}

    if args.impl == "reference":
        if rank != 0:
            return
        # weak scaling: at N GPUs the job is N x nseg segments; the reference gets all of them on the host cores
        nref = nseg * max(args.gpus, 1)
        fps, dt = run_reference(nref, K, Wm)
        config = dict(config, reference_segments=nref)
        line = {"impl": "reference", "metric": "1080p encode fps, bit-exact to ref", "value": fps, "unit": "frames/s",
                "n_gpus": args.gpus, "steps": K, "warmup": Wm, "ms_per_step": dt / K * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": min(cores, nref), "kind": "reference",
                                 "sample": "%d steps x %d segments, one process per segment on %d host cores" % (K, nref, cores)},
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
    nframes = min(GOP, Wm + K)

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

    # clips in pinned host memory (one per segment of this rank)
    clips = []
    for s in range(nseg):
        pinned = torch.empty((nframes, FRAME_BYTES), dtype=torch.uint8, pin_memory=True)
        pinned.numpy()[:] = make_clip(rank * nseg + s, nframes)
        clips.append(pinned)

    def run(resident):
        encs = [B.Encoder(lib, W, H, GOP) for _ in range(nseg)]
        rps = [e.run_param(qp=QP) for e in encs]
        if resident:
            for e, c in zip(encs, clips):
                err = lib.lib.H264E_preload(C.c_void_p(e.persist), nframes, C.c_void_p(c.data_ptr()))
                assert err == 0
        n = nseg
        yuvs = [B.IoYuv() for _ in range(n)]
        nxt = [B.IoYuv() for _ in range(n)]
        P = (C.c_void_p * n)(*[e.persist for e in encs])
        S = (C.c_void_p * n)(*[e.scratch for e in encs])
        R = (C.c_void_p * n)(*[C.addressof(r) for r in rps])
        Y = (C.c_void_p * n)(*[C.addressof(y) for y in yuvs])
        D = (C.c_void_p * n)()
        N = (C.c_int * n)()
        tm = (C.c_float * 4)()
        kern = [0.0, 0.0, 0.0, 0.0]
        out_bytes = 0

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
            out_bytes += sum(N[i] for i in range(n))

        for t in range(Wm):
            step(t)
        out_bytes = 0
        launches0 = lib.launch_count()
        barrier()
        t0 = time.perf_counter()
        for t in range(Wm, Wm + K):
            step(t)
            lib.lib.h264b200_last_timing(tm)
            for i in range(4):
                kern[i] += tm[i]
        barrier()
        dt = time.perf_counter() - t0
        launches = lib.launch_count() - launches0
        for e in encs:
            e.close()
        return dt, kern, launches, out_bytes

    sampler = ClockSampler(local_rank)
    sampler.start()
    dt_res, kern, launches, _ = run(resident=True)
    clocks = sampler.stop()
    lib.lib.h264b200_prefetch_hits.restype = C.c_long
    hits0 = lib.lib.h264b200_prefetch_hits()
    dt_e2e, kern_e2e, _, out_bytes = run(resident=False)
    prefetch_hits = int(lib.lib.h264b200_prefetch_hits() - hits0)
    dt_res = max_over_ranks(dt_res)
    dt_e2e = max_over_ranks(dt_e2e)
    if rank != 0:
        return
    total_frames = nseg * K * world
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # algorithmic HBM bytes of a 1080p frame: input + padded reference + recon + border + bitstream
    w16, h16 = 1920, 1088
    b_in, b_ref = 1.5 * W * H, 1.5 * (w16 + 32) * (h16 + 32)
    b_rec = 1.5 * w16 * h16 + (b_ref - 1.5 * w16 * h16)
    p_frames = sum(1 for t in range(Wm, Wm + K) if t % GOP)
    i_frames = K - p_frames
    alg_bytes_step = nseg * ((b_in + b_ref + b_rec) * p_frames + (b_in + b_rec) * i_frames) / K + out_bytes / K
    # DRAM traffic of the dominant kernel, per launch, from the committed ncu --set full capture of this workload
    traffic = None
    inst = None
    try:
        km = json.load(open(os.path.join(ROOT, "profiles", "r01c_k_encode_rows_10stream_keymetrics.json")))
        unit = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
        traffic = sum(float(km[k][0]) * unit[km[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum")) if nseg == NSEG else None
        # integer-ALU view of the same launch (north_star: ME / transform kernels against the SM issue peak):
        # warp instructions of the captured launch / the live launch duration, against 4 issue slots x 148 SMs x SM clock
        inst = float(km["smsp__inst_executed.sum"][0]) if nseg == NSEG and "smsp__inst_executed.sum" in km else None
    except Exception:
        traffic = None
        inst = None
    k_enc_ms = kern[1] / K
    achieved = alg_bytes_step / (k_enc_ms * 1e-3) / 1e9 if k_enc_ms > 0 else 0.0
    alu = None
    if inst and k_enc_ms > 0:
        sm_mhz = float(clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0)
        peak_ginst = 148 * 4 * sm_mhz * 1e6 / 1e9
        ach_ginst = inst / (k_enc_ms * 1e-3) / 1e9
        alu = {"achieved": ach_ginst, "peak": peak_ginst, "unit": "G warp-instructions/s", "frac": ach_ginst / peak_ginst,
               "source": "smsp__inst_executed.sum of the committed 10-frame capture / live launch time; peak = 148 SMs x 4 issue slots x SM clock"}
    line = {
        "metric": "1080p encode fps, bit-exact to ref", "value": total_frames / dt_res, "unit": "frames/s",
        "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": dt_res / K * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
        "megapixels_per_s": total_frames / dt_res * W * H / 1e6,
        "e2e": {"value": total_frames / dt_e2e, "unit": "frames/s", "h2d_bytes_per_step": nseg * FRAME_BYTES,
                "d2h_bytes_per_step": int(out_bytes / K), "ms_per_step": dt_e2e / K * 1e3,
                "device_ms_per_step": kern_e2e[0] / K, "k_encode_rows_ms_per_step": kern_e2e[1] / K},
        "gpu_launches": launches, "e2e_prefetch_hits": prefetch_hits,
        "kernel_ms_per_step": {"device_total": kern[0] / K, "k_encode_rows": kern[1] / K, "k_deblock_rows+k_borders": kern[2] / K,
                               "k_cavlc+k_scan+k_pack": kern[3] / K},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic, "traffic_unit": "bytes per k_encode_rows launch (ncu dram__bytes_read+write, profiles/r01c_k_encode_rows_10stream_keymetrics.json)",
                     "algorithmic_bytes_per_launch": alg_bytes_step, "kernel": "k_encode_rows (sweep 0 + repair waves of the step)", "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                     "alu": alu,
                     "note": "integer/latency-bound wavefront: HBM is not the limiter (SURVEY 8(d)); "
                             "us per wavefront step = %.2f" % (k_enc_ms * 1e3 / (120 + 2 * 67) if k_enc_ms else 0)},
        "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        try:
            cpu_steps = 4
            fps, dt = run_reference(nseg, cpu_steps, 1)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": min(cores, nseg), "kind": "reference",
                                    "sample": "frames 1..%d of each of the %d segments (P frames), one process per segment on %d host cores"
                                              % (cpu_steps, nseg, cores)}
        except Exception as ex:      # the oracle always exists; report loudly if it does not run
            line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": "failed: %r" % (ex,)}
    json_out.write(json.dumps(line) + "\n")
    json_out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
