"""GPU suite (pytest -m gpu): the product library (host C + sm_100a kernels), called
through its C ABI (H264E_sizeof / H264E_init / H264E_encode / H264E_encode_batch), must
reproduce the compiled reference (oracle/_ref) byte for byte -- bit stream, per-frame
sizes, reconstruction planes -- and the committed golden digests (tests/golden)."""
import hashlib
import json
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "manifest.json")


def _check(case, binding, cuda_lib, ref):
    name, kind, w, h, n, gop, kw = case
    frames = cases.make(kind, w, h, n)
    rbs, rsizes, rrec, _ = ref.encode_sequence(frames, w, h, gop, **kw)
    bs, sizes, rec = binding.encode_sequence(cuda_lib, frames, w, h, gop, **kw)
    assert cuda_lib.backend() == "cuda-sm_100a"
    assert list(sizes) == list(rsizes), "per-frame sizes differ"
    assert bs == rbs, "bit stream differs"
    assert np.array_equal(rec, rrec), "reconstruction differs"
    return bs, rec


@pytest.mark.parametrize("case", cases.SMALL, ids=lambda c: c[0])
def test_small_cases(case, binding, cuda_lib, ref):
    _check(case, binding, cuda_lib, ref)


@pytest.mark.parametrize("case", cases.CIF_FOREMAN_SUBSTITUTE, ids=lambda c: c[0])
def test_cif_reference_script_qps(case, binding, cuda_lib, ref):
    _check(case, binding, cuda_lib, ref)


@pytest.mark.parametrize("case", cases.LARGE, ids=lambda c: c[0])
def test_large_cases(case, binding, cuda_lib, ref):
    _check(case, binding, cuda_lib, ref)


def test_golden_digests(binding, cuda_lib):
    """Committed digests of reference output (made by tests/golden/make_golden.py where
    /root/reference exists): independent of oracle/_ref being present on the box."""
    man = json.load(open(GOLDEN))
    for g in man["cases"]:
        frames = cases.make(g["kind"], g["width"], g["height"], g["frames"])
        assert hashlib.md5(frames.tobytes()).hexdigest() == g["input_md5"]
        bs, sizes, rec = binding.encode_sequence(cuda_lib, frames, g["width"], g["height"], g["gop"], **g["kw"])
        assert hashlib.md5(bs).hexdigest() == g["bitstream_md5"], g["name"]
        assert hashlib.md5(rec.tobytes()).hexdigest() == g["recon_md5"], g["name"]


def test_batch_equals_sequential(binding, cuda_lib, ref):
    """Closed-GOP segments encoded concurrently in one submission == each segment on its own
    (and == the reference run once per segment)."""
    w, h, seglen, nseg = 352, 288, 5, 4
    frames = cases.make("panning", w, h, seglen * nseg)
    encs = [binding.Encoder(cuda_lib, w, h, seglen) for _ in range(nseg)]
    rps = [e.run_param(qp=28) for e in encs]
    outs = [b"" for _ in range(nseg)]
    for t in range(seglen):
        fr = [frames[s * seglen + t].copy() for s in range(nseg)]
        res = binding.encode_batch(cuda_lib, encs, fr, rps)
        for s in range(nseg):
            outs[s] += res[s]
    for s in range(nseg):
        rbs, _, rrec, _ = ref.encode_sequence(frames[s * seglen:(s + 1) * seglen], w, h, seglen, qp=28)
        assert outs[s] == rbs, "segment %d" % s
        assert np.array_equal(encs[s].recon(), rrec[-1])
    for e in encs:
        e.close()


def test_round_trip_decodes(binding, cuda_lib, tmp_path):
    """Size-independent property at a BASELINE size: the 1080p stream parses in a real
    decoder (FFmpeg through cv2) into the right number of frames of the cropped size."""
    cv2 = pytest.importorskip("cv2")
    w, h, n = 1920, 1080, 3
    frames = cases.make("panning", w, h, n)
    bs, sizes, rec = binding.encode_sequence(cuda_lib, frames, w, h, 3, qp=28)
    p = tmp_path / "o.264"
    p.write_bytes(bs)
    cap = cv2.VideoCapture(str(p), cv2.CAP_FFMPEG)
    cnt = 0
    while True:
        ok, img = cap.read()
        if not ok:
            break
        assert img.shape[0] == h and img.shape[1] == w
        cnt += 1
    assert cnt == n


def test_in_place_recon_and_errors(binding, cuda_lib, ref):
    """const_input_flag = 0 overwrites the caller's planes with the reconstruction (H:6719-6723)."""
    w, h = 352, 288
    frames = cases.make("panning", w, h, 2)
    enc = binding.Encoder(cuda_lib, w, h, 2, const_input=0)
    rp = enc.run_param(qp=30)
    _, _, rrec, _ = ref.encode_sequence(frames, w, h, 2, qp=30)
    for i in range(2):
        f = frames[i].copy()
        enc.encode(f, rp)
        assert np.array_equal(f, rrec[i])
    enc.close()


def test_in_place_transparent_frames(binding, cuda_lib, ref):
    """VBV overflow with vbv_overflow_empty_frame_flag: transparent frames (one skip run, H:6497-6508); in in-place mode
    the caller's planes receive the unchanged reference picture.  Also through H264E_encode_batch."""
    w, h, n = 352, 288, 6
    frames = cases.make("noise", w, h, n)
    rbs, rsizes, rrec, _ = ref.encode_sequence(frames, w, h, n, kbps=2500, empty_frames=1)
    assert list(rsizes[1:4]) == [10, 10, 10]          # the case does produce transparent frames
    enc = binding.Encoder(cuda_lib, w, h, n, const_input=0, vbv_overflow_empty_frame_flag=1)
    rp = enc.run_param(kbps=2500)
    out = b""
    for i in range(n):
        f = frames[i].copy()
        out += enc.encode(f, rp)
        assert np.array_equal(f, rrec[i]), i
    enc.close()
    assert out == rbs
    encs = [binding.Encoder(cuda_lib, w, h, n, vbv_overflow_empty_frame_flag=1) for _ in range(2)]
    rps = [e.run_param(kbps=2500) for e in encs]
    outs = [b"", b""]
    for i in range(n):
        res = binding.encode_batch(cuda_lib, encs, [frames[i].copy(), frames[i].copy()], rps)
        outs = [o + r for o, r in zip(outs, res)]
    assert outs[0] == rbs and outs[1] == rbs
    for e in encs:
        e.close()


def test_many_independent_streams_in_one_batch(binding, cuda_lib, ref):
    """BASELINE config 5 in small: independent streams (distinct content) batched in one device
    submission per frame; every stream equals its own reference run."""
    w, h, n, nstreams = 320, 240, 4, 8
    clips = [cases.make("panning", w, h, n, seed=100 + s) for s in range(nstreams)]
    encs = [binding.Encoder(cuda_lib, w, h, 60) for _ in range(nstreams)]
    rps = [e.run_param(qp=28) for e in encs]
    outs = [b"" for _ in range(nstreams)]
    for t in range(n):
        res = binding.encode_batch(cuda_lib, encs, [c[t].copy() for c in clips], rps)
        for s in range(nstreams):
            outs[s] += res[s]
    for s in range(nstreams):
        rbs, _, rrec, _ = ref.encode_sequence(clips[s], w, h, 60, qp=28)
        assert outs[s] == rbs, "stream %d" % s
        assert np.array_equal(encs[s].recon(), rrec[-1])
    for e in encs:
        e.close()


def test_rate_controlled_gop_shards(binding, cuda_lib, ref):
    """BASELINE config 3 in small: rate-controlled closed-GOP segments encoded concurrently, one fresh
    session per segment (the QP trajectory of every segment must match the reference's)."""
    w, h, seglen, nseg = 352, 288, 6, 3
    frames = cases.make("multi", w, h, seglen * nseg)
    encs = [binding.Encoder(cuda_lib, w, h, seglen) for _ in range(nseg)]
    rps = [e.run_param(kbps=600) for e in encs]
    outs = [b"" for _ in range(nseg)]
    for t in range(seglen):
        res = binding.encode_batch(cuda_lib, encs, [frames[s * seglen + t].copy() for s in range(nseg)], rps)
        for s in range(nseg):
            outs[s] += res[s]
    for s in range(nseg):
        rbs, rsizes, _, _ = ref.encode_sequence(frames[s * seglen:(s + 1) * seglen], w, h, seglen, kbps=600)
        assert outs[s] == rbs, "segment %d" % s
    for e in encs:
        e.close()


def test_concurrent_encoders_from_host_threads(binding, cuda_lib, ref):
    """Distinct encoders are independent (SURVEY 8(b) Threading): host threads driving their own
    streams at the same time -- one with H264E_encode, one with H264E_encode_batch -- each get the
    reference's bytes (every thread submits on its own lane of the shim)."""
    import threading
    w, h, n, nthreads = 320, 240, 6, 4
    clips = [cases.make("multi" if s & 1 else "panning", w, h, n, seed=300 + s) for s in range(nthreads)]
    outs = [b""] * nthreads
    recs = [None] * nthreads
    errs = []
    start = threading.Barrier(nthreads)

    def worker(s):
        try:
            start.wait()
            if s & 1:
                bs, _, rec = binding.encode_sequence(cuda_lib, clips[s], w, h, 60, qp=30)
                outs[s], recs[s] = bs, rec[-1]
            else:
                e = binding.Encoder(cuda_lib, w, h, 60)
                rp = e.run_param(qp=30)
                for t in range(n):
                    outs[s] += binding.encode_batch(cuda_lib, [e], [clips[s][t].copy()], [rp])[0]
                recs[s] = e.recon()
                e.close()
        except BaseException as ex:       # reported by the main thread
            errs.append((s, repr(ex)))

    th = [threading.Thread(target=worker, args=(s,)) for s in range(nthreads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for s in range(nthreads):
        rbs, _, rrec, _ = ref.encode_sequence(clips[s], w, h, 60, qp=30)
        assert outs[s] == rbs, "thread %d" % s
        assert np.array_equal(recs[s], rrec[-1]), "thread %d" % s


def test_prefetch_is_transparent(binding, cuda_lib, ref):
    """H264E_prefetch (double-buffered input upload): same bytes whether the next frame was staged, staged but a
    different frame is then encoded (miss), or not staged at all."""
    import ctypes as C
    w, h, n = 366, 250, 6                      # odd chroma width: the staged copy goes through the 2-D path
    frames = cases.make("panning", w, h, n)
    rbs, _, rrec, _ = ref.encode_sequence(frames, w, h, n, qp=26)
    enc = binding.Encoder(cuda_lib, w, h, n)
    rp = enc.run_param(qp=26)
    keep = [frames[i].copy() for i in range(n)]
    decoy = frames[0].copy()
    out = b""
    for i in range(n):
        if i + 1 < n and i != 2:
            nxt = enc.io_yuv(keep[i + 1])
            assert cuda_lib.lib.H264E_prefetch(C.c_void_p(enc.persist), C.byref(nxt)) == 0
        elif i == 2:
            nxt = enc.io_yuv(decoy)             # staged, never encoded: must be ignored
            assert cuda_lib.lib.H264E_prefetch(C.c_void_p(enc.persist), C.byref(nxt)) == 0
        out += enc.encode(keep[i], rp)
    assert out == rbs
    assert np.array_equal(enc.recon(), rrec[-1])
    enc.close()


def test_full_gop_and_idr_at_bench_size(binding, cuda_lib, ref):
    """BASELINE config 2's shape on one session: 1080p, GOP 60, 62 frames (a whole GOP, the IDR that follows it and one
    more P frame), byte for byte -- the trajectory speculation / repair machinery over 60 consecutive P frames."""
    import content
    w, h, n, gop = 1920, 1080, 62, 60
    frames = content.panning(w, h, n, seed=4242)
    variant = "_fast" if ref.have_ref("_fast") else ""
    rbs, rsz, _, _ = ref.encode_sequence(frames, w, h, gop, qp=28, want_recon=False, variant=variant)
    bs, sz, _ = binding.encode_sequence(cuda_lib, frames, w, h, gop, qp=28, want_recon=False)
    assert list(sz) == list(rsz)
    assert bs == rbs


def test_droppable_frames(binding, cuda_lib, ref):
    """Droppable frames in in-place mode + H264E_get_recon after them (see tests/test_emu_parity.py)."""
    from test_emu_parity import _droppable_sequence
    _droppable_sequence(binding, cuda_lib, ref)


# ---------------------------------------------------------------------------------------------------------------
# BASELINE configs at their STATED geometry, every unit compared with oracle/_ref
# ---------------------------------------------------------------------------------------------------------------
def _ref_units(ref, clips, w, h, gop, **kw):
    """The reference run once per unit (fresh instance per closed-GOP segment / stream), on the host cores."""
    from concurrent.futures import ThreadPoolExecutor
    variant = "_fast" if ref.have_ref("_fast") else ""
    want_recon = kw.pop("want_recon", True)

    def one(c):
        return ref.encode_sequence(c, w, h, gop, variant=variant, want_recon=want_recon, **kw)[:3]
    with ThreadPoolExecutor(max_workers=min(len(clips), os.cpu_count() or 4)) as ex:      # ctypes releases the GIL
        return list(ex.map(one, clips))


def _batch_units(binding, cuda_lib, clips, w, h, gop, **kw):
    """frame t of every unit in ONE device submission (H264E_encode_batch); returns per unit (bytes, sizes, last recon)."""
    n = len(clips)
    encs = [binding.Encoder(cuda_lib, w, h, gop) for _ in range(n)]
    rps = [e.run_param(**kw) for e in encs]
    outs, sizes = [b""] * n, [[] for _ in range(n)]
    for t in range(clips[0].shape[0]):
        res = binding.encode_batch(cuda_lib, encs, [c[t].copy() for c in clips], rps)
        for i in range(n):
            outs[i] += res[i]
            sizes[i].append(len(res[i]))
    recs = [e.recon() for e in encs]
    for e in encs:
        e.close()
    return outs, sizes, recs


def test_c3_2160p_rate_controlled_segments(binding, cuda_lib, ref):
    """BASELINE config 3 at its stated geometry: 3840x2160, --kbps 20000, GOP 30, closed-GOP segments with one fresh
    session each, encoded concurrently in one batch: bytes, per-frame sizes (the QP trajectory) and reconstruction."""
    import content
    w, h, gop, nseg, nfr = 3840, 2160, 30, 2, 8
    clips = [content.panning(w, h, nfr, seed=3000 + s) for s in range(nseg)]
    refs = _ref_units(ref, clips, w, h, gop, kbps=20000)
    outs, sizes, recs = _batch_units(binding, cuda_lib, clips, w, h, gop, kbps=20000)
    for s in range(nseg):
        assert sizes[s] == [int(x) for x in refs[s][1]], "segment %d: frame sizes (QP trajectory)" % s
        assert outs[s] == refs[s][0], "segment %d: bit stream" % s
        assert np.array_equal(recs[s], refs[s][2][-1]), "segment %d: reconstruction" % s


def test_c5_64_streams_720p(binding, cuda_lib, ref):
    """BASELINE config 5 at its stated geometry: 64 independent 1280x720 streams (distinct seeds), GOP 60, QP 28, one
    frame of every stream per submission; EVERY stream is compared with its own reference run."""
    import content
    w, h, gop, nstreams, nfr = 1280, 720, 60, 64, 4
    clips = [content.panning(w, h, nfr, seed=s) for s in range(nstreams)]
    refs = _ref_units(ref, clips, w, h, gop, qp=28)
    outs, sizes, recs = _batch_units(binding, cuda_lib, clips, w, h, gop, qp=28)
    for s in range(nstreams):
        assert outs[s] == refs[s][0], "stream %d: bit stream" % s
        assert np.array_equal(recs[s], refs[s][2][-1]), "stream %d: reconstruction" % s


def test_bench_shape_10x1080p_all_streams_and_determinism(binding, cuda_lib, ref):
    """The bench workload itself (BASELINE config 2: ten 1080p closed-GOP segments, QP 28, one submission per frame
    index): ALL ten streams compared with the reference, and the whole run repeated three times with identical
    digests -- 690 row CTAs over the resident slots in ticket order is the regime where an ordering bug of the
    wavefront protocol would show (the determinism soak stands in for compute-sanitizer, closed on this pool)."""
    import content
    w, h, gop, nseg, nfr = 1920, 1080, 60, 10, 8
    clips = [content.panning(w, h, nfr, seed=1000 + s) for s in range(nseg)]      # bench.py's clips
    refs = _ref_units(ref, clips, w, h, gop, qp=28)
    digests = []
    for rep in range(3):
        outs, sizes, recs = _batch_units(binding, cuda_lib, clips, w, h, gop, qp=28)
        for s in range(nseg):
            assert outs[s] == refs[s][0], "run %d segment %d: bit stream" % (rep, s)
            assert np.array_equal(recs[s], refs[s][2][-1]), "run %d segment %d: reconstruction" % (rep, s)
        digests.append(hashlib.md5(b"".join(outs) + b"".join(r.tobytes() for r in recs)).hexdigest())
    assert digests[0] == digests[1] == digests[2]


def test_batch_larger_than_one_submission(binding, cuda_lib, ref):
    """Any batch size is accepted: more jobs than the device holds as one wavefront submission are run as
    consecutive submissions (h264b200_encode_frames), results unchanged."""
    w, h, n = 64, 48, 3
    nstreams = 300
    clips = [cases.make("panning", w, h, n, seed=500 + (s % 7)) for s in range(nstreams)]
    refs = {}
    for k in range(7):
        refs[k] = ref.encode_sequence(clips[k], w, h, 60, qp=30)[0]
    outs, _, _ = _batch_units(binding, cuda_lib, clips, w, h, 60, qp=30)
    for s in range(nstreams):
        assert outs[s] == refs[s % 7], "stream %d" % s


KNOBS = [{}, {"H264B200_NO_SADMAP": "1"}, {"H264B200_NO_ME_PREPASS": "1"}, {"H264B200_NO_INTRA_SPEC": "1"},
         {"H264B200_NO_FAST": "1"}, {"H264B200_ME_ROUNDS": "1"}, {"H264B200_THR": "6"}, {"H264B200_THR": "40"},
         {"H264B200_NO_PREV_TRAJ": "1"}]


def test_speculation_layers_switched_off(ref):
    """Every layer of DESIGN.md 5 (SAD maps, motion-estimation pre-pass, intra speculation, decide / work fast path,
    trajectory prediction) only decides WHERE work happens: with any of them switched off, or with another threshold,
    the bytes are the reference's."""
    import subprocess
    import sys
    import knob_child
    want = []
    for kind, w, h, n, gop, kw in knob_child.CASES:
        frames = cases.make(kind, w, h, n)
        rbs, _, rrec, _ = ref.encode_sequence(frames, w, h, gop, **kw)
        want.append(hashlib.md5(rbs + rrec.tobytes()).hexdigest())
    for knob in KNOBS:
        env = dict(os.environ, **knob)
        out = subprocess.run([sys.executable, knob_child.__file__], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, (knob, out.stderr[-2000:])
        assert out.stdout.split() == want, knob


def test_heterogeneous_batch(binding, cuda_lib, ref):
    """Sessions of different picture sizes, GOP lengths (I and P frames in one submission) and quantisers in one
    H264E_encode_batch: every job has its own geometry and payload capacity (a batch-wide capacity -- the smallest
    picture's -- once failed the large ones; found by tools/stress_batch.py)."""
    specs = [("panning", 1280, 720, 3, 60, dict(qp=22)), ("noise", 32, 32, 4, 2, dict(qp=12)), ("multi", 366, 250, 4, 3, dict(qp=33)),
             ("chess", 176, 144, 2, 1, dict(kbps=200))]
    sess = []
    for kind, w, h, n, gop, kw in specs:
        frames = cases.make(kind, w, h, n)
        rbs, _, _, _ = ref.encode_sequence(frames, w, h, gop, want_recon=False, **kw)
        enc = binding.Encoder(cuda_lib, w, h, gop)
        sess.append(dict(frames=frames, n=n, ref=rbs, enc=enc, rp=enc.run_param(**kw), out=b""))
    for t in range(max(s["n"] for s in sess)):
        act = [s for s in sess if t < s["n"]]
        res = binding.encode_batch(cuda_lib, [s["enc"] for s in act], [s["frames"][t].copy() for s in act], [s["rp"] for s in act])
        for s, r in zip(act, res):
            s["out"] += r
    for i, s in enumerate(sess):
        assert s["out"] == s["ref"], specs[i]
        s["enc"].close()
