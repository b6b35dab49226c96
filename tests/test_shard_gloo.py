"""N > 1 path on CPU: two processes (torch.distributed, gloo) shard closed-GOP segments between
them (h264-lab_b200/shard.py), each encodes its own segments -- here with the TEST-ONLY host
emulation of the device code, on the GPU box the same code runs on the CUDA library, one rank
per GPU -- and rank 0 concatenates.  The result must equal the reference run once per segment."""
import importlib.util
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _worker(rank, world, port, q, cuda=False):
    if cuda:
        os.environ["CUDA_VISIBLE_DEVICES"] = str(rank)       # one rank per GPU; H264E_init takes the current device
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cases
    S = _load("h264lab_shard", os.path.join(ROOT, "h264-lab_b200", "shard.py"))
    B = S.binding()
    lib = B.Library(os.path.join(ROOT, "h264-lab_b200", "libh264lab_b200.so") if cuda
                    else os.path.join(ROOT, "tests", "_emu", "libh264lab_emu.so"))
    w, h, gop, n = 176, 144, 4, 14
    frames = cases.make("panning", w, h, n)
    units = [frames[s:s + k] for s, k in S.split_closed_gops(n, gop)]
    out = S.encode_sharded(lib, units, w, h, gop, rank=rank, world=world, dist=dist, qp=28)
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_segments(ref, emu_lib):
    _run_two_ranks(ref, cuda=False)


@pytest.mark.gpu
def test_two_ranks_shard_segments_cuda(ref, cuda_lib):
    """The same two-rank run on the PRODUCT library, one rank per GPU (skipped on a one-GPU box)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    _run_two_ranks(ref, cuda=True)


def _run_two_ranks(ref, cuda):
    import torch.multiprocessing as mp
    S = _load("h264lab_shard", os.path.join(ROOT, "h264-lab_b200", "shard.py"))
    assert S.units_of_rank(5, 0, 2) == [0, 2, 4] and S.units_of_rank(5, 1, 2) == [1, 3]
    assert S.split_closed_gops(14, 4) == [(0, 4), (4, 4), (8, 4), (12, 2)]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, cuda)) for r in range(2)]
    for p in procs:
        p.start()
    out = None
    for _ in range(600):
        try:
            out = q.get(timeout=0.5)
            break
        except Exception:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    assert out is not None, "a rank failed"
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import cases
    w, h, gop, n = 176, 144, 4, 14
    frames = cases.make("panning", w, h, n)
    assert len(out) == 4
    for (s0, k), bs in zip(S.split_closed_gops(n, gop), out):
        rbs, _, _, _ = ref.encode_sequence(frames[s0:s0 + k], w, h, gop, qp=28)
        assert bs == rbs, "segment starting at frame %d" % s0
