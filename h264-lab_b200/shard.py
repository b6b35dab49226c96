"""Multi-GPU sharding of the encode path: closed-GOP segments and independent
streams are independent units -- one process per GPU, unit k -> rank k mod world, NO
collective on the data path.  torch.distributed is used only to hand the finished bit streams
to rank 0, which concatenates them in unit order (what the reference's user would do with the
outputs of one encoder instance per segment).

Exactness condition: the reference must be run with the same segmentation (a fresh
H264E_init per closed GOP), because a long-lived instance carries mv_clusters, rate-control
state and the idr_pic_id toggle across IDRs.
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def binding():
    """The ctypes binding module (one instance per process: its ctypes classes must be shared)."""
    mod = sys.modules.get("h264lab_binding")
    if mod is None:
        spec = importlib.util.spec_from_file_location("h264lab_binding", os.path.join(HERE, "binding.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["h264lab_binding"] = mod
        spec.loader.exec_module(mod)
    return mod


def units_of_rank(n_units, rank, world):
    """Round-robin assignment: unit k is encoded by rank k mod world."""
    return [k for k in range(n_units) if k % world == rank]


def split_closed_gops(n_frames, gop):
    """[(first_frame, n_frames)] of the closed-GOP segments of an IPPP stream (gop 0 = one segment)."""
    if gop <= 0:
        return [(0, n_frames)]
    return [(s, min(gop, n_frames - s)) for s in range(0, n_frames, gop)]


def encode_units(library, units, width, height, gop, max_batch=64, **kw):
    """Encode `units` (list of uint8 arrays [frames, W*H*3/2]) concurrently: frame t of every unit in
    one device submission (H264E_encode_batch), one fresh encoder session per unit.
    Returns the list of bit streams."""
    B = binding()
    outs = [b"" for _ in units]
    for lo in range(0, len(units), max_batch):
        group = list(range(lo, min(lo + max_batch, len(units))))
        encs = [B.Encoder(library, width, height, gop) for _ in group]
        rps = [e.run_param(**kw) for e in encs]
        steps = max(units[i].shape[0] for i in group)
        for t in range(steps):
            live = [j for j, i in enumerate(group) if t < units[i].shape[0]]
            res = B.encode_batch(library, [encs[j] for j in live], [units[group[j]][t].copy() for j in live],
                                 [rps[j] for j in live])
            for j, bs in zip(live, res):
                outs[group[j]] += bs
        for e in encs:
            e.close()
    return outs


def encode_sharded(library, all_units, width, height, gop, rank=0, world=1, dist=None, **kw):
    """Every rank encodes its own units; rank 0 receives all bit streams (gather_object) and returns them in
    unit order (other ranks return None).  `dist` = torch.distributed (initialised) or None for world == 1."""
    mine = units_of_rank(len(all_units), rank, world)
    local = encode_units(library, [all_units[k] for k in mine], width, height, gop, **kw)
    if world == 1 or dist is None:
        return local
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((mine, local), gathered, dst=0)
    if rank != 0:
        return None
    out = [None] * len(all_units)
    for ids, streams in gathered:
        for k, bs in zip(ids, streams):
            out[k] = bs
    return out
