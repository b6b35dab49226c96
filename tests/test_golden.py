"""CPU suite: golden digests.  (1) the compiled reference still produces the committed
digests (the oracle is pinned); (2) the host-emulated device code produces them too."""
import hashlib
import json
import os

import pytest

import cases

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "manifest.json")
MAN = json.load(open(GOLDEN))


@pytest.mark.parametrize("g", [g for g in MAN["cases"] if g["width"] <= 720], ids=lambda g: g["name"])
def test_reference_matches_golden(g, ref):
    frames = cases.make(g["kind"], g["width"], g["height"], g["frames"])
    assert hashlib.md5(frames.tobytes()).hexdigest() == g["input_md5"]
    bs, sizes, rec, _ = ref.encode_sequence(frames, g["width"], g["height"], g["gop"], **g["kw"])
    assert hashlib.md5(bs).hexdigest() == g["bitstream_md5"]
    assert hashlib.md5(rec.tobytes()).hexdigest() == g["recon_md5"]
    assert [int(s) for s in sizes] == g["frame_sizes"]


@pytest.mark.parametrize("g", [g for g in MAN["cases"] if g["width"] <= 720], ids=lambda g: g["name"])
def test_emulation_matches_golden(g, binding, emu_lib):
    frames = cases.make(g["kind"], g["width"], g["height"], g["frames"])
    bs, sizes, rec = binding.encode_sequence(emu_lib, frames, g["width"], g["height"], g["gop"], **g["kw"])
    assert hashlib.md5(bs).hexdigest() == g["bitstream_md5"]
    assert hashlib.md5(rec.tobytes()).hexdigest() == g["recon_md5"]
