"""Drop-in boundary proof: the reference's ONLY consumer, its command-line tool
(/root/reference/src/minih264e_test.c, T:469-687), compiled UNMODIFIED against the product's declarations-only
header include/h264-lab.h and linked against libh264lab_b200.so (oracle/Makefile target _ref/encode_app_dropin),
must resolve H264E_* from the product library (CPU check) and -- on the GPU -- write the very bytes that the
reference's own build of the same tool (oracle/_ref/encode_app_ref) writes for the same command line.  The product's
re-written CLI (h264-lab_b200/encode_app) is held to the same files."""
import os
import subprocess

import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "encode_app_ref")
DROPIN_CLI = os.path.join(ROOT, "oracle", "_ref", "encode_app_dropin")
OUR_CLI = os.path.join(ROOT, "h264-lab_b200", "encode_app")


def _need(path):
    if not os.path.exists(path):
        subprocess.call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    if not os.path.exists(path):
        pytest.skip("%s not built and /root/reference absent" % os.path.basename(path))


def test_reference_cli_links_against_the_product_library():
    """H264E_sizeof / H264E_init / H264E_encode are UNDEFINED in the drop-in binary (i.e. they come from the shared
    library at load time) and the binary carries no copy of the reference's implementation."""
    _need(DROPIN_CLI)
    out = subprocess.check_output(["nm", "-D", DROPIN_CLI], text=True)
    und = {l.split()[-1] for l in out.splitlines() if " U " in l}
    assert {"H264E_sizeof", "H264E_init", "H264E_encode"} <= und
    allsyms = subprocess.check_output(["nm", DROPIN_CLI], text=True)
    for leaked in ("h264e_vlc_encode", "mb_encode", "me_search_diamond", "H264E_encode_one"):
        assert leaked not in allsyms
    needed = subprocess.check_output(["readelf", "-d", DROPIN_CLI], text=True)
    assert "libh264lab_b200.so" in needed


CLI_RUNS = [
    ("qp22", ["--gop", "20", "--qp", "22"]),          # scripts/enc_test.bat QPs, T:10-11 defaults
    ("qp33_default", []),
    ("kbps300", ["--gop", "10", "--kbps", "300"]),
    ("speed5", ["--qp", "30", "--speed", "5"]),
    ("denoise", ["--qp", "28", "--denoise"]),
]


@pytest.mark.gpu
@pytest.mark.parametrize("run", CLI_RUNS, ids=lambda r: r[0])
def test_cli_outputs_identical(run, tmp_path):
    """reference CLI (own build) == reference CLI on the product library == the product's own CLI, byte for byte."""
    _need(REF_CLI)
    _need(DROPIN_CLI)
    name, opts = run
    clip = tmp_path / "clip_cif.yuv"                    # the tools take the picture size from the file name (T:283)
    cases.make("panning", 352, 288, 12).tofile(str(clip))
    outs = {}
    for tag, exe in (("ref", REF_CLI), ("dropin", DROPIN_CLI), ("ours", OUR_CLI)):
        o = tmp_path / (tag + ".264")
        # T:205-218: every --long option takes the next argv, "--denoise" included, so flags go last
        cmd = [exe, "--input", str(clip), "--output", str(o)] + opts
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
        assert r.returncode == 0, (tag, r.stdout[-500:])
        outs[tag] = o.read_bytes()
    assert len(outs["ref"]) > 1000
    assert outs["dropin"] == outs["ref"], "reference CLI on libh264lab_b200.so differs from the reference"
    assert outs["ours"] == outs["ref"], "encode_app differs from the reference"
