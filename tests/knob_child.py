"""Child process of test_gpu_parity.py::test_speculation_layers_switched_off: the library reads its developer switches
(H264B200_NO_SADMAP ...) once per process, so every setting needs a process of its own.  Prints one md5 of bit stream +
reconstruction per case."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cases
from conftest import ROOT, load_binding

CASES = [("multi", 352, 288, 6, 6, dict(qp=28)), ("panning", 366, 250, 5, 5, dict(qp=33)),
         ("noise", 176, 144, 6, 3, dict(kbps=300))]


def main():
    binding = load_binding()
    lib = binding.Library(os.path.join(ROOT, "h264-lab_b200", "libh264lab_b200.so"))
    for kind, w, h, n, gop, kw in CASES:
        frames = cases.make(kind, w, h, n)
        bs, sizes, rec = binding.encode_sequence(lib, frames, w, h, gop, **kw)
        print(hashlib.md5(bs + rec.tobytes()).hexdigest())


if __name__ == "__main__":
    main()
