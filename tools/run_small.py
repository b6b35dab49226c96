"""Developer tool: encode a few frames of one stream (for ncu captures)."""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import content
spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "h264-lab_b200", "binding.py"))
B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
L = B.Library()
w, h, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
fr = getattr(content, sys.argv[4] if len(sys.argv) > 4 else "panning")(w, h, n)
enc = B.Encoder(L, w, h, 60)
rp = enc.run_param(qp=28)
tot = 0
for i in range(n):
    tot += len(enc.encode(fr[i].copy(), rp))
print("bytes", tot)
