#!/usr/bin/env python3
"""developer helper: static SASS footprint of a kernel by source function
usage: tools/code_footprint.py [lib.so] [kernel-substring]   (needs cuobjdump + nvdisasm)"""
import re, subprocess, sys, os, tempfile, collections, bisect, glob
lib = sys.argv[1] if len(sys.argv) > 1 else 'h264-lab_b200/libh264lab_b200.so'
kern = sys.argv[2] if len(sys.argv) > 2 else 'k_encode_rows'
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
funcs = {}
def load_funcs(path):
    if path in funcs: return
    starts = []
    try: lines = open(path, errors='replace').read().split('\n')
    except OSError: funcs[path] = ([], []); return
    for i, l in enumerate(lines, 1):
        m = re.match(r'^(?:HDN?|HD_NOINLINE|static|template|__device__|__global__|inline|extern)\b.*?([A-Za-z_][A-Za-z0-9_]*)\s*\(', l)
        if m and not l.rstrip().endswith(';'): starts.append((i, m.group(1)))
    funcs[path] = ([s[0] for s in starts], [s[1] for s in starts])
def func_of(path, line):
    load_funcs(path); ls, ns = funcs[path]
    k = bisect.bisect_right(ls, line) - 1
    return ns[k] if k >= 0 else '?'
for cub in glob.glob(tmp + '/*.cubin'):
    out = subprocess.run(['nvdisasm', '-gi', cub], capture_output=True, text=True).stdout
    inside = False; cur = ('?', 0); cnt = collections.Counter(); outer = collections.Counter(); total = 0
    for l in out.split('\n'):
        if l.startswith('//---'):
            inside = ('.text.' in l and kern in l)
            continue
        if not inside: continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            chain = [(m.group(1), int(m.group(2)))] + [(a, int(b)) for a, b in re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))]
            cur = chain
            continue
        if re.match(r'\s*/\*[0-9a-f]{4,}\*/', l):
            total += 1
            names = [func_of(p, ln) if '/csrc/' in p else None for p, ln in cur] if isinstance(cur, list) else ['?']
            names = [n for n in names if n]
            cnt[names[0] if names else 'lib'] += 1
            outer[' < '.join(names[:3])] += 1
    if not total: continue
    print(f'{os.path.basename(cub)}: {kern}: {total} instrs = {total*16/1024:.1f} KB')
    for n, c in cnt.most_common(45): print(f'  {c*16/1024:6.1f} KB  {n}')
    print('  -- by inline chain (innermost < caller < caller)')
    for n, c in outer.most_common(40): print(f'  {c*16/1024:6.1f} KB  {n}')
