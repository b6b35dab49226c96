#!/bin/bash
# developer helper (GPU box): bench several builds of the library back to back on the same box
# usage: tools/ab_bench.sh tag lib1.so lib2.so ...   (paths relative to h264-lab_b200/)
tag=$1; shift
for rep in 1 2; do
  for lib in "$@"; do
    echo -n "$lib: "
    H264B200_LIB=$PWD/h264-lab_b200/$lib python bench.py --steps 12 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.1f e2e %.1f dev %.2f enc %.2f'%(d['value'],d['e2e']['value'],d['kernel_ms_per_step']['device_total'],d['kernel_ms_per_step']['k_encode_rows']))"
  done
done > gpurun_out/${tag}_ab.log 2>&1
echo done
