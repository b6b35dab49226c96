#!/bin/bash
# developer helper run on the GPU box: parity tests, per-MB phase profile, short bench
tag=${1:-x}
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${tag}_tests.log
python tools/mb_profile.py 1920 1080 3 > gpurun_out/${tag}_mbprof.log 2>&1
python tools/mb_profile.py 1920 1080 3 panning 10 > gpurun_out/${tag}_mbprof10.log 2>&1
python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench_err.log
python tools/batch_probe.py 10 8 > gpurun_out/${tag}_probe.log 2>&1
echo done
